//! The reference's function names and signatures, implemented on the GPU through `zksc-sys`.
//! NOT COMPILED in the build image of this repository (no Rust toolchain); see rust/README.md.
//!
//! Reference items replaced (paths relative to aagbotemi/zk-cryptography):
//!   MultiComposedSumcheckProver::{calculate_poly_sum, prove_partial}   sumcheck/src/composed/multi_composed_sumcheck.rs:37-62
//!   Multilinear::{partial_evaluation, evaluation}                      polynomial/src/multilinear/evaluation_form.rs:123-175
//! Errors: the reference panics on bad shapes (evaluation_form.rs:16-20) -> ZKSC_ERR_SHAPE is turned back into a panic;
//! the prover always returns Ok (multi_composed_sumcheck.rs:119).
use ark_ff::BigInt;
use ark_test_curves::bls12_381::Fr;
use polynomial::{ComposedMultilinear, Multilinear, SparseUnivariatePolynomial, UnivariateMonomial};
use std::{cell::RefCell, ffi::CStr, ptr};
use zksc_sys::*;

/// One context (CUDA device 0, one stream) per host thread: a zksc context is not thread-safe (include/zksc.h).
struct Ctx(*mut zksc_ctx);
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { zksc_ctx_destroy(self.0) };
    }
}
thread_local! { static CTX: RefCell<Option<Ctx>> = RefCell::new(None); }

fn context() -> *mut zksc_ctx {
    CTX.with(|c| {
        let mut c = c.borrow_mut();
        if c.is_none() {
            let mut h = ptr::null_mut();
            let rc = unsafe { zksc_ctx_create(0, &mut h) };
            assert_eq!(rc, ZKSC_OK, "zksc_ctx_create: {} (there is no CPU fallback)", last_error(ptr::null()));
            *c = Some(Ctx(h));
        }
        c.as_ref().unwrap().0
    })
}

fn last_error(ctx: *const zksc_ctx) -> String {
    unsafe { CStr::from_ptr(zksc_last_error(ctx)) }.to_string_lossy().into_owned()
}

fn check(ctx: *mut zksc_ctx, rc: i32) {
    if rc == ZKSC_ERR_SHAPE {
        panic!("{}", last_error(ctx)); // the reference's assert!/panic! sites
    }
    assert_eq!(rc, ZKSC_OK, "{}", last_error(ctx));
}

/// `&[Fr]` -> `*const u64`: Fp256<MontBackend<FrConfig, 4>> is a BigInt<4> = [u64; 4] in Montgomery form: no conversion.
fn limbs(x: &[Fr]) -> *const u64 {
    x.as_ptr() as *const u64
}
fn fr(l: &[u64]) -> Fr {
    Fr::new_unchecked(BigInt::new([l[0], l[1], l[2], l[3]]))
}

/// Device-resident tables of `Vec<ComposedMultilinear<Fr>>`; freed on drop.
struct Tables {
    h: *mut zksc_tables,
    n_vars: u32,
    deg: Vec<u32>,
}
impl Tables {
    fn upload(poly: &[ComposedMultilinear<Fr>]) -> Self {
        let ctx = context();
        let n_vars = poly[0].n_vars() as u32; // multi_composed_sumcheck.rs:76
        let deg: Vec<u32> = poly.iter().map(|p| p.max_degree() as u32).collect();
        let ptrs: Vec<*const u64> = poly.iter().flat_map(|p| p.polys.iter().map(|m| limbs(&m.evaluations))).collect();
        let mut h = ptr::null_mut();
        let rc = unsafe { zksc_tables_upload(ctx, n_vars, 1, deg.len() as u32, deg.as_ptr(), ptrs.as_ptr(), &mut h) };
        check(ctx, rc);
        Tables { h, n_vars, deg }
    }
}
impl Drop for Tables {
    fn drop(&mut self) {
        unsafe { zksc_tables_free(self.h) };
    }
}

pub struct ComposedSumcheckProof {
    pub round_polys: Vec<SparseUnivariatePolynomial<Fr>>,
    pub sum: Fr,
}

pub struct MultiComposedSumcheckProver;

impl MultiComposedSumcheckProver {
    /// multi_composed_sumcheck.rs:37-45
    pub fn calculate_poly_sum(poly: &Vec<ComposedMultilinear<Fr>>) -> Fr {
        let t = Tables::upload(poly);
        let mut out = [0u64; 4];
        check(context(), unsafe { zksc_poly_sum(t.h, out.as_mut_ptr()) });
        fr(&out)
    }

    /// multi_composed_sumcheck.rs:56-62 (fresh transcript, the claimed sum is absorbed, tables are not)
    pub fn prove_partial(poly: &Vec<ComposedMultilinear<Fr>>, sum: &Fr) -> Result<(ComposedSumcheckProof, Vec<Fr>), &'static str> {
        let ctx = context();
        let t = Tables::upload(poly);
        let n = t.n_vars as usize;
        let stride = unsafe { zksc_msg_stride(ZKSC_PROTO_MULTI_PARTIAL, t.deg.len() as u32, t.deg.as_ptr()) } as usize;
        let mut msgs = vec![0u64; n * stride * 4];
        let mut lens = vec![0u32; n];
        let mut chal = vec![0u64; n * 4];
        let s = sum.0 .0; // Montgomery limbs of the caller's claimed sum
        let rc = unsafe { zksc_prove(t.h, ZKSC_PROTO_MULTI_PARTIAL, s.as_ptr(), msgs.as_mut_ptr(), lens.as_mut_ptr(), chal.as_mut_ptr()) };
        check(ctx, rc);
        let round_polys = (0..n)
            .map(|r| SparseUnivariatePolynomial {
                monomial: (0..lens[r] as usize)
                    .map(|m| {
                        let o = (r * stride + 2 * m) * 4;
                        UnivariateMonomial { coeff: fr(&msgs[o..o + 4]), pow: fr(&msgs[o + 4..o + 8]) }
                    })
                    .collect(),
            })
            .collect();
        let challenges = (0..n).map(|r| fr(&chal[4 * r..4 * r + 4])).collect();
        Ok((ComposedSumcheckProof { round_polys, sum: *sum }, challenges))
    }
}

/// The round-level seam for callers that keep their own transcript: evaluations of round j, then bind.
pub struct RoundProver(Tables);
impl RoundProver {
    pub fn new(poly: &[ComposedMultilinear<Fr>]) -> Self {
        RoundProver(Tables::upload(poly))
    }
    /// p.partial_evaluation(F::from(i), 0).element_wise_product().iter().sum() for i = 0..=deg, every product (:81-89)
    pub fn round_evals(&mut self) -> Vec<Fr> {
        let e: usize = self.0.deg.iter().map(|d| *d as usize + 1).sum();
        let mut out = vec![0u64; e * 4];
        check(context(), unsafe { zksc_round_evals(self.0.h, out.as_mut_ptr()) });
        out.chunks(4).map(fr).collect()
    }
    /// current_poly[i].partial_evaluation(&r, &0) for every table (:103-105); fused into the next round on the device
    pub fn bind(&mut self, r: &Fr) {
        check(context(), unsafe { zksc_bind(self.0.h, r.0 .0.as_ptr()) });
    }
}

/// Multilinear::partial_evaluation(eval_point, variable_index)   evaluation_form.rs:123-141
pub fn partial_evaluation(m: &Multilinear<Fr>, eval_point: &Fr, variable_index: &usize) -> Multilinear<Fr> {
    let n = m.evaluations.len();
    let mut out = vec![0u64; n / 2 * 4];
    let ctx = context();
    let rc = unsafe { zksc_ml_partial_evaluation(ctx, limbs(&m.evaluations), n as u64, eval_point.0 .0.as_ptr(), *variable_index as u32, out.as_mut_ptr()) };
    check(ctx, rc);
    Multilinear::new(out.chunks(4).map(fr).collect())
}

/// Multilinear::evaluation(points)   evaluation_form.rs:162-175
pub fn evaluation(m: &Multilinear<Fr>, points: &[Fr]) -> Fr {
    let mut out = [0u64; 4];
    let ctx = context();
    let rc = unsafe { zksc_ml_evaluation(ctx, limbs(&m.evaluations), m.evaluations.len() as u64, limbs(points), points.len() as u32, out.as_mut_ptr()) };
    check(ctx, rc);
    fr(&out)
}

/// GKRProtocol::prove   gkr/src/protocol.rs:21-113 -- the whole proof in one library call (`zksc_gkr_prove`): layer tables
/// built in HBM, layer sumchecks on the GPU, W(b*), W(c*) on the GPU, outer Fiat-Shamir transcript on the host.
/// `gates`: per circuit layer (output layer first) the gates as (is_mul, in0, in1)   circuit/src/{gate,circuit}.rs
/// `circuit_evaluation`: `Circuit::evaluation`'s result (circuit.rs:32-55), [0] = output ... [n] = input.
/// Returns (sumcheck_proofs, wb_s, wc_s, w_0) -- the fields of the reference's `GKRProof` (protocol.rs:10-15).
pub fn gkr_prove(gates: &[Vec<(bool, usize, usize)>], circuit_evaluation: &Vec<Vec<Fr>>) -> (Vec<ComposedSumcheckProof>, Vec<Fr>, Vec<Fr>, Vec<Fr>) {
    let l = gates.len();
    assert_eq!(circuit_evaluation.len(), l + 1, "circuit_evaluation must hold one vector per layer plus the input");
    let n_gates: Vec<u32> = gates.iter().map(|g| g.len() as u32).collect();
    let gate_type: Vec<u8> = gates.iter().flatten().map(|g| g.0 as u8).collect();
    let in0: Vec<u32> = gates.iter().flatten().map(|g| g.1 as u32).collect();
    let in1: Vec<u32> = gates.iter().flatten().map(|g| g.2 as u32).collect();
    let vals: Vec<*const u64> = circuit_evaluation.iter().map(|v| limbs(v)).collect();
    let vlen: Vec<u64> = circuit_evaluation.iter().map(|v| v.len() as u64).collect();
    let rounds = unsafe { zksc_gkr_total_rounds(l as u32) } as usize;
    let (mut w0, mut sums, mut wb, mut wc) = (vec![0u64; 8], vec![0u64; 4 * l], vec![0u64; 4 * l], vec![0u64; 4 * l]);
    let (mut msgs, mut lens, mut chal) = (vec![0u64; rounds * 6 * 4], vec![0u32; rounds], vec![0u64; rounds * 4]);
    let ctx = context();
    let rc = unsafe {
        zksc_gkr_prove(ctx, l as u32, n_gates.as_ptr(), gate_type.as_ptr(), in0.as_ptr(), in1.as_ptr(), vals.as_ptr(), vlen.as_ptr(), w0.as_mut_ptr(),
                       sums.as_mut_ptr(), wb.as_mut_ptr(), wc.as_mut_ptr(), msgs.as_mut_ptr(), lens.as_mut_ptr(), chal.as_mut_ptr())
    };
    check(ctx, rc);
    let mut proofs = Vec::with_capacity(l);
    let mut off = 0usize;
    for li in 0..l {
        let n = 2 * (li + 1);
        let round_polys = (off..off + n).map(|r| SparseUnivariatePolynomial {
            monomial: (0..lens[r] as usize).map(|m| {
                let o = (r * 6 + 2 * m) * 4;
                UnivariateMonomial { coeff: fr(&msgs[o..o + 4]), pow: fr(&msgs[o + 4..o + 8]) }
            }).collect(),
        }).collect();
        proofs.push(ComposedSumcheckProof { round_polys, sum: fr(&sums[4 * li..4 * li + 4]) });
        off += n;
    }
    (proofs, wb.chunks(4).map(fr).collect(), wc.chunks(4).map(fr).collect(), w0.chunks(4).map(fr).collect())
}
