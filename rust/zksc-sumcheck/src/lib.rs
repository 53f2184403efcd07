//! The reference's function names and signatures, implemented on the GPU through `zksc-sys`.
//! NOT COMPILED in the build image of this repository (no Rust toolchain); see rust/README.md.
//!
//! Reference items replaced (paths relative to aagbotemi/zk-cryptography):
//!   MultilinearTrait / ComposedMultilinearTrait impls (as `Gpu<T>`)            polynomial/src/interface.rs:9-19
//!   Multilinear::{partial_evaluation, partial_evaluations, evaluation}         polynomial/src/multilinear/evaluation_form.rs:123-175
//!   Sumcheck::{new, poly_sum, prove, verify}, SumcheckProof                    sumcheck/src/sumcheck.rs:6-95
//!   composed::ComposedSumcheck::{new, calculate_poly_sum, prove, verify}       sumcheck/src/composed/composed_sumcheck.rs:9-95
//!   MultiComposedSumcheckProver::{calculate_poly_sum, prove, prove_partial}    sumcheck/src/composed/multi_composed_sumcheck.rs:36-62
//!   MultiComposedSumcheckVerifier::{verify, verify_partial}, SubClaim          sumcheck/src/composed/multi_composed_sumcheck.rs:18-22,126-149
//!   ComposedSumcheckProof::to_bytes                                            sumcheck/src/composed/multi_composed_sumcheck.rs:24-32
//!   GKRProtocol::prove                                                         gkr/src/protocol.rs:21-113
//! One visibility change is needed in the reference for an out-of-crate shim: `ComposedMultilinear::polys` is private
//! (polynomial/src/composed/composed_multilinear.rs:8); add `pub fn polys(&self) -> &[Multilinear<F>] { &self.polys }` there
//!, which is what `tables_of` below calls.
//! Errors: the reference panics on bad shapes (evaluation_form.rs:16-20) -> ZKSC_ERR_SHAPE is turned back into a panic;
//! the prover always returns Ok (multi_composed_sumcheck.rs:119).
use ark_ff::BigInt;
use ark_test_curves::bls12_381::Fr;
use polynomial::{ComposedMultilinear, ComposedMultilinearTrait, Multilinear, MultilinearTrait, SparseUnivariatePolynomial, UnivariateMonomial};
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

/// The factor tables of a product (needs the accessor named in the module documentation).
fn tables_of(p: &ComposedMultilinear<Fr>) -> &[Multilinear<Fr>] {
    p.polys()
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
        let ptrs: Vec<*const u64> = poly.iter().flat_map(|p| tables_of(p).iter().map(|m| limbs(&m.evaluations))).collect();
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

    /// multi_composed_sumcheck.rs:47-54 (the tables are absorbed first: converted to canonical big-endian bytes on the device)
    pub fn prove(poly: &Vec<ComposedMultilinear<Fr>>, sum: &Fr) -> Result<(ComposedSumcheckProof, Vec<Fr>), &'static str> {
        Self::prove_with(poly, sum, ZKSC_PROTO_MULTI_FULL)
    }

    /// multi_composed_sumcheck.rs:56-62 (fresh transcript, the claimed sum is absorbed, tables are not)
    pub fn prove_partial(poly: &Vec<ComposedMultilinear<Fr>>, sum: &Fr) -> Result<(ComposedSumcheckProof, Vec<Fr>), &'static str> {
        Self::prove_with(poly, sum, ZKSC_PROTO_MULTI_PARTIAL)
    }

    fn prove_with(poly: &Vec<ComposedMultilinear<Fr>>, sum: &Fr, protocol: i32) -> Result<(ComposedSumcheckProof, Vec<Fr>), &'static str> {
        let t = Tables::upload(poly);
        let (msgs, lens, chal, stride) = t.prove(protocol, Some(sum));
        let n = t.n_vars as usize;
        let round_polys = (0..n).map(|r| sparse_of(&msgs[r * stride * 4..], lens[r] as usize)).collect();
        Ok((ComposedSumcheckProof { round_polys, sum: *sum }, chal))
    }
}

impl Tables {
    /// zksc_prove: the reference's round loop with the transcript on the host (inside the library), tables on the GPU
    fn prove(&self, protocol: i32, sum: Option<&Fr>) -> (Vec<u64>, Vec<u32>, Vec<Fr>, usize) {
        let ctx = context();
        let n = self.n_vars as usize;
        let stride = unsafe { zksc_msg_stride(protocol, self.deg.len() as u32, self.deg.as_ptr()) } as usize;
        let mut msgs = vec![0u64; n * stride * 4];
        let mut lens = vec![0u32; n];
        let mut chal = vec![0u64; n * 4];
        let s = sum.map(|s| s.0 .0); // Montgomery limbs of the caller's claimed sum
        let sp = s.as_ref().map_or(ptr::null(), |s| s.as_ptr());
        let rc = unsafe { zksc_prove(self.h, protocol, sp, msgs.as_mut_ptr(), lens.as_mut_ptr(), chal.as_mut_ptr()) };
        check(ctx, rc);
        (msgs, lens, chal.chunks(4).map(fr).collect(), stride)
    }
    /// sum_p prod_k f_{p,k}(points): the verifiers' oracle check, n folds on the device (zksc_evaluate)
    fn evaluate(&self, points: &[Fr]) -> Fr {
        let mut out = [0u64; 4];
        check(context(), unsafe { zksc_evaluate(self.h, limbs(points), out.as_mut_ptr()) });
        fr(&out)
    }
}

fn sparse_of(msg: &[u64], monomials: usize) -> SparseUnivariatePolynomial<Fr> {
    SparseUnivariatePolynomial {
        monomial: (0..monomials).map(|m| UnivariateMonomial { coeff: fr(&msg[8 * m..8 * m + 4]), pow: fr(&msg[8 * m + 4..8 * m + 8]) }).collect(),
    }
}

impl ComposedSumcheckProof {
    /// multi_composed_sumcheck.rs:24-32
    pub fn to_bytes(&self) -> Vec<u8> {
        self.round_polys.iter().flat_map(|p| p.to_bytes()).collect()
    }
    /// (coeff, pow) pairs of every round, flattened, in the layout zksc_verify_rounds reads
    fn raw(&self) -> (Vec<u64>, Vec<u32>, usize) {
        let stride = 2 * self.round_polys.iter().map(|p| p.monomial.len()).max().unwrap_or(1).max(1);
        let mut msgs = vec![0u64; self.round_polys.len() * stride * 4];
        let mut lens = Vec::with_capacity(self.round_polys.len());
        for (r, p) in self.round_polys.iter().enumerate() {
            for (m, mono) in p.monomial.iter().enumerate() {
                let o = (r * stride + 2 * m) * 4;
                msgs[o..o + 4].copy_from_slice(&mono.coeff.0 .0);
                msgs[o + 4..o + 8].copy_from_slice(&mono.pow.0 .0);
            }
            lens.push(p.monomial.len() as u32);
        }
        (msgs, lens, stride)
    }
}

/// multi_composed_sumcheck.rs:18-22
#[derive(Debug)]
pub struct SubClaim {
    pub sum: Fr,
    pub challenges: Vec<Fr>,
}

/// The transcript half of every verifier (zksc_verify_rounds, host only): replay, p(0) + p(1) chain, sub-claim.
fn verify_rounds(protocol: i32, sum: &Fr, msgs: &[u64], lens: &[u32], stride: usize, prefix: &[u8]) -> Result<SubClaim, &'static str> {
    let n = lens.len();
    let mut sub = [0u64; 4];
    let mut chal = vec![0u64; n.max(1) * 4];
    let rc = unsafe {
        zksc_verify_rounds(protocol, n as u32, stride as u32, sum.0 .0.as_ptr(), msgs.as_ptr(), lens.as_ptr(),
                           if prefix.is_empty() { ptr::null() } else { prefix.as_ptr() }, prefix.len(), sub.as_mut_ptr(), chal.as_mut_ptr())
    };
    if rc == ZKSC_ERR_VERIFY {
        return Err("Verification failed"); // multi_composed_sumcheck.rs:170
    }
    assert_eq!(rc, ZKSC_OK, "zksc_verify_rounds: malformed proof");
    Ok(SubClaim { sum: fr(&sub), challenges: chal[..4 * n].chunks(4).map(fr).collect() })
}

pub struct MultiComposedSumcheckVerifier;

impl MultiComposedSumcheckVerifier {
    /// multi_composed_sumcheck.rs:126-142: tables absorbed, rounds replayed, oracle check on the device
    pub fn verify(poly: &Vec<ComposedMultilinear<Fr>>, proof: &ComposedSumcheckProof) -> Result<bool, &'static str> {
        let prefix: Vec<u8> = poly.iter().flat_map(|p| p.to_bytes()).collect(); // composed_poly_to_bytes, sumcheck/src/utils.rs:53-59
        let (msgs, lens, stride) = proof.raw();
        let sub = verify_rounds(ZKSC_PROTO_MULTI_FULL, &proof.sum, &msgs, &lens, stride, &prefix)?;
        let t = Tables::upload(poly);
        Ok(t.evaluate(&sub.challenges) == sub.sum)
    }
    /// multi_composed_sumcheck.rs:143-149 -- what GKR calls (gkr/src/protocol.rs:162)
    pub fn verify_partial(proof: &ComposedSumcheckProof) -> Result<SubClaim, &'static str> {
        let (msgs, lens, stride) = proof.raw();
        verify_rounds(ZKSC_PROTO_MULTI_PARTIAL, &proof.sum, &msgs, &lens, stride, &[])
    }
}

/// sumcheck/src/sumcheck.rs:6-95
pub struct Sumcheck {
    poly: Multilinear<Fr>,
    pub sum: Fr,
}
pub struct SumcheckProof {
    poly: Multilinear<Fr>,
    sum: Fr,
    univariate_poly: Vec<Multilinear<Fr>>,
}
impl Sumcheck {
    pub fn new(poly: Multilinear<Fr>) -> Self {
        Sumcheck { poly, sum: Default::default() }
    }
    /// :25-27
    pub fn poly_sum(&mut self) {
        let t = Tables::upload(&[ComposedMultilinear::new(vec![self.poly.clone()])]);
        let mut out = [0u64; 4];
        check(context(), unsafe { zksc_poly_sum(t.h, out.as_mut_ptr()) });
        self.sum = fr(&out);
    }
    /// :29-61: the round message is the two half sums
    pub fn prove(&self) -> (SumcheckProof, Vec<Fr>) {
        let t = Tables::upload(&[ComposedMultilinear::new(vec![self.poly.clone()])]);
        let (msgs, _lens, chal, stride) = t.prove(ZKSC_PROTO_SUMCHECK, Some(&self.sum));
        let univariate_poly = (0..t.n_vars as usize)
            .map(|r| Multilinear::new(vec![fr(&msgs[r * stride * 4..r * stride * 4 + 4]), fr(&msgs[r * stride * 4 + 4..r * stride * 4 + 8])]))
            .collect();
        (SumcheckProof { poly: self.poly.clone(), sum: self.sum, univariate_poly }, chal)
    }
    /// :63-95
    pub fn verify(&self, proof: &SumcheckProof) -> bool {
        let msgs: Vec<u64> = proof.univariate_poly.iter().flat_map(|u| u.evaluations.iter().flat_map(|e| e.0 .0)).collect();
        let lens = vec![2u32; proof.univariate_poly.len()];
        match verify_rounds(ZKSC_PROTO_SUMCHECK, &proof.sum, &msgs, &lens, 2, &[]) {
            Err(_) => false,
            Ok(sub) => evaluation(&proof.poly, &sub.challenges) == sub.sum,
        }
    }
}

/// sumcheck/src/composed/composed_sumcheck.rs (same module path as in the reference: its proof type shares a name with the
/// multi-composed one)
pub mod composed {
    use super::*;

    #[derive(Debug, Clone)]
    pub struct ComposedSumcheck {
        pub poly: ComposedMultilinear<Fr>,
        pub sum: Fr,
    }
    pub struct ComposedSumcheckProof {
        pub poly: ComposedMultilinear<Fr>,
        pub round_polys: Vec<Vec<Fr>>,
    }
    impl ComposedSumcheck {
        pub fn new(poly: ComposedMultilinear<Fr>) -> Self {
            ComposedSumcheck { poly, sum: Default::default() }
        }
        /// :28-30
        pub fn calculate_poly_sum(poly: &ComposedMultilinear<Fr>) -> Fr {
            MultiComposedSumcheckProver::calculate_poly_sum(&vec![poly.clone()])
        }
        /// :32-67: no sum absorbed; the round message is the evaluations at 0..=d
        pub fn prove(&self) -> (ComposedSumcheckProof, Vec<Fr>) {
            let t = Tables::upload(std::slice::from_ref(&self.poly));
            let (msgs, lens, chal, stride) = t.prove(ZKSC_PROTO_COMPOSED, None);
            let round_polys = (0..t.n_vars as usize)
                .map(|r| (0..lens[r] as usize).map(|i| fr(&msgs[(r * stride + i) * 4..(r * stride + i) * 4 + 4])).collect())
                .collect();
            (ComposedSumcheckProof { poly: self.poly.clone(), round_polys }, chal)
        }
        /// :69-95
        pub fn verify(&self, proof: &ComposedSumcheckProof, sum: Fr) -> bool {
            let stride = proof.round_polys.iter().map(|r| r.len()).max().unwrap_or(1).max(1);
            let mut msgs = vec![0u64; proof.round_polys.len() * stride * 4];
            for (r, rp) in proof.round_polys.iter().enumerate() {
                for (i, e) in rp.iter().enumerate() {
                    msgs[(r * stride + i) * 4..(r * stride + i) * 4 + 4].copy_from_slice(&e.0 .0);
                }
            }
            let lens: Vec<u32> = proof.round_polys.iter().map(|r| r.len() as u32).collect();
            match verify_rounds(ZKSC_PROTO_COMPOSED, &sum, &msgs, &lens, stride, &[]) {
                Err(_) => false,
                Ok(sub) => Tables::upload(std::slice::from_ref(&proof.poly)).evaluate(&sub.challenges) == sub.sum,
            }
        }
    }
}

/// The reference's traits (polynomial/src/interface.rs:9-19) on the GPU.  `polynomial` already implements them for its own types
/// on the CPU, and a foreign trait cannot be implemented twice for a foreign type, so the GPU implementations live on a
/// transparent wrapper: `Gpu(&poly).evaluation(&points)`.  (Inside the reference workspace the bodies below simply replace the
/// CPU bodies: INTEGRATION.md section 1.)
pub struct Gpu<T>(pub T);

impl MultilinearTrait<Fr> for Gpu<Multilinear<Fr>> {
    fn partial_evaluation(&self, eval_point: &Fr, variable_index: &usize) -> Self {
        Gpu(partial_evaluation(&self.0, eval_point, variable_index))
    }
    fn partial_evaluations(&self, points: &[Fr], variable_indices: &Vec<usize>) -> Self {
        assert_eq!(points.len(), variable_indices.len(), "The length of evaluation_points and variable_indices should be the same"); // evaluation_form.rs:146-152
        let mut cur = self.0.clone();
        for (p, k) in points.iter().zip(variable_indices) {
            cur = partial_evaluation(&cur, p, k);
        }
        Gpu(cur)
    }
    fn evaluation(&self, evaluation_points: &[Fr]) -> Fr {
        evaluation(&self.0, evaluation_points)
    }
}

impl MultilinearTrait<Fr> for Gpu<ComposedMultilinear<Fr>> {
    /// composed_multilinear.rs:63-75
    fn partial_evaluation(&self, eval_point: &Fr, variable_index: &usize) -> Self {
        Gpu(ComposedMultilinear::new(tables_of(&self.0).iter().map(|m| partial_evaluation(m, eval_point, variable_index)).collect()))
    }
    fn partial_evaluations(&self, points: &[Fr], variable_indices: &Vec<usize>) -> Self {
        assert_eq!(points.len(), variable_indices.len(), "The length of evaluation_points and variable_index should be the same"); // :82-88
        let mut cur = Gpu(self.0.clone());
        for (p, k) in points.iter().zip(variable_indices) {
            cur = cur.partial_evaluation(p, k);
        }
        cur
    }
    /// composed_multilinear.rs:52-61: the product of the factors' evaluations, n folds of every table on the device
    fn evaluation(&self, points: &[Fr]) -> Fr {
        Tables::upload(std::slice::from_ref(&self.0)).evaluate(points)
    }
}

impl ComposedMultilinearTrait<Fr> for Gpu<ComposedMultilinear<Fr>> {
    /// composed_multilinear.rs:105-111
    fn element_wise_product(&self) -> Vec<Fr> {
        elementwise(tables_of(&self.0), 2)
    }
    /// :113-119
    fn element_wise_add(&self) -> Vec<Fr> {
        elementwise(tables_of(&self.0), 0)
    }
    /// :101-103
    fn max_degree(&self) -> usize {
        tables_of(&self.0).len()
    }
}

/// fold the factor tables with zksc_ml_elementwise (op 0 = add, 2 = multiply)
fn elementwise(tabs: &[Multilinear<Fr>], op: i32) -> Vec<Fr> {
    let ctx = context();
    let n = tabs[0].evaluations.len();
    let mut acc: Vec<u64> = tabs[0].evaluations.iter().flat_map(|e| e.0 .0).collect();
    for t in &tabs[1..] {
        let mut out = vec![0u64; n * 4];
        check(ctx, unsafe { zksc_ml_elementwise(ctx, op, acc.as_ptr(), limbs(&t.evaluations), n as u64, out.as_mut_ptr()) });
        acc = out;
    }
    acc.chunks(4).map(fr).collect()
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

/// A layered add/mul circuit of ANY power-of-two layer widths, resident on the device, proved in time linear in its gates
/// (`zksc_circuit_*`, `zksc_gkr_prove_linear`; BASELINE config 4 as written: "width 2^20, depth 8").  `circuit::Circuit` holds only
/// the pyramid shape (layer i: 2^i gates, circuit/src/utils.rs:1-34); on such circuits `prove` returns what `gkr_prove` returns,
/// i.e. the fields of `GKRProtocol::prove`'s `GKRProof` (gkr/src/protocol.rs:10-15, :21-113), byte for byte.
pub struct LayeredCircuit {
    h: *mut zksc_circuit,
    log_width: Vec<u32>,
}
impl LayeredCircuit {
    /// `log_width[i]` = log2(gates of layer i), output layer first, last entry = log2(inputs);
    /// `gates`: per layer the gates as (is_mul, in0, in1), inputs being wire indices of the layer below.
    pub fn new(log_width: &[u32], gates: &[Vec<(bool, usize, usize)>]) -> Self {
        assert_eq!(log_width.len(), gates.len() + 1, "one width per layer plus the input layer");
        let gate_type: Vec<u8> = gates.iter().flatten().map(|g| g.0 as u8).collect();
        let in0: Vec<u32> = gates.iter().flatten().map(|g| g.1 as u32).collect();
        let in1: Vec<u32> = gates.iter().flatten().map(|g| g.2 as u32).collect();
        for (i, g) in gates.iter().enumerate() {
            assert_eq!(g.len(), 1usize << log_width[i], "layer {} must have 2^{} gates", i, log_width[i]);
        }
        let ctx = context();
        let mut h = ptr::null_mut();
        let rc = unsafe { zksc_circuit_create(ctx, gates.len() as u32, log_width.as_ptr(), gate_type.as_ptr(), in0.as_ptr(), in1.as_ptr(), &mut h) };
        check(ctx, rc);
        LayeredCircuit { h, log_width: log_width.to_vec() }
    }
    /// `Circuit::evaluation` (circuit/src/circuit.rs:32-55) on the device; returns the output layer, keeps every layer in HBM.
    pub fn evaluate(&mut self, inputs: &[Fr]) -> Vec<Fr> {
        assert_eq!(inputs.len(), 1usize << self.log_width[self.log_width.len() - 1], "the input layer has 2^log_width values");
        let mut out = vec![0u64; 4 << self.log_width[0]];
        let rc = unsafe { zksc_circuit_evaluate(self.h, limbs(inputs), out.as_mut_ptr()) };
        check(context(), rc);
        out.chunks(4).map(fr).collect()
    }
    /// `GKRProtocol::prove` of the latest `evaluate`: (sumcheck_proofs, wb_s, wc_s, w_0).
    pub fn prove(&mut self) -> (Vec<ComposedSumcheckProof>, Vec<Fr>, Vec<Fr>, Vec<Fr>) {
        let l = self.log_width.len() - 1;
        let rounds = unsafe { zksc_circuit_total_rounds(self.h) } as usize;
        let n0 = std::cmp::max(2usize, 1usize << self.log_width[0]);
        let (mut w0, mut sums, mut wb, mut wc) = (vec![0u64; 4 * n0], vec![0u64; 4 * l], vec![0u64; 4 * l], vec![0u64; 4 * l]);
        let (mut msgs, mut lens, mut chal) = (vec![0u64; rounds * 6 * 4], vec![0u32; rounds], vec![0u64; rounds * 4]);
        let rc = unsafe {
            zksc_gkr_prove_linear(self.h, w0.as_mut_ptr(), sums.as_mut_ptr(), wb.as_mut_ptr(), wc.as_mut_ptr(), msgs.as_mut_ptr(), lens.as_mut_ptr(), chal.as_mut_ptr())
        };
        check(context(), rc);
        let mut proofs = Vec::with_capacity(l);
        let mut off = 0usize;
        for li in 0..l {
            let n = 2 * self.log_width[li + 1] as usize;
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
}
impl Drop for LayeredCircuit {
    fn drop(&mut self) {
        unsafe { zksc_circuit_free(self.h) };
    }
}
