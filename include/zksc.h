/* zksc -- B200-native sumcheck prover: the drop-in C ABI.
 *
 * The reference (aagbotemi/zk-cryptography) is pure Rust with no FFI seam of its own; the seam this
 * library replaces is the Rust API of its `polynomial`, `sumcheck` and `fiat_shamir` crates
 * (SURVEY.md section 8b).  Every entry point below names the reference item it stands in for
 * (paths relative to the reference root).  A `build.rs`-built Rust shim (rust/, source only -- there
 * is no Rust toolchain in this image) binds exactly these symbols; see INTEGRATION.md.
 *
 * Conventions
 *  - A field element is BLS12-381 Fr in ark-ff 0.4.2's in-memory form: 4 x uint64_t little-endian
 *    limbs, Montgomery form (R = 2^256) -- a `&[Fr]` can be passed as `const uint64_t*` unchanged.
 *  - Every function returns 0 (ZKSC_OK) or a negative ZKSC_ERR_* code; zksc_last_error() gives text.
 *    Nothing unwinds across the boundary.  Shape errors that `assert!`/`panic!` in the reference
 *    (evaluation_form.rs:16-20, composed_multilinear.rs:15) come back as ZKSC_ERR_SHAPE.
 *  - A context is bound to one CUDA device and one host thread at a time.  There is NO CPU fallback:
 *    without a usable CUDA device every compute entry point fails with ZKSC_ERR_NO_DEVICE.
 */
#ifndef ZKSC_H
#define ZKSC_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZKSC_OK 0
#define ZKSC_ERR_NO_DEVICE (-1)
#define ZKSC_ERR_CUDA (-2)
#define ZKSC_ERR_SHAPE (-3)
#define ZKSC_ERR_STATE (-4)
#define ZKSC_ERR_UNSUPPORTED (-5)
#define ZKSC_ERR_OOM (-6)
#define ZKSC_ERR_VERIFY (-7) /* the reference's Err("Verification failed"), multi_composed_sumcheck.rs:170 */
#define ZKSC_ERR_COMM (-8)

#define ZKSC_MAX_DEGREE 8    /* factors per product */
#define ZKSC_MAX_PRODUCTS 8

typedef struct zksc_ctx zksc_ctx;
typedef struct zksc_tables zksc_tables;

/* ---- library / context ----------------------------------------------------------------------- */
const char* zksc_version(void);
/* Number of usable CUDA devices (0 when none; never fails). */
int zksc_device_count(void);
/* Create a context on CUDA device `device`.  Fails with ZKSC_ERR_NO_DEVICE when there is none. */
int zksc_ctx_create(int device, zksc_ctx** out);
/* One context over several devices of ONE process (the reference's caller, gkr/src/protocol.rs:85, is a single process): `devices` are
 * n_devices distinct CUDA device ordinals, a power of two up to 8, with peer access to each other (NVLink / NVSwitch).  Every table of
 * a handle made on such a context is sharded over the devices (device g holds the entries i = g mod n_devices); one host thread -- the
 * caller's -- runs the one Fiat-Shamir transcript, adds the devices' partial evaluations (read from pinned host memory) and posts
 * every challenge to all devices; there is no per-round exchange between the devices, and once a table is down to
 * zksc_ctx_gather_entries() entries every device pulls the other shards over NVLink and the rest runs replicated.
 * zksc_tables_upload sends every table across PCIe once (staged on the first device, each device picks its shard over NVLink).
 * Supported on such a context: zksc_tables_upload / _synth / _reset / _free / _vars_left, zksc_poly_sum, zksc_round_evals, zksc_bind,
 * zksc_residual (once the shards are exhausted), zksc_tables_to_bytes, zksc_prove (all four protocols), zksc_evaluate, the stand-alone
 * zksc_ml_* operations and the KZG entry points (first device); everything else returns ZKSC_ERR_UNSUPPORTED.  n_devices == 1 gives an ordinary context. */
int zksc_ctx_create_multi(const int* devices, int n_devices, zksc_ctx** out);
/* Number of devices behind a context (1 unless it came from zksc_ctx_create_multi). */
int zksc_ctx_devices(const zksc_ctx* ctx);
int zksc_ctx_destroy(zksc_ctx* ctx);
/* Last error text of this context (or of the failed zksc_ctx_create when ctx == NULL). */
const char* zksc_last_error(const zksc_ctx* ctx);
/* Multi-GPU: one process per GPU.  `unique_id` is the 128-byte ncclUniqueId made by rank 0
 * (zksc_comm_unique_id) and distributed by the caller (torch.distributed / MPI / files). */
int zksc_comm_unique_id(uint8_t out_id[128]);
int zksc_comm_init(zksc_ctx* ctx, int n_ranks, int rank, const uint8_t unique_id[128]);
int zksc_ctx_rank(const zksc_ctx* ctx, int* rank, int* n_ranks);
/* 1 when the per-round partial evaluations of a sharded context travel through peer memory (CUDA IPC over
 * NVLink) inside the round kernel, 0 when they go through ncclAllGather + a host sum (single rank: 0). */
int zksc_ctx_peer_exchange(const zksc_ctx* ctx);
/* Sharded contexts: the table size (entries IN TOTAL over all ranks) at which the shards are gathered and the remaining rounds run
 * replicated on every rank, without a per-round exchange (1 on a single-rank context). */
uint64_t zksc_ctx_gather_entries(const zksc_ctx* ctx);
int zksc_ctx_synchronize(zksc_ctx* ctx);
/* Measurement hooks (bench.py; no reference counterpart).  zksc_ctx_stream: the cudaStream_t every kernel
 * of this context is launched on (so callers can record their own events on it).  zksc_ctx_launch_count:
 * kernels launched by this context so far.  zksc_ctx_timing(ctx, 1) brackets every round-kernel launch
 * with CUDA events from then on; zksc_ctx_timing_read drains the records (per launch: duration in ms,
 * product degree, whether the previous challenge's fold was fused in, pairs per table processed, proofs
 * in the launch) and returns how many were written (at most cap). */
void* zksc_ctx_stream(zksc_ctx* ctx);
unsigned long long zksc_ctx_launch_count(const zksc_ctx* ctx);
int zksc_ctx_timing(zksc_ctx* ctx, int enable);
int zksc_ctx_timing_read(zksc_ctx* ctx, uint32_t cap, uint32_t* n_out, float* ms, uint32_t* degree, uint32_t* fold, uint64_t* pairs,
                         uint64_t* proofs);
/* Host wall time, in microseconds, of every round of the latest zksc_prove on this context (device pass + transcript + bind:
 * the rounds add up to the call); at most cap values, *n_out = how many. */
int zksc_ctx_round_times(const zksc_ctx* ctx, uint32_t cap, uint32_t* n_out, double* us);
/* The integer-multiply roof of this device, measured now: 32 x 32 -> 64 bit multiply-adds (IMAD.WIDE.U32, what every limb product of
 * the field arithmetic compiles to) per second with all SMs full and nothing else in the loop (about 1 ms of GPU time). */
int zksc_int_peak(zksc_ctx* ctx, double* limb_products_per_second);

/* ---- device-resident evaluation tables --------------------------------------------------------
 * A `zksc_tables` is `n_proofs` independent instances of  sum_p prod_k f_{p,k}  : for each proof,
 * `n_products` products (Vec<ComposedMultilinear<F>>, multi_composed_sumcheck.rs:47-50), product p
 * having degree[p] factor tables (ComposedMultilinear.polys, composed_multilinear.rs:7-18) of 2^n_vars
 * entries each (Multilinear.evaluations, evaluation_form.rs:5-9).  Variable 0 is the most significant
 * index bit.  Tables are ordered proof-major, then product, then factor.
 *
 * Sharding (contexts with a communicator of G ranks): every rank holds the entries i with
 * i mod G == rank of every table, as a table of 2^n_vars / G entries ("local" tables); n_vars is always
 * the GLOBAL number of variables.                                                                 */

/* Copy host tables to the device.  host_tables[t] -> 2^n_vars elements (full table, also when
 * sharded: the rank picks out its own entries). */
int zksc_tables_upload(zksc_ctx* ctx, uint32_t n_vars, uint32_t n_proofs, uint32_t n_products, const uint32_t* degree,
                       const uint64_t* const* host_tables, zksc_tables** out);
/* Sharded contexts: as zksc_tables_upload, but host_local_tables[t] -> this rank's 2^n_vars / n_ranks
 * entries (those with index = rank mod n_ranks, in order), so no rank touches data it does not own. */
int zksc_tables_upload_local(zksc_ctx* ctx, uint32_t n_vars, uint32_t n_proofs, uint32_t n_products, const uint32_t* degree,
                             const uint64_t* const* host_local_tables, zksc_tables** out);
/* Copy new host tables into an existing handle of the same shape (and reset it): the steady-state form of
 * zksc_tables_upload (local == 0) / zksc_tables_upload_local (local != 0) for callers that prove many
 * instances of one shape. */
int zksc_tables_reupload(zksc_tables* t, const uint64_t* const* host_tables, int local);
/* The same refill, split in two so that it overlaps other work of the context (double buffering across proofs: prove handle
 * A on the compute stream while handle B is refilled on the context's copy stream).  _begin queues the copies and returns;
 * the host buffers (pinned memory for a truly asynchronous copy) stay borrowed until _end returns.  Any other call on the
 * handle finishes a pending refill first.  Sharded contexts with full (non-local) tables copy synchronously in _begin. */
int zksc_tables_reupload_begin(zksc_tables* t, const uint64_t* const* host_tables, int local);
int zksc_tables_reupload_end(zksc_tables* t);
/* This rank's copy of the tables as uploaded (n_proofs x n_tables x 2^n_vars / n_ranks elements). */
int zksc_tables_read_local(zksc_tables* t, uint64_t* out);
/* Fill tables on the device with the seeded synthetic generator (entry = f(seed + proof, table, i);
 * DESIGN.md "Synthetic inputs"); each rank generates only its shard. */
int zksc_tables_synth(zksc_ctx* ctx, uint32_t n_vars, uint32_t n_proofs, uint32_t n_products, const uint32_t* degree, uint64_t seed,
                      zksc_tables** out);
/* Tables built on the device (GKR layer polynomials, gkr/src/protocol.rs:67-93 and gkr/src/utils.rs:23-33):
 * zksc_tables_alloc gives uninitialised tables of a shape; every table must then be filled by one of
 *   zksc_tables_fill_outer  : table = a (+) b or a (x) b, the outer sum / product of two host vectors --
 *                             Multilinear::add_distinct / mul_distinct (evaluation_form.rs:28-52); na * nb == 2^n_vars
 *   zksc_tables_fill_sparse : table = 0 except table[idx[i]] = vals[i] (indices distinct) -- a wiring table
 *                             (Circuit::add_mult_mle, circuit/src/circuit.rs:57-95) after partial_evaluations over
 *                             its gate-label variables, scaled and summed (protocol.rs:70-74, 86-88)
 *   zksc_tables_fill_dense  : table = the given 2^n_vars host elements
 * `table` counts proof-major, then product, then factor.  On a sharded context every rank passes the same
 * (global) arguments and keeps its own entries. */
int zksc_tables_alloc(zksc_ctx* ctx, uint32_t n_vars, uint32_t n_proofs, uint32_t n_products, const uint32_t* degree, zksc_tables** out);
int zksc_tables_fill_outer(zksc_tables* t, uint32_t table, int mul, const uint64_t* a, uint64_t na, const uint64_t* b, uint64_t nb);
int zksc_tables_fill_sparse(zksc_tables* t, uint32_t table, const uint64_t* idx, const uint64_t* vals, uint64_t count);
int zksc_tables_fill_dense(zksc_tables* t, uint32_t table, const uint64_t* evals);
int zksc_tables_free(zksc_tables* t);
/* Forget all bound challenges: back to the tables as uploaded (the input is never modified). */
int zksc_tables_reset(zksc_tables* t);
/* Number of variables still unbound (global). */
int zksc_tables_vars_left(const zksc_tables* t, uint32_t* out);

/* The round-evaluation idiom  p.partial_evaluation(F::from(i),0).element_wise_product().sum(), i = 0..=d
 * (multi_composed_sumcheck.rs:81-89, composed_sumcheck.rs:41-49; for degree 1 this is
 * split_poly_into_two_and_sum_each_part, evaluation_form.rs:68-74), for every proof and product, with
 * any pending zksc_bind fused into the same pass.
 * out: n_proofs x (sum_p degree[p]+1) elements; proof-major, product-major, i-minor.  All ranks of a
 * sharded context receive the full (reduced) values. */
int zksc_round_evals(zksc_tables* t, uint64_t* out);
/* Bind variable 0 of every table to a challenge: Multilinear::partial_evaluation(r, 0)
 * (evaluation_form.rs:123-141, as used at multi_composed_sumcheck.rs:103-105).  challenges: n_proofs
 * elements.  The fold is deferred and fused into the next zksc_round_evals / zksc_residual. */
int zksc_bind(zksc_tables* t, const uint64_t* challenges);
/* Current contents of the tables (after all binds): n_proofs x n_tables x 2^vars_left elements.
 * On a sharded context the shards are gathered (ncclAllGather) and every rank gets the full tables. */
int zksc_residual(zksc_tables* t, uint64_t* out);
/* sum over the hypercube of sum_p prod_k f_{p,k}: MultiComposedSumcheckProver::calculate_poly_sum
 * (multi_composed_sumcheck.rs:37-45) / Sumcheck::poly_sum (sumcheck.rs:25-27).  out: n_proofs elements.
 * Must be called on unbound tables. */
int zksc_poly_sum(zksc_tables* t, uint64_t* out);
/* Multilinear::to_bytes / composed_poly_to_bytes (evaluation_form.rs:54-62, sumcheck/src/utils.rs:53-59)
 * of proof `proof`: n_tables x 2^n_vars x 32 canonical big-endian bytes.  Single-rank contexts only. */
int zksc_tables_to_bytes(zksc_tables* t, uint32_t proof, uint8_t* out);

/* ---- the provers (host round loop + SHA-256 transcript inside the library) ---------------------
 * Protocol selector: which reference prover's transcript/message format to follow. */
#define ZKSC_PROTO_SUMCHECK 0       /* Sumcheck::prove            sumcheck/src/sumcheck.rs:29-61  (1 product, degree 1) */
#define ZKSC_PROTO_COMPOSED 1       /* ComposedSumcheck::prove    composed/composed_sumcheck.rs:32-67 (1 product)       */
#define ZKSC_PROTO_MULTI_PARTIAL 2  /* MultiComposedSumcheckProver::prove_partial  multi_composed_sumcheck.rs:56-62     */
#define ZKSC_PROTO_MULTI_FULL 3     /* MultiComposedSumcheckProver::prove (absorbs every table first)  :47-54          */

/* Prove all n_proofs instances (independent transcripts).  `sums`: n_proofs claimed sums (ignored for
 * ZKSC_PROTO_COMPOSED, which absorbs no sum).  Outputs, per proof and round:
 *   round_msgs   : n_proofs x n_vars x msg_stride elements -- SUMCHECK: [h0,h1]; COMPOSED: evals at 0..d;
 *                  MULTI_*: monomials as (coeff, pow) pairs, `round_len` of them (sparse polynomial,
 *                  sparse_univariate.rs:11-20).  msg_stride = zksc_msg_stride(...) elements.
 *   round_len    : n_proofs x n_vars -- number of elements (SUMCHECK/COMPOSED) or monomials (MULTI_*)
 *   challenges   : n_proofs x n_vars elements
 * The tables end fully bound except for the last challenge; call zksc_tables_reset to reuse them. */
int zksc_prove(zksc_tables* t, int protocol, const uint64_t* sums, uint64_t* round_msgs, uint32_t* round_len, uint64_t* challenges);
uint32_t zksc_msg_stride(int protocol, uint32_t n_products, const uint32_t* degree);
/* Serialise one proof's round messages exactly as the reference feeds them to its transcripts:
 * ComposedSumcheckProof::to_bytes (multi_composed_sumcheck.rs:24-32) for MULTI_*, vec_to_bytes per
 * round otherwise.  Returns the byte count in *out_len (out may be NULL to query). */
int zksc_proof_to_bytes(int protocol, uint32_t n_vars, uint32_t msg_stride, const uint64_t* round_msgs, const uint32_t* round_len,
                        uint8_t* out, size_t* out_len);

/* MultiComposedSumcheckVerifier::verify_partial (multi_composed_sumcheck.rs:143-181) and the transcript
 * half of Sumcheck::verify / ComposedSumcheck::verify: replay the transcript, check p(0)+p(1) against
 * the running claim, return the sub-claim.  Host only (no device work).  ZKSC_ERR_VERIFY on mismatch. */
int zksc_verify_rounds(int protocol, uint32_t n_vars, uint32_t msg_stride, const uint64_t* sum, const uint64_t* round_msgs,
                       const uint32_t* round_len, const uint8_t* absorbed_prefix, size_t prefix_len, uint64_t* subclaim_sum,
                       uint64_t* challenges);
/* The oracle check of MultiComposedSumcheckVerifier::verify (:135-141) / Sumcheck::verify (:94) /
 * ComposedSumcheck::verify (:94): sum_p prod_k f_{p,k}(points), by n successive folds on the device.
 * points: n_proofs x n_vars elements; out: n_proofs elements.  Resets the tables first and after. */
int zksc_evaluate(zksc_tables* t, const uint64_t* points, uint64_t* out);

/* ---- GKR layer driver -------------------------------------------------------
 * GKRProtocol::prove (gkr/src/protocol.rs:21-113, layer one: gkr/src/utils.rs:12-57) in one call: per layer the four
 * tables [alpha add(r_b,.,.) + beta add(r_c,.,.), W(b)+W(c), alpha mul(..) + beta mul(..), W(b) W(c)] are built in HBM
 * (outer sum / product kernels; the wiring tables as a sparse scatter of sum_g eq(r, g) at (in0(g), in1(g))), the layer
 * sumcheck is MultiComposedSumcheckProver::prove_partial on the device, W(b*), W(c*) are evaluated on the device, and the
 * outer Fiat-Shamir transcript (w_0 bytes, every proof's to_bytes, n_r, alpha, beta) runs on the host.
 * Circuit (circuit/src/circuit.rs:10-25): n_layers layers, output layer first; layer i has n_gates[i] = 2^i gates whose
 * inputs are labels of i + 1 bits into layer i + 1 (circuit/src/utils.rs:12-34); gates of all layers concatenated in
 * gate_type (0 = Add, 1 = Mul, gate.rs:2-5), gate_in0, gate_in1.
 * layer_values: n_layers + 1 host vectors = Circuit::evaluation's result (circuit.rs:32-55), [0] = output ... [n_layers] =
 * input, value_len[i] = 2^i Montgomery elements.
 * Outputs (Montgomery): w0[2] = [out, 0] (protocol.rs:31-34); per layer i: sums[i] (the claimed sum of its sumcheck),
 * wb_s[i], wc_s[i] (protocol.rs:103-105); the rounds of all layers concatenated -- layer i has 2 (i + 1) rounds,
 * zksc_gkr_total_rounds(n_layers) = n_layers (n_layers + 1) in total -- as round_msgs (6 elements per round: up to three
 * (coeff, pow) monomials), round_len (monomials per round) and challenges (one element per round), the per-layer slices
 * being exactly what zksc_prove(ZKSC_PROTO_MULTI_PARTIAL) writes and zksc_proof_to_bytes / zksc_verify_rounds read.
 * Shape violations the reference's constructors panic on return ZKSC_ERR_SHAPE.  Unsharded contexts only. */
uint64_t zksc_gkr_total_rounds(uint32_t n_layers);
int zksc_gkr_prove(zksc_ctx* ctx, uint32_t n_layers, const uint32_t* n_gates, const uint8_t* gate_type, const uint32_t* gate_in0,
                   const uint32_t* gate_in1, const uint64_t* const* layer_values, const uint64_t* value_len, uint64_t* w0, uint64_t* sums,
                   uint64_t* wb_s, uint64_t* wc_s, uint64_t* round_msgs, uint32_t* round_len, uint64_t* challenges);

/* ---- GKR on layered circuits of any widths, linear time (second reading of BASELINE config 4) ---------
 * The same protocol (gkr/src/protocol.rs:21-113) and the same proof bytes as zksc_gkr_prove, without the dense 2^(2k)-entry layer
 * tables: the k rounds of a layer sumcheck that bind b run on [W, H1 | H2, 1], the k rounds that bind c on
 * [A, W(u) + W | W(u) M, W] -- tables of 2^k entries built from the gate lists (csrc/gkr_linear.cuh has the algebra) -- on one
 * transcript.  The reference's Circuit can only hold the pyramid shape (layer i: 2^i gates, circuit/src/utils.rs:1-34); here layer
 * i has 2^log_width[i] gates (log_width[0] may be 0: one output, padded to [out, 0] as protocol.rs:31-34 does) whose inputs are wire
 * indices of layer i + 1; log_width[n_layers] is the input layer.  Gates of all layers concatenated, output layer first, in
 * gate_type (0 = Add, 1 = Mul), gate_in0, gate_in1.
 * zksc_circuit_create groups the gates by either input once (host) and keeps the circuit in HBM.
 * zksc_circuit_evaluate = Circuit::evaluation (circuit/src/circuit.rs:32-55) on the device: inputs 2^log_width[n_layers] Montgomery
 *   elements (host), outputs (optional) 2^log_width[0]; the layer values stay in HBM.  zksc_circuit_layer_values reads one layer back.
 * zksc_gkr_prove_linear proves the latest evaluation.  Outputs as zksc_gkr_prove's: w0 (max(2, 2^log_width[0]) elements), per layer
 *   sums, wb_s, wc_s, and the rounds of all layers concatenated -- layer i has 2 log_width[i + 1] rounds,
 *   zksc_circuit_total_rounds in all -- as round_msgs (6 elements per round), round_len, challenges.
 * Unsharded single-device contexts only. */
typedef struct zksc_circuit zksc_circuit;
int zksc_circuit_create(zksc_ctx* ctx, uint32_t n_layers, const uint32_t* log_width, const uint8_t* gate_type, const uint32_t* gate_in0,
                        const uint32_t* gate_in1, zksc_circuit** out);
int zksc_circuit_free(zksc_circuit* c);
int zksc_circuit_evaluate(zksc_circuit* c, const uint64_t* inputs, uint64_t* outputs);
int zksc_circuit_layer_values(zksc_circuit* c, uint32_t layer, uint64_t* out);
uint64_t zksc_circuit_total_rounds(const zksc_circuit* c);
/* Verifier side of one layer (GKRProtocol::verify, gkr/src/protocol.rs:131-133, :164-171): out[0] = alpha add(r_b, b, c) + beta add(r_c, b, c),
 * out[1] = the same for mul, from the gate lists on the device (the reference evaluates dense wiring tables).  r_b, r_c: max(1, log_width[layer])
 * coordinates each (r_c / beta NULL: one point, the output layer's n_r); b, c: log_width[layer + 1] coordinates each. */
int zksc_circuit_wiring_eval(zksc_circuit* c, uint32_t layer, const uint64_t* r_b, const uint64_t* alpha, const uint64_t* r_c, const uint64_t* beta,
                             const uint64_t* b, const uint64_t* cpt, uint64_t* out);
int zksc_gkr_prove_linear(zksc_circuit* c, uint64_t* w0, uint64_t* sums, uint64_t* wb_s, uint64_t* wc_s, uint64_t* round_msgs, uint32_t* round_len,
                          uint64_t* challenges);

/* ---- stand-alone Multilinear operations on caller-owned host vectors (device compute) -----------
 * Each copies its inputs to the device, runs the CUDA kernel and copies the result back. */
/* Multilinear::partial_evaluation(r, variable_index)  evaluation_form.rs:123-141; n = 2^k entries in, n/2 out */
int zksc_ml_partial_evaluation(zksc_ctx* ctx, const uint64_t* evals, uint64_t n, const uint64_t* r, uint32_t variable_index, uint64_t* out);
/* Multilinear::evaluation(points)  evaluation_form.rs:162-175 */
int zksc_ml_evaluation(zksc_ctx* ctx, const uint64_t* evals, uint64_t n, const uint64_t* points, uint32_t n_points, uint64_t* out);
/* add_distinct / mul_distinct  evaluation_form.rs:28-52 : out[i*nb+j] = a[i] (+|*) b[j] */
int zksc_ml_outer(zksc_ctx* ctx, int mul, const uint64_t* a, uint64_t na, const uint64_t* b, uint64_t nb, uint64_t* out);
/* impl Add / Sub / Mul<F> for Multilinear (evaluation_form.rs:178-251) and element_wise_product of two
 * tables (composed_multilinear.rs:105-111): op 0 add, 1 sub, 2 mul, 3 scale by b[0] */
int zksc_ml_elementwise(zksc_ctx* ctx, int op, const uint64_t* a, const uint64_t* b, uint64_t n, uint64_t* out);

/* ---- multilinear KZG over BLS12-381 G1 (kzg/src/multilinear_kzg.rs) -----------------------------------
 * A G1 point crosses the boundary in ark-ec 0.4's in-memory form of `G1Projective` (short_weierstrass::Projective: Jacobian X, Y, Z,
 * each an Fq of 6 x uint64_t little-endian limbs in Montgomery form, R = 2^384; Z = 0 is the identity): 18 x uint64_t per point, so a
 * `&[G1Projective]` (TrustedSetup::powers_of_tau_in_g1, kzg/src/trusted_setup.rs:10-13) is passed unchanged.  Results come back
 * normalised (Z = 1 in Montgomery form, or the identity as (1, 1, 0)): equal as group elements to what the reference computes, and
 * bit-identical to it after `into_affine()`.
 * zksc_g1_msm: sum_i scalars[i] * points[i] -- MultilinearKZG::commitment (multilinear_kzg.rs:33-48: evaluations zipped with
 *   powers_of_tau_in_g1, one mul_bigint each, summed).  scalars: n Fr elements (Montgomery); points: n x 18; out: 18.
 * zksc_kzg_open: MultilinearKZG::open (multilinear_kzg.rs:50-88): per variable the quotient f(1,.) - f(0,.) (get_poly_quotient,
 *   kzg/src/utils.rs:12-17) blown up to all variables by repetition (add_to_front / duplicate_evaluation,
 *   evaluation_form.rs:86-96,112-119) is committed, and the polynomial is replaced by its remainder partial_evaluation(point, 0)
 *   (get_poly_remainder, utils.rs:5-10).  evals: 2^n_vars elements; points: n_vars elements; srs_g1: 2^n_vars x 18;
 *   out_evaluation: 1 element (= poly.evaluation(points)); out_proofs: n_vars x 18.
 * The pairing check of MultilinearKZG::verify is host-side work on n + 1 single points: zksc_pairing_check below (a Rust caller may as well
 * keep ark-ec's pairing).  Nothing of it runs on the device. */
/* The verifier's pairing equation on the host (csrc/host_pairing.hpp; no device work, no context): is prod_i e(P_i, Q_i) == 1 ?
 * g1: n x 12 u64 = affine (x, y); g2: n x 24 u64 = affine (x.c0, x.c1, y.c0, y.c1); canonical little-endian limbs, all zero = the identity.
 * MultilinearKZG::verify (multilinear_kzg.rs:90-116, sum_pairing_results kzg/src/utils.rs:42-61) is
 *   e(C - v g1, -g2) * prod_i e(proof_i, tau_i g2 - z_i g2) == 1.
 * ZKSC_ERR_SHAPE for a coordinate >= p or a point off its curve; subgroup membership is not checked (ark-ec's unchecked forms).
 * zksc_pairing: e(P, Q) itself as 12 Fq coefficients (72 u64, canonical) in tower order -- for cross-checks. */
int zksc_pairing_check(const uint64_t* g1, const uint64_t* g2, uint32_t n, int* is_one);
/* MultilinearKZG::verify (multilinear_kzg.rs:90-116) in one host call.  commitment, proofs: 18 u64 each (ark-ec G1Projective memory form, as
 * zksc_g1_msm / zksc_kzg_open return them); points (n_vars), evaluation: Montgomery Fr elements; srs_g2: the trusted setup's n_vars G2 powers
 * tau_i g2 (trusted_setup.rs:37-46), 24 u64 each as in zksc_pairing_check.  *ok = 1 (verified) / 0.  No context, no device work. */
int zksc_kzg_verify(const uint64_t* commitment, const uint64_t* points, const uint64_t* evaluation, const uint64_t* proofs, const uint64_t* srs_g2,
                    uint32_t n_vars, int* ok);
int zksc_pairing(const uint64_t* g1, const uint64_t* g2, uint64_t* out);
int zksc_g1_msm(zksc_ctx* ctx, const uint64_t* scalars, const uint64_t* points, uint64_t n, uint64_t* out);
int zksc_kzg_open(zksc_ctx* ctx, const uint64_t* evals, uint32_t n_vars, const uint64_t* points, const uint64_t* srs_g1, uint64_t* out_evaluation,
                  uint64_t* out_proofs);

/* ---- host helpers (no device) -------------------------------------------------------------------- */
/* F::from(u64) in Montgomery form; canonical <-> Montgomery; be32 (sumcheck/src/utils.rs:7-9) */
void zksc_fr_from_u64(uint64_t x, uint64_t out[4]);
void zksc_fr_from_canonical(const uint64_t canonical[4], uint64_t out[4]);
void zksc_fr_to_canonical(const uint64_t mont[4], uint64_t out[4]);
void zksc_fr_from_canonical_batch(const uint64_t* canonical, uint64_t n, uint64_t* out);
void zksc_fr_to_canonical_batch(const uint64_t* mont, uint64_t n, uint64_t* out);
void zksc_fr_to_be_bytes(const uint64_t mont[4], uint8_t out[32]);
void zksc_fr_from_be_bytes_mod_order(const uint8_t in[32], uint64_t out[4]);
void zksc_fr_add(const uint64_t a[4], const uint64_t b[4], uint64_t out[4]);
void zksc_fr_sub(const uint64_t a[4], const uint64_t b[4], uint64_t out[4]);
void zksc_fr_mul(const uint64_t a[4], const uint64_t b[4], uint64_t out[4]);
/* FiatShamirTranscript (transcripts/fiat-shamir/src/fiat_shamir.rs:5-40) */
typedef struct zksc_transcript zksc_transcript;
zksc_transcript* zksc_transcript_new(void);
void zksc_transcript_free(zksc_transcript* t);
void zksc_transcript_commit(zksc_transcript* t, const uint8_t* data, size_t len);
void zksc_transcript_challenge(zksc_transcript* t, uint8_t out[32]);
void zksc_transcript_challenge_field(zksc_transcript* t, uint64_t out[4]);
/* SparseUnivariatePolynomial::interpolation over x = 0..n-1 (sparse_univariate.rs:40-63): returns the
 * number of monomials written to out_mono as (coeff, pow) pairs (zero coefficients dropped). */
uint32_t zksc_sparse_interpolate(const uint64_t* ys, uint32_t n, uint64_t* out_mono);
/* impl Add for SparseUnivariatePolynomial (sparse_univariate.rs:159-203); returns monomial count */
uint32_t zksc_sparse_add(const uint64_t* a_mono, uint32_t na, const uint64_t* b_mono, uint32_t nb, uint64_t* out_mono);
/* SparseUnivariatePolynomial::evaluate (sparse_univariate.rs:90-106) */
void zksc_sparse_evaluate(const uint64_t* mono, uint32_t n, const uint64_t point[4], uint64_t out[4]);
/* What the round kernels deliver for one product -> its evaluations at 0..degree, in place (degree + 1 Montgomery elements).
 * The kernels evaluate degree-2 products at {0, 1, inf} and degree-3 products at {0, 1, -1, inf} (cheaper operands than 2 and 3;
 * inf = the leading coefficient); zksc_round_evals / zksc_prove apply this conversion themselves -- it is exported so that the
 * identity can be tested without a device.  Other degrees are left as they are. */
void zksc_round_slots_to_evals(uint32_t degree, uint64_t* values);
/* Seeded synthetic table entry (canonical value -> Montgomery), same convention as the device generator */
void zksc_synth_entry(uint64_t seed, uint64_t table, uint64_t index, uint64_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* ZKSC_H */
