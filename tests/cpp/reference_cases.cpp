// The reference's own #[test] cases for the sumcheck path, restated against include/zksc.hpp (the C++ mirror of the Rust
// API over the C ABI).  Each case asserts what the Rust test asserts (sums, verify == true) and prints one line
//     <name> <hex of the bytes the transcript absorbed per round>
// which tests/test_cpp_host.py compares with the oracle on the same inputs.  Exit code 0 = every assertion held.
//
//   sumcheck/src/sumcheck.rs:107-202                      test_sum_calculation, test_sum_check_proof{,_2,_3}
//   sumcheck/src/composed/composed_sumcheck.rs:108-241    test_sum_calculation, test_sum_check_proof{,1,_2,_3}
//   sumcheck/src/composed/multi_composed_sumcheck.rs:194-311
//   polynomial/src/multilinear/evaluation_form.rs:264-462 (primitive known answers)
//   gkr/src/protocol.rs:209-285                           test_gkr_protocol_1 / _2, and Circuit::random as in gkr/benches
#include <cstdio>
#include <string>

#include "../../include/zksc.hpp"

using namespace zk;

static int g_failed = 0;
#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            g_failed++;                                                      \
        }                                                                    \
    } while (0)

static std::vector<Fr> frs(std::initializer_list<uint64_t> v) {
    std::vector<Fr> out;
    for (uint64_t x : v) out.push_back(Fr::from(x));
    return out;
}
static Multilinear ml(std::initializer_list<uint64_t> v) { return Multilinear::new_(frs(v)); }
static void emit(const char* name, const std::vector<uint8_t>& bytes) {
    std::printf("%s ", name);
    for (uint8_t b : bytes) std::printf("%02x", b);
    std::printf("\n");
}
static std::vector<uint8_t> bytes_of(const SumcheckProof& p) {   // what Sumcheck::prove absorbs per round (sumcheck.rs:44-47)
    std::vector<uint8_t> out;
    for (const auto& u : p.univariate_poly) { append(out, u.evaluations[0]); append(out, u.evaluations[1]); }
    return out;
}
static std::vector<uint8_t> bytes_of(const composed::ComposedSumcheckProof& p) {   // vec_to_bytes per round (composed_sumcheck.rs:51)
    std::vector<uint8_t> out;
    for (const auto& r : p.round_polys)
        for (const auto& y : r) append(out, y);
    return out;
}
static std::vector<uint8_t> with_challenges(std::vector<uint8_t> b, const std::vector<Fr>& ch) {
    for (const auto& c : ch) append(b, c);
    return b;
}

// ---- polynomial/src/multilinear/evaluation_form.rs ---------------------------------------------------------------
static void test_multilinear_primitives() {
    // test_add_mul_distinct (:264-312)
    CHECK(ml({0, 0, 2, 2}).add_distinct(ml({0, 3, 0, 3})) == ml({0, 3, 0, 3, 0, 3, 0, 3, 2, 5, 2, 5, 2, 5, 2, 5}));
    CHECK(ml({0, 0, 2, 2}).mul_distinct(ml({0, 3, 0, 3})) == ml({0, 0, 0, 0, 0, 0, 0, 0, 0, 6, 0, 6, 0, 6, 0, 6}));
    {   // test_partial_evaluation_1 (:315-325): [3, 1, 2, 5] at 5 on variable 0 -> [-2, 21]
        Multilinear got = ml({3, 1, 2, 5}).partial_evaluation(Fr::from(5), 0);
        CHECK(got == Multilinear::new_({Fr::zero() - Fr::from(2), Fr::from(21)}));
    }
    // test_evaluation_1 / _2 (:362-405)
    CHECK(ml({3, 1, 2, 5}).evaluation(frs({5, 6})) == Fr::from(136));
    CHECK(ml({3, 9, 7, 13, 6, 12, 10, 18}).evaluation(frs({2, 3, 1})) == Fr::from(39));
    CHECK(ml({0, 0, 0, 3, 0, 0, 2, 5}).evaluation(frs({2, 3, 4})) == Fr::from(48));
    // test_split_poly_into_two_and_sum_each_part (:408-438), test_sum_over_boolean_hypercube (:441-462)
    CHECK(ml({0, 0, 0, 2, 2, 2, 2, 4}).split_poly_into_two_and_sum_each_part() == ml({2, 10}));
    CHECK(ml({0, 0, 2, 7, 3, 3, 6, 11}).split_poly_into_two_and_sum_each_part() == ml({9, 23}));
    CHECK(ml({1, 2, 3, 4, 5, 6, 7, 8}).sum_over_the_boolean_hypercube() == Fr::from(36));
    // element_wise_product (composed_multilinear.rs:159-170)
    CHECK(ComposedMultilinear::new_({ml({0, 1, 2, 3}), ml({0, 0, 0, 1})}).element_wise_product() == frs({0, 0, 0, 3}));
    // Multilinear::new panics unless the length is a power of two (:16-20); ComposedMultilinear::new on unequal arity (:15)
    bool threw = false;
    try { Multilinear::new_(frs({1, 2, 3})); } catch (const Error& e) { threw = e.code == ZKSC_ERR_SHAPE; }
    CHECK(threw);
    threw = false;
    try { ComposedMultilinear::new_({ml({1, 2}), ml({1, 2, 3, 4})}); } catch (const Error& e) { threw = e.code == ZKSC_ERR_SHAPE; }
    CHECK(threw);
    // convert_field_to_byte of 1 and 100 (sumcheck/src/utils.rs tests): 32 big-endian bytes
    CHECK(Fr::from(1).to_bytes_be()[31] == 1 && Fr::from(100).to_bytes_be()[31] == 100 && Fr::from(100).to_bytes_be()[0] == 0);
}

// ---- sumcheck/src/sumcheck.rs ----------------------------------------------------------------------------------------
static void sumcheck_case(const char* name, std::initializer_list<uint64_t> evals) {
    Sumcheck sumcheck = Sumcheck::new_(ml(evals));
    sumcheck.poly_sum();
    auto [proof, challenges] = sumcheck.prove();
    bool verifer = sumcheck.verify(proof);
    CHECK(verifer == true);
    emit(name, with_challenges(bytes_of(proof), challenges));
    SumcheckProof bad = proof;
    bad.sum = bad.sum + Fr::one();
    CHECK(sumcheck.verify(bad) == false);
}
static void test_sumcheck() {
    {   // test_sum_calculation (:107-123)
        Sumcheck prover = Sumcheck::new_(ml({0, 0, 0, 2, 2, 2, 2, 4}));
        prover.poly_sum();
        CHECK(prover.sum == Fr::from(12));
    }
    sumcheck_case("sumcheck_proof", {0, 0, 2, 7, 3, 3, 6, 11});                                       // :126-144
    sumcheck_case("sumcheck_proof_2", {0, 0, 0, 0, 0, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0});              // :147-173
    sumcheck_case("sumcheck_proof_3", {1, 3, 5, 7, 2, 4, 6, 8, 3, 5, 7, 9, 4, 6, 8, 10});             // :176-202
}

// ---- sumcheck/src/composed/composed_sumcheck.rs ----------------------------------------------------------------------
static void composed_case(const char* name, std::vector<Multilinear> factors) {
    ComposedMultilinear composed = ComposedMultilinear::new_(std::move(factors));
    ComposedSumcheck sumcheck = ComposedSumcheck::new_(composed);
    Fr sum = ComposedSumcheck::calculate_poly_sum(composed);
    auto [proof, challenges] = sumcheck.prove();
    bool verifer = sumcheck.verify(proof, sum);
    CHECK(verifer == true);
    emit(name, with_challenges(bytes_of(proof), challenges));
    CHECK(sumcheck.verify(proof, sum + Fr::one()) == false);
}
static void test_composed_sumcheck() {
    // test_sum_calculation (:108-140)
    CHECK(ComposedSumcheck::calculate_poly_sum(ComposedMultilinear::new_({ml({0, 1, 2, 3}), ml({0, 0, 0, 1})})) == Fr::from(3));
    CHECK(ComposedSumcheck::calculate_poly_sum(ComposedMultilinear::new_({ml({3, 3, 5, 5}), ml({0, 0, 0, 1})})) == Fr::from(5));
    CHECK(ComposedSumcheck::calculate_poly_sum(ComposedMultilinear::new_({ml({0, 1, 2, 3})})) == Fr::from(6));
    CHECK(ComposedSumcheck::calculate_poly_sum(ComposedMultilinear::new_({ml({0, 0, 0, 2, 2, 2, 2, 4})})) == Fr::from(12));
    composed_case("composed_proof", {ml({3, 3, 5, 5}), ml({0, 0, 0, 1})});                                            // :143-164
    composed_case("composed_proof1", {ml({0, 0, 2, 7, 3, 3, 6, 11})});                                                // :167-185
    composed_case("composed_proof_2", {ml({0, 0, 0, 0, 0, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0})});                        // :188-215
    composed_case("composed_proof_3", {ml({1, 3, 5, 7, 2, 4, 6, 8, 3, 5, 7, 9, 4, 6, 8, 10})});                       // :218-241
}

// ---- sumcheck/src/composed/multi_composed_sumcheck.rs ----------------------------------------------------------------
static void multi_case(const char* name, const std::vector<ComposedMultilinear>& multi_composed) {
    Fr sum = MultiComposedSumcheckProver::calculate_poly_sum(multi_composed);
    auto [proof, challenges] = MultiComposedSumcheckProver::prove(multi_composed, sum).unwrap();
    bool verify = MultiComposedSumcheckVerifier::verify(multi_composed, proof).unwrap();
    CHECK(verify);
    emit(name, with_challenges(proof.to_bytes(), challenges));
    // prove_partial (the entry point gkr calls, :56-62) and its verifier
    auto [pproof, pchal] = MultiComposedSumcheckProver::prove_partial(multi_composed, sum).unwrap();
    SubClaim sub = MultiComposedSumcheckVerifier::verify_partial(pproof).unwrap();
    CHECK(sub.challenges == pchal);
    Fr at = Fr::zero();
    for (const auto& p : multi_composed) at += p.evaluation(pchal);
    CHECK(at == sub.sum);
    emit((std::string(name) + "_partial").c_str(), with_challenges(pproof.to_bytes(), pchal));
    // a wrong claimed sum must come back as Err("Verification failed") (:170)
    ComposedSumcheckProof bad = pproof;
    bad.sum = bad.sum + Fr::one();
    auto r = MultiComposedSumcheckVerifier::verify_partial(bad);
    CHECK(!r.is_ok() && std::string(r.err) == "Verification failed");
}
static void test_multi_composed_sumcheck() {
    // test_sum_calculation (:194-214)
    CHECK(MultiComposedSumcheckProver::calculate_poly_sum({ComposedMultilinear::new_({ml({0, 1, 2, 3})}), ComposedMultilinear::new_({ml({0, 0, 0, 1})})}) == Fr::from(7));
    CHECK(MultiComposedSumcheckProver::calculate_poly_sum({ComposedMultilinear::new_({ml({0, 0, 0, 2})}), ComposedMultilinear::new_({ml({0, 3, 0, 3})})}) == Fr::from(8));
    Multilinear poly1 = ml({0, 0, 0, 2}), poly2 = ml({0, 3, 0, 3});
    multi_case("multi_proof", {ComposedMultilinear::new_({poly1}), ComposedMultilinear::new_({poly2})});                                          // :217-231
    multi_case("multi_proof_1", {ComposedMultilinear::new_({poly1}), ComposedMultilinear::new_({poly2}), ComposedMultilinear::new_({poly2})});    // :234-248
    multi_case("multi_proof_2", {ComposedMultilinear::new_({poly1, poly2}), ComposedMultilinear::new_({poly2, poly1})});                          // :251-264
    {   // test_multi_composed_sum_check_proof_2_on_gkr_example (:267-311)
        Multilinear add_i = ml({4, 4, 7, 7, 4, 4, 7, 9}), w_b = ml({0, 4}), w_c = ml({0, 3}), mul_i = ml({3, 3, 3, 4, 3, 3, 5, 6});
        ComposedMultilinear lhs_poly = ComposedMultilinear::new_({add_i.partial_evaluation(Fr::from(2), 0), w_b.add_distinct(w_c)});
        ComposedMultilinear rhs_poly = ComposedMultilinear::new_({mul_i.partial_evaluation(Fr::from(2), 0), w_b.mul_distinct(w_c)});
        CHECK(MultiComposedSumcheckProver::calculate_poly_sum({lhs_poly, rhs_poly}) == Fr::from(213));   // SURVEY 8(c) derived check value
        multi_case("multi_proof_gkr_example", {lhs_poly, rhs_poly});
    }
}

// ---- gkr/src/protocol.rs ---------------------------------------------------------------------------------------------
static std::vector<uint8_t> bytes_of(const GKRProof& p) {
    std::vector<uint8_t> out;
    for (const auto& e : p.w_0_mle.evaluations) append(out, e);
    for (size_t i = 0; i < p.sumcheck_proofs.size(); i++) {
        auto b = p.sumcheck_proofs[i].to_bytes();
        out.insert(out.end(), b.begin(), b.end());
        append(out, p.wb_s[i]);
        append(out, p.wc_s[i]);
    }
    return out;
}
static void gkr_case(const char* name, const Circuit& circuit, const std::vector<Fr>& input) {
    auto evaluation = circuit.evaluation(input);
    GKRProof proof = GKRProtocol::prove(circuit, evaluation);
    CHECK(GKRProtocol::verify(circuit, input, proof));
    emit(name, bytes_of(proof));
    GKRProof bad = proof;
    bad.wb_s.back() = bad.wb_s.back() + Fr::one();
    CHECK(!GKRProtocol::verify(circuit, input, bad));
    // the linear-time prover for layered circuits of any widths gives the same proof on the reference's pyramid circuits, and its
    // verifier (wiring polynomials from the gate lists on the device) agrees with GKRProtocol::verify
    LayeredCircuit lc = LayeredCircuit::from_circuit(circuit);
    CHECK(lc.evaluate(input) == evaluation[0]);
    GKRProof lin = lc.prove();
    CHECK(lin.sumcheck_proofs.size() == proof.sumcheck_proofs.size());
    for (size_t i = 0; i < lin.sumcheck_proofs.size() && i < proof.sumcheck_proofs.size(); i++) CHECK(lin.sumcheck_proofs[i].to_bytes() == proof.sumcheck_proofs[i].to_bytes());
    CHECK(lin.wb_s == proof.wb_s && lin.wc_s == proof.wc_s);
    CHECK(lc.verify(input, lin));
    CHECK(!lc.verify(input, bad));
}
static void test_layered_uniform() {
    // a uniform-width circuit (2^5 gates in every layer, 3 layers): the reference's Circuit cannot hold it
    std::vector<CircuitLayer> layers;
    uint64_t x = 88172645463325252ull;
    auto next = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    for (int li = 0; li < 3; li++) {
        CircuitLayer l;
        for (int g = 0; g < 32; g++) l.layer.push_back(Gate{(next() & 1) ? GateType::Mul : GateType::Add, {size_t(next() % 32), size_t(next() % 32)}});
        layers.push_back(l);
    }
    std::vector<Fr> input;
    for (uint64_t i = 0; i < 32; i++) input.push_back(Fr::from(0xD1B54A32D192ED03ull * (i + 1)));
    LayeredCircuit lc({5, 5, 5, 5}, layers);
    std::vector<Fr> host = Circuit::new_(layers).evaluation(input)[0];
    CHECK(lc.evaluate(input) == host);
    GKRProof proof = lc.prove();
    CHECK(proof.sumcheck_proofs.size() == 3 && proof.sumcheck_proofs[0].round_polys.size() == 10);
    CHECK(lc.verify(input, proof));
    proof.wc_s[1] = proof.wc_s[1] + Fr::one();
    CHECK(!lc.verify(input, proof));
}
static void test_gkr() {
    using G = Gate;
    const GateType A = GateType::Add, M = GateType::Mul;
    {   // test_gkr_protocol_1 (:209-232)
        Circuit circuit = Circuit::new_({CircuitLayer::new_({G::new_(M, {0, 1})}), CircuitLayer::new_({G::new_(A, {0, 1}), G::new_(M, {2, 3})})});
        gkr_case("gkr_protocol_1", circuit, frs({2, 3, 4, 5}));
    }
    {   // test_gkr_protocol_2 (:235-285)
        Circuit circuit = Circuit::new_({
            CircuitLayer::new_({G::new_(A, {0, 1})}),
            CircuitLayer::new_({G::new_(M, {0, 1}), G::new_(A, {2, 3})}),
            CircuitLayer::new_({G::new_(A, {0, 1}), G::new_(M, {2, 3}), G::new_(M, {4, 5}), G::new_(M, {6, 7})}),
            CircuitLayer::new_({G::new_(M, {0, 1}), G::new_(M, {2, 3}), G::new_(M, {4, 5}), G::new_(A, {6, 7}), G::new_(M, {8, 9}), G::new_(A, {10, 11}),
                                G::new_(M, {12, 13}), G::new_(M, {14, 15})}),
        });
        gkr_case("gkr_protocol_2", circuit, frs({2, 1, 3, 1, 4, 1, 2, 2, 3, 3, 4, 4, 2, 3, 3, 4}));
    }
    {   // Circuit::random(6) on the bench's kind of input (gkr/benches/gkr_benchmark.rs:11-20)
        std::vector<Fr> input;
        for (uint64_t i = 0; i < 64; i++) input.push_back(Fr::from(0x9E3779B97F4A7C15ull * (i + 1)));
        gkr_case("gkr_random_6", Circuit::random(6), input);
    }
}

int main() {
    try {
        test_multilinear_primitives();
        test_sumcheck();
        test_composed_sumcheck();
        test_multi_composed_sumcheck();
        // (the GKR driver runs on single-device contexts: its layer tables are far below the size where sharding pays)
        if (zksc_ctx_devices(Context::get()) == 1) { test_gkr(); test_layered_uniform(); }
    } catch (const Error& e) {
        std::fprintf(stderr, "zk::Error %d: %s\n", e.code, e.what());
        return 2;
    }
    if (g_failed) { std::fprintf(stderr, "%d assertion(s) failed\n", g_failed); return 1; }
    std::printf("ALL REFERENCE CASES OK\n");
    return 0;
}
