//! The reference's own protocol tests, restated against the GPU shim (source only: see rust/README.md; run with rust/check.sh on
//! a machine that has cargo, nvcc and a B200).  Inputs and assertions are those of
//!   sumcheck/src/sumcheck.rs:108-202, sumcheck/src/composed/composed_sumcheck.rs:108-241,
//!   sumcheck/src/composed/multi_composed_sumcheck.rs:195-311
//! with `Fr` = ark_test_curves::bls12_381::Fr (the reference's tests wrap it in field-tracker's op counter `Ft<4, Fr>`, a
//! test-only instrument).  tests/cpp/reference_cases.cpp holds the same cases against the C++ mirror and DOES run in this
//! repository's GPU test suite.
use ark_test_curves::bls12_381::Fr;
use polynomial::{ComposedMultilinear, Multilinear, MultilinearTrait};
use zksc_sumcheck::composed::ComposedSumcheck;
use zksc_sumcheck::{Gpu, MultiComposedSumcheckProver, MultiComposedSumcheckVerifier, Sumcheck};

fn ml(v: &[u64]) -> Multilinear<Fr> {
    Multilinear::new(v.iter().map(|x| Fr::from(*x)).collect())
}

#[test]
fn test_sum_calculation() {
    let mut prover = Sumcheck::new(ml(&[0, 0, 0, 2, 2, 2, 2, 4]));
    prover.poly_sum();
    assert_eq!(prover.sum, Fr::from(12));
}

#[test]
fn test_sum_check_proof() {
    for evals in [&[0u64, 0, 2, 7, 3, 3, 6, 11][..], &[0, 0, 0, 0, 0, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0][..], &[1, 3, 5, 7, 2, 4, 6, 8, 3, 5, 7, 9, 4, 6, 8, 10][..]] {
        let mut sumcheck = Sumcheck::new(ml(evals));
        sumcheck.poly_sum();
        let (proof, _challenges) = sumcheck.prove();
        assert!(sumcheck.verify(&proof));
    }
}

#[test]
fn test_composed_sumcheck_sum_and_proof() {
    let (a, b) = (ml(&[0, 0, 0, 2]), ml(&[0, 3, 0, 3]));
    let poly = ComposedMultilinear::new(vec![a, b]);
    assert_eq!(ComposedSumcheck::calculate_poly_sum(&poly), Fr::from(6));
    let sumcheck = ComposedSumcheck::new(poly.clone());
    let (proof, _challenges) = sumcheck.prove();
    assert!(sumcheck.verify(&proof, Fr::from(6)));
    assert!(!sumcheck.verify(&proof, Fr::from(7)));
}

#[test]
fn test_multi_composed_sum_check_proof_2_on_gkr_example() {
    // multi_composed_sumcheck.rs:267-311
    let add_i = ml(&[0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]);
    let mul_i = ml(&[0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0]);
    let w = ml(&[3, 2, 3, 1]);
    let r = Fr::from(2u64);
    let add_rbc = Gpu(add_i).partial_evaluation(&r, &0).0;
    let mul_rbc = Gpu(mul_i).partial_evaluation(&r, &0).0;
    let wb_add_wc = w.add_distinct(&w);
    let wb_mul_wc = w.mul_distinct(&w);
    let polys = vec![ComposedMultilinear::new(vec![add_rbc, wb_add_wc]), ComposedMultilinear::new(vec![mul_rbc, wb_mul_wc])];
    let sum = MultiComposedSumcheckProver::calculate_poly_sum(&polys);
    let (proof, challenges) = MultiComposedSumcheckProver::prove(&polys, &sum).unwrap();
    assert!(MultiComposedSumcheckVerifier::verify(&polys, &proof).unwrap());
    let (partial, _) = MultiComposedSumcheckProver::prove_partial(&polys, &sum).unwrap();
    let sub = MultiComposedSumcheckVerifier::verify_partial(&partial).unwrap();
    assert_eq!(sub.challenges.len(), challenges.len());
    // a wrong claimed sum is rejected with the reference's error
    let mut bad = partial;
    bad.sum += Fr::from(1u64);
    assert_eq!(MultiComposedSumcheckVerifier::verify_partial(&bad).unwrap_err(), "Verification failed");
}

/// gkr/src/protocol.rs:209-233 (test_gkr_protocol_1): the linear-time prover on the reference's circuit gives the dense prover's proof
#[test]
fn test_gkr_linear_prover_matches_the_dense_prover_on_the_reference_circuit() {
    use zksc_sumcheck::{gkr_prove, LayeredCircuit};
    let gates = vec![vec![(true, 0usize, 1usize)], vec![(false, 0, 1), (true, 2, 3)]];
    let input: Vec<Fr> = [2u64, 3, 4, 5].iter().map(|&x| Fr::from(x)).collect();
    let mut lc = LayeredCircuit::new(&[0, 1, 2], &gates);
    assert_eq!(lc.evaluate(&input), vec![Fr::from(100u64)]);
    let evaluation = vec![vec![Fr::from(100u64)], vec![Fr::from(5u64), Fr::from(20u64)], input.clone()];
    let (proofs, wb, wc, w0) = lc.prove();
    let (dproofs, dwb, dwc, dw0) = gkr_prove(&gates, &evaluation);
    assert_eq!(proofs.len(), dproofs.len());
    for (a, b) in proofs.iter().zip(dproofs.iter()) {
        assert_eq!(a.to_bytes(), b.to_bytes());
    }
    assert_eq!((wb, wc, w0), (dwb, dwc, dw0));
}
