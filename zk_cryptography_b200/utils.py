"""Mirror of the reference's `sumcheck/src/utils.rs` helpers over the device API.

  convert_field_to_byte                  :7-9     element.into_bigint().to_bytes_be()
  skip_first_and_sum_all                 :11-27   test-only, exponential: every vertex of {0,1}^(n-1) bound on variable 1
  convert_round_poly_to_uni_poly_format  :29-35   [(F::from(i), y_i)]
  vec_to_bytes                           :37-43
  sum_over_boolean_hypercube             :45-51   sum_x sum_p prod_k f_{p,k}(x)   (device: zksc_poly_sum)
  composed_poly_to_bytes                 :53-59
  boolean_hypercube                      polynomial/src/utils.rs:141-157
Scalars are Python ints (canonical residues), tables are `Multilinear` / `ComposedMultilinear` of api.py."""
from . import _lib
from .api import Multilinear, MultiComposedSumcheckProver

R = _lib.R_MOD


def convert_field_to_byte(element):
    return (int(element) % R).to_bytes(32, "big")


def boolean_hypercube(n):
    """vertices in counting order, most significant bit first (polynomial/src/utils.rs:141-157)"""
    return [[(i >> j) & 1 for j in range(n - 1, -1, -1)] for i in range(1 << n)]


def skip_first_and_sum_all(current_poly):
    """sum over every assignment of all variables but the first: a 1-variable table [sum f(0, x), sum f(1, x)].
    Literal restatement (2^(n-1) chains of partial evaluations on variable index 1, on the device): for tests on small tables;
    `Multilinear.split_poly_into_two_and_sum_each_part` gives the same two values in one pass."""
    rounds = current_poly.n_vars - 1
    bh_sum = Multilinear([0, 0])                      # Multilinear::additive_identity(1)
    for vertex in boolean_hypercube(rounds):
        sum_except_first = current_poly
        for bit in vertex:
            sum_except_first = sum_except_first.partial_evaluation(bit, 1)
        bh_sum = bh_sum + sum_except_first
    return bh_sum


def convert_round_poly_to_uni_poly_format(round_poly):
    return [(i, int(v) % R) for i, v in enumerate(round_poly)]


def vec_to_bytes(poly):
    return b"".join(convert_field_to_byte(p) for p in poly)


def sum_over_boolean_hypercube(poly):
    return MultiComposedSumcheckProver.calculate_poly_sum(list(poly))


def composed_poly_to_bytes(poly):
    return b"".join(p.to_bytes() for p in poly)
