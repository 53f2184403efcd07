"""TEST INFRASTRUCTURE ONLY: ctypes loader for oracle/libzkref.so (the C restatement, zkref.c).

Values cross as canonical little-endian 4 x u64 (numpy (n,4) uint64) or Python ints via helpers."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_u64p = ctypes.POINTER(ctypes.c_uint64)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libzkref.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libzkref.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.zkref_max_threads.restype = ctypes.c_int
    return _LIB


def ints_to_canon(values):
    buf = b"".join(int(v).to_bytes(32, "little") for v in values)
    return np.frombuffer(buf, dtype=np.uint64).reshape(len(values), 4).copy()


def canon_to_ints(arr):
    b = np.ascontiguousarray(arr, dtype=np.uint64).tobytes()
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def _p64(a):
    return a.ctypes.data_as(_u64p)


def set_threads(n):
    lib().zkref_set_threads(int(n))


def max_threads():
    """All host threads this process may use: the CPUs of its affinity mask, NOT omp_get_max_threads() -- torchrun exports
    OMP_NUM_THREADS=1 to every rank, which would silently turn the all-cores CPU baseline into a single-thread one."""
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, n) if lib().zkref_max_threads() >= 1 else 1


def synth_table(seed, table, n_vars):
    out = np.zeros((1 << n_vars, 4), dtype=np.uint64)
    lib().zkref_synth_table(ctypes.c_uint64(seed), ctypes.c_uint64(table), ctypes.c_uint32(n_vars), _p64(out))
    return out


def eval_synth(seed, table, n_vars, pts):
    """Multilinear::evaluation of the seeded synthetic table at pts (python ints), streamed: the table is never materialised."""
    out = np.zeros(4, dtype=np.uint64)
    p = ints_to_canon(pts) if pts else np.zeros((1, 4), dtype=np.uint64)
    lib().zkref_eval_synth(ctypes.c_uint64(seed), ctypes.c_uint64(table), ctypes.c_uint32(n_vars), _p64(p), _p64(out))
    return canon_to_ints(out)[0]


def prove(protocol, n_vars, degrees, tables_canon, sum_int):
    """tables_canon: (sum(deg) * 2^n, 4) canonical uint64 array (products concatenated).
    -> (proof_bytes, challenges as python ints)"""
    deg = np.asarray(degrees, dtype=np.uint32)
    tabs = np.ascontiguousarray(tables_canon, dtype=np.uint64)
    s = ints_to_canon([sum_int])
    dmax = int(deg.max())
    buf = np.zeros(max(1, n_vars) * 64 * (dmax + 1) * 2, dtype=np.uint8)
    ln = ctypes.c_size_t()
    chal = np.zeros((max(1, n_vars), 4), dtype=np.uint64)
    rc = lib().zkref_prove(protocol, ctypes.c_uint32(n_vars), ctypes.c_uint32(len(deg)), deg.ctypes.data_as(_u32p), _p64(tabs), _p64(s),
                           buf.ctypes.data_as(_u8p), ctypes.byref(ln), _p64(chal))
    assert rc == 0
    return buf[:ln.value].tobytes(), canon_to_ints(chal[:n_vars])


def poly_sum(n_vars, degrees, tables_canon):
    deg = np.asarray(degrees, dtype=np.uint32)
    tabs = np.ascontiguousarray(tables_canon, dtype=np.uint64)
    out = np.zeros(4, dtype=np.uint64)
    lib().zkref_poly_sum(ctypes.c_uint32(n_vars), ctypes.c_uint32(len(deg)), deg.ctypes.data_as(_u32p), _p64(tabs), _p64(out))
    return canon_to_ints(out)[0]


def partial_evaluation(evals_ints, r, var):
    a = ints_to_canon(evals_ints)
    out = np.zeros((len(evals_ints) // 2, 4), dtype=np.uint64)
    lib().zkref_partial_evaluation(_p64(a), ctypes.c_uint64(len(evals_ints)), _p64(ints_to_canon([r])), ctypes.c_uint64(var), _p64(out))
    return canon_to_ints(out)


def evaluation(evals_ints, pts):
    a = ints_to_canon(evals_ints)
    out = np.zeros(4, dtype=np.uint64)
    p = ints_to_canon(pts) if pts else np.zeros((1, 4), dtype=np.uint64)
    lib().zkref_evaluation(_p64(a), ctypes.c_uint64(len(evals_ints)), _p64(p), _p64(out))
    return canon_to_ints(out)[0]


def interpolate(ys):
    y = ints_to_canon(ys)
    out = np.zeros((2 * len(ys), 4), dtype=np.uint64)
    n = lib().zkref_interpolate(_p64(y), ctypes.c_uint32(len(ys)), _p64(out))
    v = canon_to_ints(out[:2 * n])
    return [(v[2 * i], v[2 * i + 1]) for i in range(n)]


def be32(x):
    out = np.zeros(32, dtype=np.uint8)
    lib().zkref_be32(_p64(ints_to_canon([x])), out.ctypes.data_as(_u8p))
    return out.tobytes()


def fr_mul(a, b):
    out = np.zeros(4, dtype=np.uint64)
    lib().zkref_fr_mul(_p64(ints_to_canon([a])), _p64(ints_to_canon([b])), _p64(out))
    return canon_to_ints(out)[0]


def transcript_two_challenges(data):
    o1 = np.zeros(32, dtype=np.uint8)
    o2 = np.zeros(32, dtype=np.uint8)
    d = np.frombuffer(bytes(data), dtype=np.uint8).copy() if data else np.zeros(1, dtype=np.uint8)
    lib().zkref_transcript_test(d.ctypes.data_as(_u8p), ctypes.c_size_t(len(data)), o1.ctypes.data_as(_u8p), o2.ctypes.data_as(_u8p))
    return o1.tobytes(), o2.tobytes()


def verify_partial(n_vars, sum_int, round_polys):
    """round_polys: list (per round) of [(coeff, pow)] python ints.  -> (ok, sub_sum, challenges)"""
    lens = np.asarray([len(rp) for rp in round_polys], dtype=np.uint32)
    flat = [v for rp in round_polys for cp in rp for v in cp]
    mono = ints_to_canon(flat) if flat else np.zeros((1, 4), dtype=np.uint64)
    sub = np.zeros(4, dtype=np.uint64)
    chal = np.zeros((max(1, n_vars), 4), dtype=np.uint64)
    rc = lib().zkref_verify_partial(ctypes.c_uint32(n_vars), _p64(ints_to_canon([sum_int])), _p64(mono), lens.ctypes.data_as(_u32p), _p64(sub), _p64(chal))
    return rc == 0, canon_to_ints(sub)[0], canon_to_ints(chal[:n_vars])


def verify_synth_proof(n_vars, degree, seed, sum_int, proof_bytes, challenges):
    """Oracle-side check of a prove_partial proof of ONE product of `degree` seeded synthetic tables that is too large for the
    oracle's prover: replay the transcript with the oracle's verifier (verify_internal, multi_composed_sumcheck.rs:151-181:
    challenges re-derived from the proof bytes, p(0) + p(1) chain) and close the final claim against the oracle's OWN streamed
    evaluation of the tables at those challenges (Multilinear::evaluation, evaluation_form.rs:162-175).  A round polynomial
    that differed from the honest prover's survives this with probability ~ n d / |F|: an accepted proof is the reference
    prover's proof.  -> (ok, reason)"""
    per_round = 64 * (degree + 1)          # random tables: every round polynomial has degree + 1 monomials of (coeff, pow)
    if len(proof_bytes) != n_vars * per_round:
        return False, "proof length %d != %d rounds x %d bytes" % (len(proof_bytes), n_vars, per_round)
    rps = []
    for r in range(n_vars):
        blob = proof_bytes[r * per_round:(r + 1) * per_round]
        rps.append([(int.from_bytes(blob[64 * i:64 * i + 32], "big"), int.from_bytes(blob[64 * i + 32:64 * i + 64], "big")) for i in range(degree + 1)])
    ok, sub, ch = verify_partial(n_vars, sum_int, rps)
    if not ok:
        return False, "the oracle's verifier rejects the proof (p(0) + p(1) != claim in some round)"
    if list(ch) != list(challenges):
        return False, "challenges differ from the oracle's replay of the transcript"
    R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    val = 1
    for k in range(degree):
        val = val * eval_synth(seed, k, n_vars, ch) % R
    if val != sub:
        return False, "final claim does not match the oracle's evaluation of the tables at the challenges"
    return True, "ok"
