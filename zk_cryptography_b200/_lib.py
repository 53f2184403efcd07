"""ctypes binding of include/zksc.h (the reference-side binding a Rust maintainer would write with
`extern "C"` is shown in INTEGRATION.md; this is the Python twin used by tests and bench.py)."""
import ctypes
import os

import numpy as np

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001

PROTO_SUMCHECK, PROTO_COMPOSED, PROTO_MULTI_PARTIAL, PROTO_MULTI_FULL = 0, 1, 2, 3

_ERR = {-1: "NO_DEVICE", -2: "CUDA", -3: "SHAPE", -4: "STATE", -5: "UNSUPPORTED", -6: "OOM", -7: "VERIFY", -8: "COMM"}


class ZkscError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__(f"zksc error {_ERR.get(code, code)}: {msg}")


_LIB = None
_u64p = ctypes.POINTER(ctypes.c_uint64)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def lib_path():
    # ZKSC_LIB: an alternative build of the same library (kernel experiments: tools/gpu_variants.sh)
    return os.environ.get("ZKSC_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libzksc.so")


def lib():
    """Load libzksc.so (built in-tree by __graft_entry__.build()).  Fails loudly when it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ZkscError(-1, f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
    L = ctypes.CDLL(path)
    vp = ctypes.c_void_p
    sig = {
        "zksc_version": (ctypes.c_char_p, []),
        "zksc_device_count": (ctypes.c_int, []),
        "zksc_ctx_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(vp)]),
        "zksc_ctx_create_multi": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.POINTER(vp)]),
        "zksc_ctx_devices": (ctypes.c_int, [vp]),
        "zksc_ctx_destroy": (ctypes.c_int, [vp]),
        "zksc_last_error": (ctypes.c_char_p, [vp]),
        "zksc_comm_unique_id": (ctypes.c_int, [_u8p]),
        "zksc_comm_init": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, _u8p]),
        "zksc_ctx_rank": (ctypes.c_int, [vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
        "zksc_ctx_synchronize": (ctypes.c_int, [vp]),
        "zksc_ctx_stream": (vp, [vp]),
        "zksc_ctx_launch_count": (ctypes.c_ulonglong, [vp]),
        "zksc_ctx_peer_exchange": (ctypes.c_int, [vp]),
        "zksc_ctx_timing": (ctypes.c_int, [vp, ctypes.c_int]),
        "zksc_ctx_timing_read": (ctypes.c_int, [vp, ctypes.c_uint32, _u32p, ctypes.POINTER(ctypes.c_float), _u32p, _u32p, _u64p, _u64p]),
        "zksc_ctx_gather_entries": (ctypes.c_uint64, [vp]),
        "zksc_ctx_round_times": (ctypes.c_int, [vp, ctypes.c_uint32, _u32p, ctypes.POINTER(ctypes.c_double)]),
        "zksc_int_peak": (ctypes.c_int, [vp, ctypes.POINTER(ctypes.c_double)]),
        "zksc_tables_reupload": (ctypes.c_int, [vp, ctypes.POINTER(_u64p), ctypes.c_int]),
        "zksc_tables_reupload_begin": (ctypes.c_int, [vp, ctypes.POINTER(_u64p), ctypes.c_int]),
        "zksc_tables_reupload_end": (ctypes.c_int, [vp]),
        "zksc_tables_upload_local": (ctypes.c_int, [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, _u32p, ctypes.POINTER(_u64p), ctypes.POINTER(vp)]),
        "zksc_tables_read_local": (ctypes.c_int, [vp, _u64p]),
        "zksc_tables_upload": (ctypes.c_int, [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, _u32p, ctypes.POINTER(_u64p), ctypes.POINTER(vp)]),
        "zksc_tables_synth": (ctypes.c_int, [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, _u32p, ctypes.c_uint64, ctypes.POINTER(vp)]),
        "zksc_tables_alloc": (ctypes.c_int, [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, _u32p, ctypes.POINTER(vp)]),
        "zksc_tables_fill_outer": (ctypes.c_int, [vp, ctypes.c_uint32, ctypes.c_int, _u64p, ctypes.c_uint64, _u64p, ctypes.c_uint64]),
        "zksc_tables_fill_sparse": (ctypes.c_int, [vp, ctypes.c_uint32, _u64p, _u64p, ctypes.c_uint64]),
        "zksc_tables_fill_dense": (ctypes.c_int, [vp, ctypes.c_uint32, _u64p]),
        "zksc_tables_free": (ctypes.c_int, [vp]),
        "zksc_tables_reset": (ctypes.c_int, [vp]),
        "zksc_tables_vars_left": (ctypes.c_int, [vp, _u32p]),
        "zksc_round_evals": (ctypes.c_int, [vp, _u64p]),
        "zksc_bind": (ctypes.c_int, [vp, _u64p]),
        "zksc_residual": (ctypes.c_int, [vp, _u64p]),
        "zksc_poly_sum": (ctypes.c_int, [vp, _u64p]),
        "zksc_tables_to_bytes": (ctypes.c_int, [vp, ctypes.c_uint32, _u8p]),
        "zksc_prove": (ctypes.c_int, [vp, ctypes.c_int, _u64p, _u64p, _u32p, _u64p]),
        "zksc_msg_stride": (ctypes.c_uint32, [ctypes.c_int, ctypes.c_uint32, _u32p]),
        "zksc_proof_to_bytes": (ctypes.c_int, [ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, _u64p, _u32p, _u8p, ctypes.POINTER(ctypes.c_size_t)]),
        "zksc_verify_rounds": (ctypes.c_int, [ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, _u64p, _u64p, _u32p, _u8p, ctypes.c_size_t, _u64p, _u64p]),
        "zksc_evaluate": (ctypes.c_int, [vp, _u64p, _u64p]),
        "zksc_gkr_total_rounds": (ctypes.c_uint64, [ctypes.c_uint32]),
        "zksc_gkr_prove": (ctypes.c_int, [vp, ctypes.c_uint32, _u32p, _u8p, _u32p, _u32p, ctypes.POINTER(_u64p), _u64p, _u64p, _u64p, _u64p, _u64p, _u64p, _u32p, _u64p]),
        "zksc_g1_msm": (ctypes.c_int, [vp, _u64p, _u64p, ctypes.c_uint64, _u64p]),
        "zksc_kzg_open": (ctypes.c_int, [vp, _u64p, ctypes.c_uint32, _u64p, _u64p, _u64p, _u64p]),
        "zksc_pairing_check": (ctypes.c_int, [_u64p, _u64p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_int)]),
        "zksc_pairing": (ctypes.c_int, [_u64p, _u64p, _u64p]),
        "zksc_kzg_verify": (ctypes.c_int, [_u64p, _u64p, _u64p, _u64p, _u64p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_int)]),
        "zksc_circuit_create": (ctypes.c_int, [vp, ctypes.c_uint32, _u32p, _u8p, _u32p, _u32p, ctypes.POINTER(vp)]),
        "zksc_circuit_free": (ctypes.c_int, [vp]),
        "zksc_circuit_evaluate": (ctypes.c_int, [vp, _u64p, _u64p]),
        "zksc_circuit_layer_values": (ctypes.c_int, [vp, ctypes.c_uint32, _u64p]),
        "zksc_circuit_total_rounds": (ctypes.c_uint64, [vp]),
        "zksc_gkr_prove_linear": (ctypes.c_int, [vp, _u64p, _u64p, _u64p, _u64p, _u64p, _u32p, _u64p]),
        "zksc_circuit_wiring_eval": (ctypes.c_int, [vp, ctypes.c_uint32, _u64p, _u64p, _u64p, _u64p, _u64p, _u64p, _u64p]),
        "zksc_ml_partial_evaluation": (ctypes.c_int, [vp, _u64p, ctypes.c_uint64, _u64p, ctypes.c_uint32, _u64p]),
        "zksc_ml_evaluation": (ctypes.c_int, [vp, _u64p, ctypes.c_uint64, _u64p, ctypes.c_uint32, _u64p]),
        "zksc_ml_outer": (ctypes.c_int, [vp, ctypes.c_int, _u64p, ctypes.c_uint64, _u64p, ctypes.c_uint64, _u64p]),
        "zksc_ml_elementwise": (ctypes.c_int, [vp, ctypes.c_int, _u64p, _u64p, ctypes.c_uint64, _u64p]),
        "zksc_fr_from_u64": (None, [ctypes.c_uint64, _u64p]),
        "zksc_fr_from_canonical": (None, [_u64p, _u64p]),
        "zksc_fr_to_canonical": (None, [_u64p, _u64p]),
        "zksc_fr_from_canonical_batch": (None, [_u64p, ctypes.c_uint64, _u64p]),
        "zksc_fr_to_canonical_batch": (None, [_u64p, ctypes.c_uint64, _u64p]),
        "zksc_fr_to_be_bytes": (None, [_u64p, _u8p]),
        "zksc_fr_from_be_bytes_mod_order": (None, [_u8p, _u64p]),
        "zksc_fr_add": (None, [_u64p, _u64p, _u64p]),
        "zksc_fr_sub": (None, [_u64p, _u64p, _u64p]),
        "zksc_fr_mul": (None, [_u64p, _u64p, _u64p]),
        "zksc_transcript_new": (vp, []),
        "zksc_transcript_free": (None, [vp]),
        "zksc_transcript_commit": (None, [vp, ctypes.c_char_p, ctypes.c_size_t]),
        "zksc_transcript_challenge": (None, [vp, _u8p]),
        "zksc_transcript_challenge_field": (None, [vp, _u64p]),
        "zksc_sparse_interpolate": (ctypes.c_uint32, [_u64p, ctypes.c_uint32, _u64p]),
        "zksc_sparse_add": (ctypes.c_uint32, [_u64p, ctypes.c_uint32, _u64p, ctypes.c_uint32, _u64p]),
        "zksc_sparse_evaluate": (None, [_u64p, ctypes.c_uint32, _u64p, _u64p]),
        "zksc_round_slots_to_evals": (None, [ctypes.c_uint32, _u64p]),
        "zksc_synth_entry": (None, [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, _u64p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here == a symbol include/zksc.h declares is missing
        fn.restype = res
        fn.argtypes = args
    L._zksc_signatures = sig
    _LIB = L
    return L


def p64(a):
    return a.ctypes.data_as(_u64p)


def p32(a):
    return a.ctypes.data_as(_u32p)


def p8(a):
    return a.ctypes.data_as(_u8p)


def ints_to_limbs(values):
    """canonical python ints -> (n,4) uint64 little-endian limbs (still canonical)."""
    n = len(values)
    buf = bytearray(32 * n)
    for i, v in enumerate(values):
        buf[32 * i:32 * i + 32] = (v % R_MOD).to_bytes(32, "little")
    return np.frombuffer(bytes(buf), dtype=np.uint64).reshape(n, 4).copy()


def limbs_to_ints(arr):
    b = np.ascontiguousarray(arr, dtype=np.uint64).tobytes()
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def to_mont(values):
    """python ints (any sign/size) -> (n,4) uint64 Montgomery-form array (ark-ff memory layout)."""
    single = isinstance(values, int)
    vals = [values] if single else list(values)
    c = ints_to_limbs(vals)
    out = np.empty_like(c)
    lib().zksc_fr_from_canonical_batch(p64(c), len(vals), p64(out))
    return out[0] if single else out


def from_mont(arr):
    """(n,4) (or (4,)) Montgomery array -> canonical python ints."""
    a = np.ascontiguousarray(arr, dtype=np.uint64)
    single = a.ndim == 1
    a2 = a.reshape(-1, 4)
    out = np.empty_like(a2)
    lib().zksc_fr_to_canonical_batch(p64(a2), a2.shape[0], p64(out))
    r = limbs_to_ints(out)
    return r[0] if single else r


class Context:
    """zksc_ctx: one CUDA device, one stream, optional NCCL communicator (one process per GPU)."""

    def __init__(self, device=0, devices=None):
        """device: one CUDA device ordinal; devices: a list of them -- ONE context (and one host thread) over several GPUs of this
        process (zksc_ctx_create_multi): tables are sharded over them, the caller sees a single logical context."""
        self._h = ctypes.c_void_p()
        if devices is not None:
            arr = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
            rc = lib().zksc_ctx_create_multi(arr, len(devices), ctypes.byref(self._h))
        else:
            rc = lib().zksc_ctx_create(device, ctypes.byref(self._h))
        if rc != 0:
            raise ZkscError(rc, (lib().zksc_last_error(None) or b"").decode())
        self.rank, self.n_ranks = 0, 1

    def devices(self):
        return int(lib().zksc_ctx_devices(self._h))

    def check(self, rc):
        if rc != 0:
            raise ZkscError(rc, (lib().zksc_last_error(self._h) or b"").decode())

    def comm_init(self, n_ranks, rank, unique_id):
        uid = np.frombuffer(bytes(unique_id), dtype=np.uint8).copy()
        self.check(lib().zksc_comm_init(self._h, n_ranks, rank, p8(uid)))
        self.rank, self.n_ranks = rank, n_ranks

    @staticmethod
    def unique_id():
        uid = np.zeros(128, dtype=np.uint8)
        rc = lib().zksc_comm_unique_id(p8(uid))
        if rc != 0:
            raise ZkscError(rc, (lib().zksc_last_error(None) or b"").decode())
        return uid.tobytes()

    def synchronize(self):
        self.check(lib().zksc_ctx_synchronize(self._h))

    def stream_handle(self):
        """The cudaStream_t (as int) all kernels of this context run on."""
        return int(lib().zksc_ctx_stream(self._h) or 0)

    def peer_exchange(self):
        return bool(lib().zksc_ctx_peer_exchange(self._h))

    def launch_count(self):
        return int(lib().zksc_ctx_launch_count(self._h))

    def timing(self, enable):
        self.check(lib().zksc_ctx_timing(self._h, 1 if enable else 0))

    def timing_read(self, cap=4096):
        """-> list of dicts {ms, degree, fold, pairs, proofs}, one per round-kernel launch since the last read."""
        n = ctypes.c_uint32()
        ms = np.zeros(cap, dtype=np.float32)
        deg = np.zeros(cap, dtype=np.uint32)
        fold = np.zeros(cap, dtype=np.uint32)
        pairs = np.zeros(cap, dtype=np.uint64)
        proofs = np.zeros(cap, dtype=np.uint64)
        self.check(lib().zksc_ctx_timing_read(self._h, cap, ctypes.byref(n), ms.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), p32(deg), p32(fold),
                                              p64(pairs), p64(proofs)))
        return [dict(ms=float(ms[i]), degree=int(deg[i]), fold=int(fold[i]), pairs=int(pairs[i]), proofs=int(proofs[i])) for i in range(n.value)]

    def gather_entries(self):
        return int(lib().zksc_ctx_gather_entries(self._h))

    def round_times(self, cap=64):
        """-> host wall time (us) of every round of the latest prove on this context."""
        n = ctypes.c_uint32()
        us = np.zeros(cap, dtype=np.float64)
        self.check(lib().zksc_ctx_round_times(self._h, cap, ctypes.byref(n), us.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
        return [float(x) for x in us[:n.value]]

    def int_peak(self):
        """-> measured IMAD.WIDE.U32 rate of this device, limb products per second."""
        v = ctypes.c_double()
        self.check(lib().zksc_int_peak(self._h, ctypes.byref(v)))
        return v.value

    def close(self):
        if self._h:
            lib().zksc_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Tables:
    """zksc_tables: device-resident  sum_p prod_k f_{p,k}  for n_proofs independent instances."""

    def __init__(self, ctx, n_vars, degrees, handle, n_proofs):
        self.ctx, self.n_vars, self.degrees, self._h, self.n_proofs = ctx, n_vars, list(degrees), handle, n_proofs
        self.n_tables = sum(self.degrees)
        self.n_evals = sum(d + 1 for d in self.degrees)

    @staticmethod
    def upload(ctx, n_vars, degrees, tables, n_proofs=1):
        """tables: list (proof-major, then product, then factor) of (2^n_vars, 4) uint64 Montgomery arrays."""
        deg = np.asarray(degrees, dtype=np.uint32)
        arrs = [np.ascontiguousarray(t, dtype=np.uint64) for t in tables]
        if len(arrs) != n_proofs * int(deg.sum()):
            raise ZkscError(-3, "number of tables does not match n_proofs * sum(degrees)")
        for a in arrs:
            if a.shape != (1 << n_vars, 4):
                raise ZkscError(-3, "Number of evaluations must be a power of 2 and equal for every table")
        ptrs = (_u64p * len(arrs))(*[p64(a) for a in arrs])
        h = ctypes.c_void_p()
        ctx.check(lib().zksc_tables_upload(ctx._h, n_vars, n_proofs, len(deg), p32(deg), ptrs, ctypes.byref(h)))
        return Tables(ctx, n_vars, degrees, h, n_proofs)

    def reupload(self, tables, local=False):
        """Copy new host tables (same shapes) into this handle and reset it.  local: the arrays are this
        rank's shards (2^n_vars / n_ranks entries each) instead of full tables."""
        arrs = [np.ascontiguousarray(t, dtype=np.uint64) for t in tables]
        n = (1 << self.n_vars) // (self.ctx.n_ranks if local else 1)
        if len(arrs) != self.n_proofs * self.n_tables or any(a.shape != (n, 4) for a in arrs):
            raise ZkscError(-3, "tables do not match the shape this handle was created with")
        ptrs = (_u64p * len(arrs))(*[p64(a) for a in arrs])
        self.ctx.check(lib().zksc_tables_reupload(self._h, ptrs, 1 if local else 0))

    def reupload_begin(self, tables, local=False):
        """Start refilling this handle on the context's copy stream (zksc_tables_reupload_begin); the arrays must stay alive
        and untouched until reupload_end()."""
        arrs = [np.ascontiguousarray(t, dtype=np.uint64) for t in tables]
        n = (1 << self.n_vars) // (self.ctx.n_ranks if local else 1)
        if len(arrs) != self.n_proofs * self.n_tables or any(a.shape != (n, 4) for a in arrs):
            raise ZkscError(-3, "tables do not match the shape this handle was created with")
        self._refill = arrs
        ptrs = (_u64p * len(arrs))(*[p64(a) for a in arrs])
        self.ctx.check(lib().zksc_tables_reupload_begin(self._h, ptrs, 1 if local else 0))

    def reupload_end(self):
        self.ctx.check(lib().zksc_tables_reupload_end(self._h))
        self._refill = None

    @staticmethod
    def upload_local(ctx, n_vars, degrees, local_tables, n_proofs=1):
        """Sharded contexts: local_tables[t] holds this rank's entries (index = rank mod n_ranks) of table t."""
        deg = np.asarray(degrees, dtype=np.uint32)
        arrs = [np.ascontiguousarray(t, dtype=np.uint64) for t in local_tables]
        n = (1 << n_vars) // ctx.n_ranks
        if len(arrs) != n_proofs * int(deg.sum()) or any(a.shape != (n, 4) for a in arrs):
            raise ZkscError(-3, "local tables must hold 2^n_vars / n_ranks entries each")
        ptrs = (_u64p * len(arrs))(*[p64(a) for a in arrs])
        h = ctypes.c_void_p()
        ctx.check(lib().zksc_tables_upload_local(ctx._h, n_vars, n_proofs, len(deg), p32(deg), ptrs, ctypes.byref(h)))
        return Tables(ctx, n_vars, degrees, h, n_proofs)

    def read_local(self):
        """-> (n_proofs, n_tables, 2^n_vars / n_ranks, 4): this rank's copy of the tables as uploaded."""
        out = np.zeros((self.n_proofs, self.n_tables, (1 << self.n_vars) // self.ctx.n_ranks, 4), dtype=np.uint64)
        self.ctx.check(lib().zksc_tables_read_local(self._h, p64(out)))
        return out

    @staticmethod
    def synth(ctx, n_vars, degrees, seed, n_proofs=1):
        deg = np.asarray(degrees, dtype=np.uint32)
        h = ctypes.c_void_p()
        ctx.check(lib().zksc_tables_synth(ctx._h, n_vars, n_proofs, len(deg), p32(deg), seed, ctypes.byref(h)))
        return Tables(ctx, n_vars, degrees, h, n_proofs)

    @staticmethod
    def alloc(ctx, n_vars, degrees, n_proofs=1):
        """Uninitialised device tables, to be filled in place by fill_outer / fill_sparse / fill_dense."""
        deg = np.asarray(degrees, dtype=np.uint32)
        h = ctypes.c_void_p()
        ctx.check(lib().zksc_tables_alloc(ctx._h, n_vars, n_proofs, len(deg), p32(deg), ctypes.byref(h)))
        return Tables(ctx, n_vars, degrees, h, n_proofs)

    def fill_outer(self, table, mul, a, b):
        """table = a (+) b (mul False: add_distinct) or a (x) b (mul True: mul_distinct), computed on the device"""
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
        self.ctx.check(lib().zksc_tables_fill_outer(self._h, table, 1 if mul else 0, p64(a), a.shape[0], p64(b), b.shape[0]))

    def fill_sparse(self, table, idx, vals):
        """table = 0 except table[idx[i]] = vals[i] (Montgomery elements)"""
        idx = np.ascontiguousarray(idx, dtype=np.uint64).reshape(-1)
        vals = np.ascontiguousarray(vals, dtype=np.uint64).reshape(-1, 4)
        if idx.shape[0] != vals.shape[0]:
            raise ZkscError(-3, "one value per index")
        self.ctx.check(lib().zksc_tables_fill_sparse(self._h, table, p64(idx), p64(vals), idx.shape[0]))

    def fill_dense(self, table, evals):
        evals = np.ascontiguousarray(evals, dtype=np.uint64).reshape(-1, 4)
        if evals.shape[0] != 1 << self.n_vars:
            raise ZkscError(-3, "Number of evaluations must be 2^n_vars")
        self.ctx.check(lib().zksc_tables_fill_dense(self._h, table, p64(evals)))

    def reset(self):
        self.ctx.check(lib().zksc_tables_reset(self._h))

    def vars_left(self):
        v = ctypes.c_uint32()
        self.ctx.check(lib().zksc_tables_vars_left(self._h, ctypes.byref(v)))
        return v.value

    def round_evals(self):
        out = np.zeros((self.n_proofs, self.n_evals, 4), dtype=np.uint64)
        self.ctx.check(lib().zksc_round_evals(self._h, p64(out)))
        return out

    def bind(self, challenges):
        c = np.ascontiguousarray(challenges, dtype=np.uint64).reshape(self.n_proofs, 4)
        self.ctx.check(lib().zksc_bind(self._h, p64(c)))

    def residual(self):
        n = 1 << self.vars_left()
        out = np.zeros((self.n_proofs, self.n_tables, n, 4), dtype=np.uint64)
        self.ctx.check(lib().zksc_residual(self._h, p64(out)))
        return out

    def poly_sum(self):
        out = np.zeros((self.n_proofs, 4), dtype=np.uint64)
        self.ctx.check(lib().zksc_poly_sum(self._h, p64(out)))
        return out

    def to_bytes(self, proof=0):
        out = np.zeros(self.n_tables * (1 << self.n_vars) * 32, dtype=np.uint8)
        self.ctx.check(lib().zksc_tables_to_bytes(self._h, proof, p8(out)))
        return out.tobytes()

    def msg_stride(self, protocol):
        deg = np.asarray(self.degrees, dtype=np.uint32)
        return lib().zksc_msg_stride(protocol, len(deg), p32(deg))

    def prove(self, protocol, sums=None):
        """-> (round_msgs (B, n, stride, 4), round_len (B, n), challenges (B, n, 4)), all Montgomery."""
        stride = self.msg_stride(protocol)
        B, n = self.n_proofs, self.n_vars
        msgs = np.zeros((B, n, stride, 4), dtype=np.uint64)
        lens = np.zeros((B, n), dtype=np.uint32)
        chal = np.zeros((B, n, 4), dtype=np.uint64)
        s = None
        if sums is not None:
            s = np.ascontiguousarray(sums, dtype=np.uint64).reshape(B, 4)
        self.ctx.check(lib().zksc_prove(self._h, protocol, p64(s) if s is not None else None, p64(msgs), p32(lens), p64(chal)))
        return msgs, lens, chal

    def evaluate(self, points):
        pts = np.ascontiguousarray(points, dtype=np.uint64).reshape(self.n_proofs, self.n_vars, 4)
        out = np.zeros((self.n_proofs, 4), dtype=np.uint64)
        self.ctx.check(lib().zksc_evaluate(self._h, p64(pts), p64(out)))
        return out

    def free(self):
        if self._h:
            if self.ctx._h:          # a handle must not outlive its context: after Context.close() there is nothing left to free into
                lib().zksc_tables_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def proof_to_bytes(protocol, msgs, lens):
    """One proof: msgs (n, stride, 4), lens (n,) -> the bytes the reference feeds its transcripts."""
    msgs = np.ascontiguousarray(msgs, dtype=np.uint64)
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    n, stride = msgs.shape[0], msgs.shape[1] if msgs.ndim == 3 else 0
    ln = ctypes.c_size_t()
    rc = lib().zksc_proof_to_bytes(protocol, n, stride, p64(msgs), p32(lens), None, ctypes.byref(ln))
    if rc != 0:
        raise ZkscError(rc)
    out = np.zeros(max(ln.value, 1), dtype=np.uint8)
    rc = lib().zksc_proof_to_bytes(protocol, n, stride, p64(msgs), p32(lens), p8(out), ctypes.byref(ln))
    if rc != 0:
        raise ZkscError(rc)
    return out[:ln.value].tobytes()


def verify_rounds(protocol, sum_mont, msgs, lens, prefix=b""):
    """-> (subclaim_sum (4,), challenges (n,4)); raises ZkscError(VERIFY) like the reference's Err."""
    msgs = np.ascontiguousarray(msgs, dtype=np.uint64)
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    n = msgs.shape[0]
    stride = msgs.shape[1] if n else 0
    s = np.ascontiguousarray(sum_mont, dtype=np.uint64).reshape(4)
    sub = np.zeros(4, dtype=np.uint64)
    chal = np.zeros((max(n, 1), 4), dtype=np.uint64)
    pre = np.frombuffer(prefix, dtype=np.uint8).copy() if prefix else None
    rc = lib().zksc_verify_rounds(protocol, n, stride, p64(s), p64(msgs), p32(lens), p8(pre) if pre is not None else None, len(prefix), p64(sub), p64(chal))
    if rc != 0:
        raise ZkscError(rc, "Verification failed" if rc == -7 else "")
    return sub, chal[:n]
