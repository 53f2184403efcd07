// Auxiliary (non-template) kernels: stand-alone fold, element-wise / outer operations, serialisation
// and the seeded synthetic-table generator.  Included by zksc.cu only.
#pragma once
#include "kernels.cuh"

namespace zksc {

// ---- stand-alone fold: Multilinear::partial_evaluation(r, k) for any variable index k -------------
// (polynomial/src/multilinear/evaluation_form.rs:123-141 + polynomial/src/utils.rs:26-53)
// out[o] = fold(in[i], in[i + s]),  s = n / 2^(k+1),  i = (o / s) * 2s + (o % s).  `out` may alias `in`
// only for k == 0.
struct FoldArgs {
    const Fr* in;
    Fr* out;
    unsigned long long in_tab_stride, in_proof_stride, out_tab_stride, out_proof_stride;
    unsigned long long n_out;   // n / 2
    unsigned long long s;       // pair distance
    unsigned int n_tabs;
    Fr chal[kMaxBatch];
};
__global__ void __launch_bounds__(256) fold_kernel(const __grid_constant__ FoldArgs args) {
    const int proof = blockIdx.y;
    const Fr r = args.chal[proof];
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned int k = 0; k < args.n_tabs; k++) {
        const Fr* in = args.in + (size_t)proof * args.in_proof_stride + (size_t)k * args.in_tab_stride;
        Fr* out = args.out + (size_t)proof * args.out_proof_stride + (size_t)k * args.out_tab_stride;
        for (unsigned long long o = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; o < args.n_out; o += stride) {
            unsigned long long i = (o / args.s) * 2 * args.s + (o % args.s);
            Fr y1 = ld256(in + i), y2 = ld256(in + i + args.s);
            st256(out + o, fr_fold(y1, y2, r));
        }
    }
}

// ---- element-wise helpers behind the Multilinear operators ---------------------------------------
enum EwOp : int { EW_ADD = 0, EW_SUB = 1, EW_MUL = 2, EW_SCALE = 3, EW_TO_MONT = 4, EW_FROM_MONT = 5 };
// out[i] = a[i] (op) b[i]      (EW_SCALE / conversions: b is a single element or unused)
__global__ void __launch_bounds__(256) ew_kernel(int op, const Fr* a, const Fr* b, Fr* out, unsigned long long n) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    Fr s = fr_zero();
    if (op == EW_SCALE) s = ld256(b);
    Fr one = fr_zero();
    one.l[0] = 1u;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Fr x = ld256(a + i), y;
        switch (op) {
            case EW_ADD: y = fr_add(x, ld256(b + i)); break;
            case EW_SUB: y = fr_sub(x, ld256(b + i)); break;
            case EW_MUL: y = fr_mul(x, ld256(b + i)); break;
            case EW_SCALE: y = fr_mul(x, s); break;
            case EW_TO_MONT: y = fr_mul(fr_canon(x), fr_r2()); break;
            default: y = fr_mul(x, one); break;  // EW_FROM_MONT
        }
        st256(out + i, y);
    }
}

// add_distinct / mul_distinct (evaluation_form.rs:28-52): out[i * nb + j] = a[i] (+|*) b[j]
__global__ void __launch_bounds__(256) outer_kernel(int mul, const Fr* a, unsigned long long na, const Fr* b, unsigned long long nb, Fr* out) {
    const unsigned long long n = na * nb, stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long o = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += stride) {
        Fr x = ld256(a + o / nb), y = ld256(b + o % nb);
        st256(out + o, mul ? fr_mul(x, y) : fr_add(x, y));
    }
}

// The same, written into this rank's shard of a table of a zksc_tables handle: out[i'] = a[i / nb] (+|*) b[i % nb],
// i = i' * shard_count + shard_index  (GKR: W(b) + W(c), W(b) * W(c), gkr/src/protocol.rs:80-81).
__global__ void __launch_bounds__(256) outer_fill_kernel(int mul, const Fr* a, const Fr* b, unsigned long long nb, Fr* out, unsigned long long n_local,
                                                         unsigned long long shard_index, unsigned long long shard_count) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long il = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; il < n_local; il += stride) {
        const unsigned long long o = il * shard_count + shard_index;
        Fr x = ld256(a + o / nb), y = ld256(b + o % nb);
        st256(out + il, mul ? fr_mul(x, y) : fr_add(x, y));
    }
}
// out[idx[i]] = vals[i] for the entries of this rank's shard (the table has been zeroed): what folding the gate-label
// variables of a 0/1 wiring table leaves behind (circuit/src/circuit.rs:57-95 + gkr/src/protocol.rs:70-74, 86-88).
__global__ void __launch_bounds__(256) scatter_kernel(const unsigned long long* idx, const Fr* vals, unsigned long long count, Fr* out,
                                                      unsigned long long shard_index, unsigned long long shard_count) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count && idx[i] % shard_count == shard_index) st256(out + idx[i] / shard_count, ld256(vals + i));
}

// Multilinear::to_bytes (evaluation_form.rs:54-62): Montgomery -> canonical -> 32 big-endian bytes
__global__ void __launch_bounds__(256) to_bytes_kernel(const Fr* a, Fr* out, unsigned long long n) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    Fr one = fr_zero();
    one.l[0] = 1u;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Fr x = fr_mul(ld256(a + i), one), y;
#pragma unroll
        for (int k = 0; k < 8; k++) y.l[k] = __byte_perm(x.l[7 - k], 0, 0x0123);
        st256(out + i, y);
    }
}

// ---- seeded synthetic tables (same convention as oracle/pymodel.py synth_entry) -------------------
ZKSC_DEV unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// out[i'] = entry(seed, table, i' * shard_count + shard_index), Montgomery form
__global__ void __launch_bounds__(256) synth_kernel(Fr* out, unsigned long long n_local, unsigned long long seed, unsigned long long table,
                                                    unsigned long long shard_index, unsigned long long shard_count) {
    const unsigned long long base = splitmix64(seed ^ splitmix64(table * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull));
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long il = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; il < n_local; il += stride) {
        const unsigned long long i = il * shard_count + shard_index;
        Fr x;
#pragma unroll
        for (int limb = 0; limb < 4; limb++) {
            unsigned long long w = splitmix64(base + (i * 4 + limb) * 0x9E3779B97F4A7C15ull);
            x.l[2 * limb] = (uint32_t)w;
            x.l[2 * limb + 1] = (uint32_t)(w >> 32);
        }
        st256(out + il, fr_mul(fr_canon(x), fr_r2()));
    }
}

}  // namespace zksc
