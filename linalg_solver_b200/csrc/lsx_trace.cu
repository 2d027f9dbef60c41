// Step trace of row_reduce: the reference records, next to the result, one label and
// one intermediate matrix per elementary step (linalg.py:544-629: S = row swap, N = normalisation of the pivot row,
// E = elimination below / above a pivot).  This kernel replays exactly that operation order modulo each of K primes
// -- one CTA per prime, the residue matrix in shared memory -- and writes the op log and the residue matrix after
// every recorded step; the host checks that all primes logged the same steps (a prime dividing an intermediate
// value would not) and lifts the residues to rationals (CRT + rational reconstruction, convert.py).
// It is a small-matrix feature (the reference uses it for step-by-step output), not a throughput path: plain
// 64-bit modular arithmetic, no Montgomery form.
#include "lsx_internal.h"

namespace {

constexpr int TR_THREADS = 128;
constexpr int TR_MAX_CELLS = 4096;     // m * n words of shared memory

struct TraceArgs {
    const int32_t* A;        // [m][n]
    const PrimeRec* primes;  // [K]
    const uint32_t* den;     // [K] residues of the common denominator of the input (NULL: 1)
    int m, n, bar, max_ops;
    int32_t* ops;            // [K][max_ops][4]: kind (1 S, 2 N, 3 E below, 4 E above), a, b, 0
    uint32_t* frames;        // [K][max_ops][m][n] plain residues after each op
    int32_t* n_ops;          // [K]
    int32_t* pivot_col;      // [K][min(m, bar)], -1 padded
};

__device__ __forceinline__ uint32_t mulmod(uint32_t a, uint32_t b, uint32_t p) { return (uint32_t)((uint64_t)a * b % p); }
__device__ uint32_t invmod(uint32_t a, uint32_t p) {
    uint32_t r = 1u, e = p - 2u;
    while (e) {
        if (e & 1u) r = mulmod(r, a, p);
        a = mulmod(a, a, p);
        e >>= 1;
    }
    return r;
}

__global__ void __launch_bounds__(TR_THREADS) k_rref_trace(TraceArgs a) {
    __shared__ uint32_t W[TR_MAX_CELLS];
    __shared__ uint32_t fcol[TR_MAX_CELLS];   // factors of one column (m <= TR_MAX_CELLS)
    __shared__ int s_src, s_hit;
    const int m = a.m, n = a.n, bar = a.bar, tid = threadIdx.x, k = blockIdx.x;
    const uint32_t p = a.primes[k].p;
    const int slots = m < bar ? m : bar;
    int32_t* ops = a.ops + (int64_t)k * a.max_ops * 4;
    uint32_t* frames = a.frames + (int64_t)k * a.max_ops * m * n;
    // rational input: the matrix is A / D with one common denominator D; its residues are A * D^-1.  A prime that
    // divides D cannot represent the matrix: it reports n_ops = -1 and the host drops it.
    const uint32_t dres = a.den ? a.den[k] % p : 1u;
    if (dres == 0u) {
        if (tid == 0) a.n_ops[k] = -1;
        return;
    }
    const uint32_t dinv = dres == 1u ? 1u : invmod(dres, p);
    for (int e = tid; e < m * n; e += TR_THREADS) W[e] = mulmod(word_of_int_any(a.A[e], p), dinv, p);
    for (int e = tid; e < slots; e += TR_THREADS) a.pivot_col[(int64_t)k * slots + e] = -1;
    __syncthreads();
    int step = 0;
    auto record = [&](int kind, int x, int y) {        // called by every thread after a barrier
        if (step < a.max_ops) {
            if (tid == 0) {
                ops[step * 4 + 0] = kind;
                ops[step * 4 + 1] = x;
                ops[step * 4 + 2] = y;
                ops[step * 4 + 3] = 0;
            }
            for (int e = tid; e < m * n; e += TR_THREADS) frames[(int64_t)step * m * n + e] = W[e];
        }
        ++step;
    };
    // eliminate column c with pivot row pr from the rows [k0, k1): row -= f * pivot_row on columns >= c
    auto eliminate = [&](int pr, int c, int k0, int k1) -> bool {
        if (tid == 0) s_hit = 0;
        __syncthreads();
        for (int r = k0 + tid; r < k1; r += TR_THREADS) {
            const uint32_t f = W[r * n + c];
            fcol[r] = f;
            if (f) s_hit = 1;
        }
        __syncthreads();
        const int width = n - c;
        for (int e = tid; e < (k1 - k0) * width; e += TR_THREADS) {
            const int r = k0 + e / width, j = c + e % width;
            const uint32_t f = fcol[r];
            if (f) {
                const uint32_t sub = mulmod(f, W[pr * n + j], p);
                const uint32_t w = W[r * n + j];
                W[r * n + j] = w >= sub ? w - sub : w + p - sub;
            }
        }
        __syncthreads();
        return s_hit != 0;
    };
    int pi = 0, pj = 0;
    while (pi < m && pj < bar) {
        if (W[pi * n + pj] == 0u) {
            if (tid == 0) {
                int src = -1;
                for (int i = pi + 1; i < m; ++i)
                    if (W[i * n + pj] != 0u) {
                        src = i;
                        break;
                    }
                s_src = src;
            }
            __syncthreads();
            const int src = s_src;
            __syncthreads();
            if (src < 0) {
                ++pj;
                continue;
            }
            for (int j = tid; j < n; j += TR_THREADS) {
                const uint32_t t0 = W[pi * n + j];
                W[pi * n + j] = W[src * n + j];
                W[src * n + j] = t0;
            }
            __syncthreads();
            record(1, pi, src);
            __syncthreads();
        }
        const uint32_t lead = W[pi * n + pj];
        __syncthreads();
        if (lead != 1u) {
            const uint32_t inv = invmod(lead, p);
            for (int j = pj + tid; j < n; j += TR_THREADS) W[pi * n + j] = mulmod(W[pi * n + j], inv, p);
            __syncthreads();
            record(2, pi, 0);
            __syncthreads();
        }
        if (eliminate(pi, pj, pi + 1, m)) {
            record(3, pj, 0);
            __syncthreads();
        }
        if (tid == 0) a.pivot_col[(int64_t)k * slots + pi] = pj;
        ++pi;
        ++pj;
    }
    __syncthreads();
    for (int idx = pi - 1; idx >= 0; --idx) {
        const int c = a.pivot_col[(int64_t)k * slots + idx];
        if (eliminate(idx, c, 0, idx)) {
            record(4, c, 0);
            __syncthreads();
        }
    }
    if (tid == 0) a.n_ops[k] = step;
}

}  // namespace

extern "C" {

int lsx_rref_trace_max_ops(int m, int n, int bar_col) {
    if (m < 1 || n < 1 || bar_col < 1 || bar_col > n) return LSX_ERR_BAD_SHAPE;
    const int r = m < bar_col ? m : bar_col;
    return 4 * r;          // per pivot at most S, N, E below and, in the backward sweep, E above
}

int lsx_rref_trace_q(lsx_ctx* ctx, const int32_t* A, const uint32_t* den_residues, int m, int n, int bar_col,
                     int n_primes, int mem, int32_t* ops, uint32_t* frames, int32_t* n_ops, int32_t* pivot_col) {
    if (!ctx) return LSX_ERR_NULL;
    if (!A || !ops || !frames || !n_ops || !pivot_col) return lsx_fail(ctx, LSX_ERR_NULL, "rref_trace: NULL buffer");
    if (m < 1 || n < 1 || bar_col < 1 || bar_col > n || (int64_t)m * n > TR_MAX_CELLS || m > 4096)
        return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "rref_trace: needs 1 <= bar_col <= n and m * n <= %d", TR_MAX_CELLS);
    if (n_primes < 1 || n_primes > LSX_TABLE_PRIMES) return lsx_fail(ctx, LSX_ERR_BOUND, "rref_trace: bad prime count");
    if (mem != LSX_MEM_HOST && mem != LSX_MEM_DEVICE) return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "bad mem flag %d", mem);
    LSX_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int max_ops = lsx_rref_trace_max_ops(m, n, bar_col), slots = m < bar_col ? m : bar_col;
    const size_t b_a = (size_t)m * n * 4, b_ops = (size_t)n_primes * max_ops * 16,
                 b_fr = (size_t)n_primes * max_ops * m * n * 4, b_n = (size_t)n_primes * 4, b_pc = (size_t)n_primes * slots * 4;
    TraceArgs t{};
    t.primes = ctx->d_primes;
    t.m = m, t.n = n, t.bar = bar_col, t.max_ops = max_ops;
    if (mem == LSX_MEM_DEVICE) {
        t.A = A, t.den = den_residues, t.ops = ops, t.frames = frames, t.n_ops = n_ops, t.pivot_col = pivot_col;
        k_rref_trace<<<n_primes, TR_THREADS, 0, ctx->stream>>>(t);
        ctx->launches++;
        LSX_CUDA_TRY(ctx, cudaGetLastError());
        return LSX_OK;
    }
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    int rc = lsx_ws_reserve(ctx, up(b_a) + up(b_ops) + up(b_fr) + up(b_n) + up(b_pc) + up(b_n));
    if (rc != LSX_OK) return rc;
    char* base = (char*)ctx->d_ws;
    int32_t* dA = (int32_t*)base;
    t.A = dA;
    t.ops = (int32_t*)(base + up(b_a));
    t.frames = (uint32_t*)(base + up(b_a) + up(b_ops));
    t.n_ops = (int32_t*)(base + up(b_a) + up(b_ops) + up(b_fr));
    t.pivot_col = (int32_t*)(base + up(b_a) + up(b_ops) + up(b_fr) + up(b_n));
    uint32_t* d_den = (uint32_t*)(base + up(b_a) + up(b_ops) + up(b_fr) + up(b_n) + up(b_pc));
    LSX_CUDA_TRY(ctx, cudaMemcpyAsync(dA, A, b_a, cudaMemcpyHostToDevice, ctx->stream));
    if (den_residues) {
        LSX_CUDA_TRY(ctx, cudaMemcpyAsync(d_den, den_residues, b_n, cudaMemcpyHostToDevice, ctx->stream));
        t.den = d_den;
    }
    LSX_CUDA_TRY(ctx, cudaMemsetAsync(t.ops, 0, b_ops, ctx->stream));
    k_rref_trace<<<n_primes, TR_THREADS, 0, ctx->stream>>>(t);
    ctx->launches++;
    LSX_CUDA_TRY(ctx, cudaGetLastError());
    LSX_CUDA_TRY(ctx, cudaMemcpyAsync(ops, t.ops, b_ops, cudaMemcpyDeviceToHost, ctx->stream));
    LSX_CUDA_TRY(ctx, cudaMemcpyAsync(frames, t.frames, b_fr, cudaMemcpyDeviceToHost, ctx->stream));
    LSX_CUDA_TRY(ctx, cudaMemcpyAsync(n_ops, t.n_ops, b_n, cudaMemcpyDeviceToHost, ctx->stream));
    LSX_CUDA_TRY(ctx, cudaMemcpyAsync(pivot_col, t.pivot_col, b_pc, cudaMemcpyDeviceToHost, ctx->stream));
    LSX_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LSX_OK;
}

int lsx_rref_trace(lsx_ctx* ctx, const int32_t* A, int m, int n, int bar_col, int n_primes, int mem, int32_t* ops,
                   uint32_t* frames, int32_t* n_ops, int32_t* pivot_col) {
    return lsx_rref_trace_q(ctx, A, nullptr, m, n, bar_col, n_primes, mem, ops, frames, n_ops, pivot_col);
}

}  // extern "C"
