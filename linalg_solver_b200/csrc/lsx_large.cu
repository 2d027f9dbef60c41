// One large determinant, sharded by prime (config 5 of BASELINE.json): residues det(A) mod p for a
// range of table primes, and the CRT of all residues to a signed multi-limb integer.
//
// The reference has no feasible route for this size (its determinant is a plan search over
// sparsity patterns, determinant.rs:575-665, or an n! sum, linalg.py:264-345); the value is the
// signed product of the forward-sweep pivots of linalg.py:547-609, computed here modulo each prime.
#include <algorithm>
#include <cmath>

#include "lsx_internal.h"

namespace {

// ---- Garner CRT for up to LSX_TABLE_PRIMES residues, one CTA ----------------------------------------
// Step j fixes the mixed-radix digit v_j = t_j and updates every later residue
// t_k <- (t_k - v_j) * p_j^{-1} mod p_k.  The inverses come from a table built by k_inv_table.
__global__ void k_inv_table(const PrimeRec* primes, int K, uint32_t* tab) {
    // tab[j * K + k] = (p_j^{-1} mod p_k) * R mod p_k  for j < k  (Montgomery word of the inverse)
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)K * K) return;
    const int j = (int)(t / K), k = (int)(t % K);
    if (j >= k) return;
    const PrimeRec P = primes[k];
    uint32_t pj = primes[j].p % P.p;
    const uint32_t w = mont_mul(pj, P.r2, P.p, P.pinv);                 // word of p_j
    tab[t] = mont_pow(w, P.p - 2u, P.one, P.p, P.pinv);                 // word of p_j^{-1}
}

__global__ void __launch_bounds__(1024) k_garner_big(const PrimeRec* primes, const uint32_t* tab, const uint32_t* res,
                                                      int K, int L, uint32_t* out, uint64_t* limb_ws) {
    extern __shared__ uint32_t sm[];
    uint32_t* t = sm;          // [K] running residues -> mixed-radix digits
    __shared__ int s_neg;
    const int tid = threadIdx.x, T = blockDim.x;
    for (int k = tid; k < K; k += T) t[k] = res[k] % primes[k].p;
    __syncthreads();
    for (int j = 0; j < K; ++j) {
        const uint32_t vj = t[j];
        for (int k = j + 1 + tid; k < K; k += T) {
            const PrimeRec P = primes[k];
            uint32_t v = vj % P.p;
            uint32_t x = t[k];
            x = x >= v ? x - v : x + P.p - v;
            t[k] = mont_mul(x, tab[(int64_t)j * K + k], P.p, P.pinv);
        }
        __syncthreads();
    }
    // sign: X > (M-1)/2 <=> digits compare above ((p_i - 1)/2)_i from the top
    if (tid == 0) {
        int neg = 0;
        for (int i = K - 1; i >= 0; --i) {
            const uint32_t h = (primes[i].p - 1u) >> 1;
            if (t[i] != h) {
                neg = t[i] > h;
                break;
            }
        }
        s_neg = neg;
    }
    __syncthreads();
    const bool neg = s_neg != 0;
    // negative: X - M = -(Y + 1) with Y = sum (p_i - 1 - v_i) P_i, so the result is ~Y
    if (neg)
        for (int k = tid; k < K; k += T) t[k] = primes[k].p - 1u - t[k];
    __syncthreads();
    // Horner from the top digit in a redundant limb form: limb l is a 64-bit word below 3 * 2^32
    // (low 32 bits + a small overflow part).  acc <- acc * p_i + v_i with limb l handled by thread l;
    // the two buffers live in global scratch (L can reach a few thousand limbs).
    uint64_t* cur = limb_ws;
    uint64_t* nxt = limb_ws + L;
    for (int l = tid; l < L; l += T) cur[l] = 0;
    __syncthreads();
    for (int i = K - 1; i >= 0; --i) {
        const uint64_t p = primes[i].p;
        const uint32_t v = t[i];
        for (int l = tid; l < L; l += T) {
            const uint64_t c = cur[l];
            uint64_t y = ((c & 0xffffffffull) * p) & 0xffffffffull;      // low word of lo * p
            if (l) {
                const uint64_t b = cur[l - 1];
                y += ((b & 0xffffffffull) * p) >> 32;                    // high word of the limb below
                y += (b >> 32) * p;                                      // its overflow part (< 3) times p
            } else {
                y += v;
            }
            nxt[l] = y;                                                  // < 2^32 + 2^32 + 2^32
        }
        __syncthreads();
        uint64_t* tmp = cur;
        cur = nxt;
        nxt = tmp;
    }
    if (tid == 0) {
        uint64_t carry = 0;
        for (int l = 0; l < L; ++l) {
            const uint64_t y = cur[l] + carry;
            const uint32_t w = (uint32_t)y;
            out[l] = neg ? ~w : w;
            carry = y >> 32;
        }
    }
}

// sums of squares of the rows ([0, n)) and of the columns ([n, 2n)) of an n x n int32 matrix.  Accumulated in
// double: a 64-bit integer sum wraps for full-range entries (n = 65 with |a| near 2^31 already exceeds 2^64), and
// the relative rounding error of n additions (n * 2^-53) is far below the slack the caller adds to the bound.
__global__ void k_sq_norms(const int32_t* __restrict__ A, int n, double* __restrict__ out) {
    const int i = blockIdx.x, tid = threadIdx.x;
    const bool is_col = i >= n;
    const int idx = is_col ? i - n : i;
    double s = 0.0;
    for (int t = tid; t < n; t += blockDim.x) {
        const double v = (double)(is_col ? A[(int64_t)t * n + idx] : A[(int64_t)idx * n + t]);
        s += v * v;
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double part[8];
    if ((tid & 31) == 0) part[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += part[w];
        out[i] = tot;
    }
}

// ---- rank of one matrix of any size: fraction-free elimination in global memory -----------------------------
// The batched rank keeps one residue tile per CTA in shared memory (m <= 254).  Beyond that the reference's rank()
// (linalg.py:745-747) still has to have an answer, so this is the plain route: G primes side by side in HBM
// ([G][m][n] words), per column one launch that finds each prime's first non-zero row at or below its pivot
// count (rows are exchanged through a permutation, not physically) and one launch that applies
// row <- piv * row - f * pivot_row to the rows below on the columns right of the pivot.  Montgomery's R^-1 per
// update scales whole rows and does not change which entries are zero.  rank over Q = max over primes of the rank
// modulo p as soon as the primes' product exceeds the Hadamard bound of the minors (a non-zero minor cannot vanish
// modulo all of them); full rank modulo ONE prime already proves full rank, which ends the loop early.
__global__ void k_rl_absmax(const int32_t* __restrict__ A, int64_t count, unsigned int* out) {
    unsigned int m = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = A[i];
        const unsigned int a = v < 0 ? 0u - (unsigned int)v : (unsigned int)v;
        m = a > m ? a : m;
    }
    for (int o = 16; o; o >>= 1) {
        const unsigned int t = __shfl_xor_sync(0xffffffffu, m, o);
        m = t > m ? t : m;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

__global__ void k_rl_load(const int32_t* __restrict__ A, int64_t cells, int m, const PrimeRec* primes, int prime0,
                          uint32_t* __restrict__ W, int32_t* __restrict__ perm, int32_t* __restrict__ npiv) {
    const int g = blockIdx.y;
    const uint32_t p = primes[prime0 + g].p;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (int64_t)gridDim.x * blockDim.x)
        W[(int64_t)g * cells + i] = word_of_int_any(A[i], p);
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < m; i += blockDim.x) perm[(int64_t)g * m + i] = i;
        if (threadIdx.x == 0) npiv[g] = 0;
    }
}

// one CTA per prime: first row (in permuted order) at or below the pivot count with a non-zero entry in column c
__global__ void __launch_bounds__(256) k_rl_pivot(const uint32_t* __restrict__ W, int m, int n, int c, int32_t* perm,
                                                   const int32_t* __restrict__ npiv, int32_t* __restrict__ pinfo) {
    const int g = blockIdx.x, tid = threadIdx.x;
    const int r = npiv[g];
    __shared__ int s_min;
    if (tid == 0) s_min = INT32_MAX;
    __syncthreads();
    int32_t* pm = perm + (int64_t)g * m;
    const uint32_t* Wg = W + (int64_t)g * m * n;
    if (r < m) {
        int best = INT32_MAX;
        for (int i = r + tid; i < m; i += 256)
            if (Wg[(int64_t)pm[i] * n + c] != 0u) {
                best = i;
                break;                                        // this thread's rows ascend: its first hit is its minimum
            }
        if (best != INT32_MAX) atomicMin(&s_min, best);
    }
    __syncthreads();
    if (tid == 0) {
        const int src = s_min;
        if (src == INT32_MAX) {
            pinfo[2 * g] = -1;                                // no pivot in this column (or the prime is done)
        } else {
            const int32_t a = pm[r], b = pm[src];
            pm[r] = b;
            pm[src] = a;
            pinfo[2 * g] = b;                                 // physical pivot row
            pinfo[2 * g + 1] = r;                             // its position
        }
    }
}

// rows below the pivot: row <- piv * row - f * pivot_row on the columns right of c
__global__ void __launch_bounds__(256) k_rl_elim(uint32_t* __restrict__ W, int m, int n, int c,
                                                  const int32_t* __restrict__ perm, int32_t* __restrict__ npiv,
                                                  const int32_t* __restrict__ pinfo, const PrimeRec* primes, int prime0) {
    const int g = blockIdx.y;
    const int prow = pinfo[2 * g];
    if (prow < 0) return;
    const int r = pinfo[2 * g + 1];
    const PrimeRec P = primes[prime0 + g];
    uint32_t* Wg = W + (int64_t)g * m * n;
    const uint32_t* pr = Wg + (int64_t)prow * n;
    const uint32_t piv = pr[c];
    const int32_t* pm = perm + (int64_t)g * m;
    // the rows below position r are split over the blocks of the x dimension, one warp per row at a time
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    for (int i = r + 1 + blockIdx.x * warps + warp; i < m; i += gridDim.x * warps) {
        uint32_t* row = Wg + (int64_t)pm[i] * n;
        const uint32_t f = row[c];
        if (f == 0u) continue;
        const uint32_t y = P.p - f;
        for (int j = c + 1 + lane; j < n; j += 32) row[j] = mont_fma2(piv, row[j], y, pr[j], P.p, P.pinv);
        __syncwarp();
        if (lane == 0) row[c] = 0u;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) npiv[g] = r + 1;
}

}  // namespace

extern "C" {

int lsx_det_large_prime_count_for(lsx_ctx* ctx, const int32_t* A, int n, int mem, int* n_primes, double* log2_bound) {
    if (!ctx) return LSX_ERR_NULL;
    if (!A || !n_primes) return lsx_fail(ctx, LSX_ERR_NULL, "det_large_prime_count_for: NULL buffer");
    if (n < 1) return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "det_large_prime_count_for: bad n");
    std::vector<double> sq((size_t)2 * n, 0.0);
    if (mem == LSX_MEM_HOST) {
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                const double v = (double)A[(int64_t)i * n + j];
                sq[i] += v * v;
                sq[n + j] += v * v;
            }
    } else {
        LSX_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
        int rc = lsx_ws_reserve(ctx, (size_t)2 * n * 8);
        if (rc != LSX_OK) return rc;
        k_sq_norms<<<2 * n, 256, 0, ctx->stream>>>(A, n, (double*)ctx->d_ws);
        ctx->launches++;
        LSX_CUDA_TRY(ctx, cudaMemcpyAsync(sq.data(), ctx->d_ws, (size_t)2 * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        LSX_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    double rows = 0.0, cols = 0.0;
    bool zero = false;
    for (int i = 0; i < n; ++i) {
        if (sq[i] == 0.0 || sq[n + i] == 0.0) zero = true;
        else rows += 0.5 * std::log2(sq[i]), cols += 0.5 * std::log2(sq[n + i]);
    }
    // a relative slack far above the rounding error of the double sums of squares (n * 2^-53 each), of 2n
    // logarithms and of their sum keeps the bound rigorous
    const double bits = zero ? 0.0 : std::min(rows, cols) * (1.0 + 1e-9) + 1e-6;
    int K, L;
    lsx_bits_to_plan(bits, &K, &L);
    if (K > LSX_TABLE_PRIMES) return LSX_ERR_BOUND;
    *n_primes = K;
    if (log2_bound) *log2_bound = bits;
    return LSX_OK;
}

int lsx_det_large_prime_count(int n, int64_t a_abs_max, int* n_primes, double* log2_bound) {
    if (n < 1 || a_abs_max < 0 || !n_primes) return LSX_ERR_BAD_SHAPE;
    const double a = (double)(a_abs_max < 1 ? 1 : a_abs_max);
    const double bits = n * (0.5 * std::log2((double)n) + std::log2(a));
    int K, L;
    lsx_bits_to_plan(bits, &K, &L);
    if (K > LSX_TABLE_PRIMES) return LSX_ERR_BOUND;
    *n_primes = K;
    if (log2_bound) *log2_bound = bits;
    return LSX_OK;
}

int lsx_det_large_residues(lsx_ctx* ctx, const int32_t* A, int n, int prime_begin, int prime_count, int mem,
                           uint32_t* residues, uint32_t* primes_out) {
    if (!ctx) return LSX_ERR_NULL;
    if (!A || !residues) return lsx_fail(ctx, LSX_ERR_NULL, "det_large: NULL buffer");
    if (n < 1 || prime_begin < 0 || prime_count < 0 || prime_begin + prime_count > LSX_TABLE_PRIMES)
        return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "det_large: bad n or prime range");
    LSX_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (primes_out) {
        if (mem == LSX_MEM_HOST)
            memcpy(primes_out, ctx->primes.data() + prime_begin, (size_t)prime_count * 4);
        else
            LSX_CUDA_TRY(ctx, cudaMemcpyAsync(primes_out, ctx->primes.data() + prime_begin, (size_t)prime_count * 4,
                                              cudaMemcpyHostToDevice, ctx->stream));
    }
    if (prime_count == 0) return LSX_OK;
    const int32_t* dA = A;
    uint32_t* dres = residues;
    void* stage = nullptr;
    if (mem == LSX_MEM_HOST) {
        const size_t abytes = (size_t)n * n * 4, rbytes = (size_t)prime_count * 4;
        LSX_CUDA_TRY(ctx, cudaMalloc(&stage, abytes + rbytes + 256));
        dA = (const int32_t*)stage;
        dres = (uint32_t*)((char*)stage + (abytes + 255) / 256 * 256);
        cudaError_t e = cudaMemcpyAsync(stage, A, abytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) {
            cudaFree(stage);
            return lsx_fail(ctx, LSX_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(e));
        }
    }
    int rc;
    if (lsx_tile_fits(n, n) && !getenv("LSX_FORCE_BLOCKED"))
        rc = lsx_tile_det_residues(ctx, dA, n, prime_begin, prime_count, dres);
    else
        rc = lsx_blocked_det_residues(ctx, dA, n, prime_begin, prime_count, dres);
    if (rc == LSX_OK && mem == LSX_MEM_HOST) {
        cudaError_t e = cudaMemcpyAsync(residues, dres, (size_t)prime_count * 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = lsx_fail(ctx, LSX_ERR_CUDA, "D2H copy failed: %s", cudaGetErrorString(e));
    }
    if (stage) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(stage);
    }
    return rc;
}

int lsx_rank_large(lsx_ctx* ctx, const int32_t* A, int m, int n, int mem, int32_t* rank, int32_t* primes_used) {
    if (!ctx) return LSX_ERR_NULL;
    if (!A || !rank) return lsx_fail(ctx, LSX_ERR_NULL, "rank_large: NULL buffer");
    if (m < 1 || n < 1) return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "rank_large: bad shape %d x %d", m, n);
    if (mem != LSX_MEM_HOST && mem != LSX_MEM_DEVICE) return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "bad mem flag %d", mem);
    LSX_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int64_t cells = (int64_t)m * n;
    const int full = m < n ? m : n;
    // group size: about 2 GiB of residues, at least one prime
    int G = (int)std::max<int64_t>(1, std::min<int64_t>(32, ((int64_t)2 << 30) / (cells * 4)));
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t b_a = up((size_t)cells * 4), b_w = up((size_t)G * cells * 4), b_perm = up((size_t)G * m * 4),
                 b_small = up((size_t)G * 16 + 64);
    int rc = lsx_ws_reserve(ctx, b_a + b_w + b_perm + b_small);
    if (rc != LSX_OK) return rc;
    char* base = (char*)ctx->d_ws;
    const int32_t* dA = A;
    if (mem == LSX_MEM_HOST) {
        LSX_CUDA_TRY(ctx, cudaMemcpyAsync(base, A, (size_t)cells * 4, cudaMemcpyHostToDevice, ctx->stream));
        dA = (const int32_t*)base;
    }
    uint32_t* W = (uint32_t*)(base + b_a);
    int32_t* perm = (int32_t*)(base + b_a + b_w);
    int32_t* npiv = (int32_t*)(base + b_a + b_w + b_perm);
    int32_t* pinfo = npiv + G;
    unsigned int* d_amax = (unsigned int*)(pinfo + 2 * G);
    // declared magnitude = the actual one; the prime count for a rigorous answer follows from the Hadamard bound
    LSX_CUDA_TRY(ctx, cudaMemsetAsync(d_amax, 0, 4, ctx->stream));
    k_rl_absmax<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(dA, cells, d_amax);
    ctx->launches++;
    unsigned int amax = 0;
    LSX_CUDA_TRY(ctx, cudaMemcpyAsync(&amax, d_amax, 4, cudaMemcpyDeviceToHost, ctx->stream));
    LSX_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    int best = 0, used = 0;
    if (amax != 0) {
        const double bits = lsx_log2_minor_bound(m, n, false, (int64_t)amax, 0, false, full);
        int K, L;
        lsx_bits_to_plan(bits, &K, &L);
        if (K > LSX_TABLE_PRIMES) return lsx_fail(ctx, LSX_ERR_BOUND, "rank_large: the bound needs %d primes", K);
        std::vector<int32_t> h_npiv(G);
        const int steps = n;
        const unsigned row_blocks = (unsigned)std::max(1, std::min(ctx->sm_count * 2, (m + 7) / 8));
        for (int p0 = 0; p0 < K && best < full; p0 += G) {
            const int g = std::min(G, K - p0);
            k_rl_load<<<dim3((unsigned)std::min<int64_t>(1024, (cells + 255) / 256), g), 256, 0, ctx->stream>>>(
                dA, cells, m, ctx->d_primes, p0, W, perm, npiv);
            ctx->launches++;
            for (int c = 0; c < steps; ++c) {
                k_rl_pivot<<<g, 256, 0, ctx->stream>>>(W, m, n, c, perm, npiv, pinfo);
                k_rl_elim<<<dim3(row_blocks, g), 256, 0, ctx->stream>>>(W, m, n, c, perm, npiv, pinfo, ctx->d_primes, p0);
                ctx->launches += 2;
            }
            LSX_CUDA_TRY(ctx, cudaGetLastError());
            LSX_CUDA_TRY(ctx, cudaMemcpyAsync(h_npiv.data(), npiv, (size_t)g * 4, cudaMemcpyDeviceToHost, ctx->stream));
            LSX_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            for (int i = 0; i < g; ++i) best = std::max(best, (int)h_npiv[i]);
            used += g;
        }
    }
    *rank = best;
    if (primes_used) *primes_used = used;
    return LSX_OK;
}

int lsx_crt_signed(lsx_ctx* ctx, const uint32_t* residues, int count, int limbs, int mem, uint32_t* out) {
    if (!ctx) return LSX_ERR_NULL;
    if (!residues || !out) return lsx_fail(ctx, LSX_ERR_NULL, "crt: NULL buffer");
    if (count < 1 || count > LSX_TABLE_PRIMES || limbs < 1 || limbs > 4 * LSX_TABLE_PRIMES)
        return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "crt: bad count or limbs");
    // the Horner accumulator needs room for the full product of the primes
    const int Lfull = count + 1;
    const int L = limbs > Lfull ? limbs : Lfull;
    LSX_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = (off + bytes + 255) / 256 * 256;
        return o;
    };
    const size_t o_tab = take((size_t)count * count * 4), o_res = take((size_t)count * 4),
                 o_out = take((size_t)L * 4), o_limb = take((size_t)2 * L * 8);
    int rc = lsx_ws_reserve(ctx, off);
    if (rc != LSX_OK) return rc;
    char* base = (char*)ctx->d_ws;
    uint32_t* d_res = (uint32_t*)(base + o_res);
    LSX_CUDA_TRY(ctx, cudaMemcpyAsync(d_res, residues, (size_t)count * 4,
                                      mem == LSX_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                                      ctx->stream));
    const int64_t nt = (int64_t)count * count;
    k_inv_table<<<(unsigned)((nt + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_primes, count, (uint32_t*)(base + o_tab));
    ctx->launches++;
    k_garner_big<<<1, 1024, (size_t)count * 4, ctx->stream>>>(ctx->d_primes, (const uint32_t*)(base + o_tab), d_res, count,
                                                             L, (uint32_t*)(base + o_out), (uint64_t*)(base + o_limb));
    ctx->launches++;
    LSX_CUDA_TRY(ctx, cudaGetLastError());
    // out gets `limbs` words: the low limbs of the (sign-extended) L-limb value
    if (limbs <= L) {
        LSX_CUDA_TRY(ctx, cudaMemcpyAsync(out, base + o_out, (size_t)limbs * 4,
                                          mem == LSX_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice,
                                          ctx->stream));
    }
    if (mem == LSX_MEM_HOST) LSX_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LSX_OK;
}

}  // extern "C"
