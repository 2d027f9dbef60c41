// Blocked right-looking modular LU for ONE large integer matrix and MANY primes at once (config 5 of
// BASELINE.json: a 4096 x 4096 determinant, primes sharded over the GPUs).
//
// det(A) mod p = sign * product of the pivots of the forward sweep of reference linalg.py:547-609
// (any non-zero pivot gives the same determinant; the first non-zero at or below the diagonal is
// used, like the reference).  For a group of G primes the residue matrices W_g = A mod p_g live side
// by side in HBM ([G][n][n] words) and every kernel below works on all of them in one launch.
//
// Recursive blocking (host side, lu() / trsm() below):
//   outer blocks of 256 columns; inside a block the panel is halved down to 8-column base panels
//   k_load      A -> residues
//   k_lu8       base panel: one CTA per prime, the rows x 8 panel lives in REGISTERS (row per thread slot),
//               pivot = first non-zero row at or below the diagonal (block-wide min), multipliers are stored
//               as (p - l) * R mod p ("negated Montgomery form") so that every later update is
//               redc((w << 32) + lneg * u) = w - l * u  with ONE reduction
//   k_swap      row swaps of a pivot range applied to a column range (left L columns and right columns)
//   k_trsm      unit-lower triangular solve with a <= 64-wide diagonal block held in shared memory
//   gemm()      C -= L * U on a rectangular region: contraction depth 64 / 128 / 256 -> tcgen05 int8-split
//               tensor-core kernel (lsx_tc.cuh: stacked byte planes, 7 TMEM accumulators x 2 tiles, cp.async.bulk staging);
//               smaller depths (8 / 16 / 32, inside a 64-column panel) -> k_gemm_int on the integer pipe
//               (64-bit accumulators, lazy high-word reduction, one REDC at the end)
//   k_finish    sign, zero flag, Montgomery -> plain residue
#include <algorithm>
#include <atomic>
#include <thread>

#include "lsx_internal.h"
#include "lsx_tc.cuh"

namespace {

constexpr int NB_OUT = 256;       // outer block (tensor-core contraction depth of the trailing update)
constexpr int NB_BASE = 8;        // register-resident base panel
constexpr int LU8_T = 512;        // threads of the base-panel CTA (1024 x 4 rows measured slower: 55 vs 47 us per launch)
constexpr int LU8_RPT = 8;        // rows per thread kept in registers -> 4096 rows
constexpr int PANEL_T = 1024;     // threads of the global-memory fallback panel (more than 4096 rows)
constexpr int GM = 128, GN = 64;  // k_gemm_int tile per CTA (256 threads, 8 x 4 outputs each)
constexpr int KI = 64;            // largest contraction depth of k_gemm_int / k_trsm

__device__ __forceinline__ uint64_t mac_lazy(uint64_t acc, uint32_t a, uint32_t b, uint32_t p) {
    acc += (uint64_t)a * b;                       // acc < p*2^32 before, product < 2^62: no overflow
    uint32_t hi = (uint32_t)(acc >> 32);
    hi = min(hi, hi - p);                         // subtract p*2^32 once if possible: keeps acc < p*2^32
    return ((uint64_t)hi << 32) | (uint32_t)acc;
}

struct LargeArgs {
    uint32_t* W;          // [G][n][n]
    const PrimeRec* primes;   // [G] (already offset to the group's first prime)
    int32_t* piv_row;     // [G][n]
    uint32_t* detM;       // [G] running product of pivots (Montgomery form)
    int32_t* flags;       // [G] bit0: odd number of swaps, bit1: zero determinant
    int n, G;
};

// any int32 into [0, p) for p > 2^30: |v| <= 2^31 < 2p, two conditional corrections, no division
__device__ __forceinline__ uint32_t residue_of(int32_t v, uint32_t p) {
    int64_t t = v;
    if (t < 0) t += p;
    if (t < 0) t += p;
    if (t >= (int64_t)p) t -= p;
    return (uint32_t)t;
}
__global__ void k_load(const int32_t* __restrict__ A, LargeArgs a) {
    const int64_t nn = (int64_t)a.n * a.n;
    const int g = blockIdx.y;
    const uint32_t p = a.primes[g].p;
    uint32_t* Wg = a.W + (int64_t)g * nn;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p > (1u << 30) && (nn & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0) {   // 16-byte accesses
        const int4* A4 = reinterpret_cast<const int4*>(A);
        uint4* W4 = reinterpret_cast<uint4*>(Wg);
        for (int64_t i = t0; i < (nn >> 2); i += stride) {
            const int4 v = __ldg(A4 + i);
            W4[i] = make_uint4(residue_of(v.x, p), residue_of(v.y, p), residue_of(v.z, p), residue_of(v.w, p));
        }
    } else {
        for (int64_t i = t0; i < nn; i += stride)
            Wg[i] = p > (1u << 30) ? residue_of(A[i], p) : word_of_int_any(A[i], p);      // tiny test primes
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.detM[g] = a.primes[g].one;
        a.flags[g] = 0;
    }
}

// c = piv^-1 * R^2 as a word, so that mont_mul(w, c) = (w / piv) * R; pivM = word of piv * R
__device__ __forceinline__ uint32_t pivot_scale(uint32_t piv, const PrimeRec& P, uint32_t* pivM_out) {
    const uint32_t pivM = mont_mul(piv, P.r2, P.p, P.pinv);
    // Fermat inverse, deliberately NOT unrolled: the panel kernels inline this once per column and their code
    // must stay inside the instruction cache
    uint32_t invM = P.one;
    const uint32_t e = P.p - 2u;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
        invM = mont_mul(invM, invM, P.p, P.pinv);
        if ((e >> bit) & 1u) invM = mont_mul(invM, pivM, P.p, P.pinv);
    }
    *pivM_out = pivM;
    return mont_mul(invM, P.r2, P.p, P.pinv);
}

// ---- base panel in registers: rows [k0, n) x columns [k0, k0 + nb), nb <= 8, n - k0 <= LU8_T * LU8_RPT ----
// The eight pivot steps run FRACTION-FREE (row <- piv * row - f * pivot_row, no division), so the Fermat inversion
// -- a serial chain of ~60 Montgomery products that every warp would wait for -- is needed once per panel instead
// of once per column: at step j every row from j on carries the common factor S_j = piv'_0 ... piv'_{j-1}, the true
// multiplier is l = f / piv'_j (same step, same factor), the true pivot row is row'_j / S_j and the true pivot is
// piv'_j / S_j.  One inversion of the product of all piv' gives every 1 / piv'_j and 1 / S_j (Montgomery's trick);
// a final pass rescales each register once.
// Pivot search, the common case first: the diagonal entry itself is the first non-zero at or below the diagonal
// unless it is zero (probability 1 / p once the entries are residues of an eliminated matrix), so the owner of row j
// publishes its row together with a "diagonal is non-zero" flag and the block-wide search (two shuffle reductions, two
// more barriers, the exchange through shared memory) only runs when the flag says it must.  The published row is
// double-buffered by the parity of the column, which also removes the barrier at the end of a column: ONE barrier per
// column on the common path instead of three.
// After the panel the same CTA applies its (rare) row swaps to the columns [col_left, k0) left of the panel -- the multipliers
// of the earlier panels of the outer block -- which used to be a k_swap launch of its own after every base panel.
// fuse_prev: the previous launch factored the full 8-column panel [k0 - 8, k0) of the same 16-column pair.  This launch
// then brings its own columns up to date first -- the row swaps of that panel, U12 = L11^-1 A12 for the rows k0 - 8 ..
// k0 - 1 (an 8 x 8 unit-lower solve in shared memory), and A22 -= L21 U12 on the register-resident rows (64 multiply-adds
// per row, U12 as broadcast shared loads) -- which used to be a k_trsm32 and a k_gemm_narrow launch between the two base
// panels (8 + 18 us of mostly launch and load latency per pair).
__global__ void __launch_bounds__(LU8_T) k_lu8(LargeArgs a, int k0, int nb, int col_left, int fuse_prev) {
    __shared__ uint32_t rowbuf[2][2][NB_BASE];   // [column parity][0: old row j, 1: pivot row (old row src)]
    __shared__ int s_nz[2];                      // [column parity] the diagonal entry is non-zero
    __shared__ int red[LU8_T / 32];
    __shared__ __align__(16) uint32_t s_u[NB_BASE][NB_BASE];     // fuse_prev: A12, then U12 ([k][c])
    __shared__ uint32_t s_l11[NB_BASE][NB_BASE];                 // fuse_prev: multipliers of the previous diagonal block
    __shared__ int s_prev_piv[NB_BASE];
    const int n = a.n, g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const PrimeRec P = a.primes[g];
    const uint32_t p = P.p, pinv = P.pinv;
    uint32_t* Wg = a.W + (int64_t)g * n * n;
    const bool vec = nb == NB_BASE && (n & 3) == 0 && (k0 & 3) == 0;
    if (fuse_prev) {
        const int kp = k0 - NB_BASE;
        if (tid < NB_BASE) s_prev_piv[tid] = a.piv_row[(int64_t)g * n + kp + tid];
        __syncthreads();
        bool any_swap = false;
#pragma unroll
        for (int i = 0; i < NB_BASE; ++i) any_swap |= s_prev_piv[i] != kp + i;
        if (any_swap) {                                           // uniform; rare
            if (tid < nb) {
                const int c = k0 + tid;
                for (int i = 0; i < NB_BASE; ++i) {
                    const int j = kp + i, src = s_prev_piv[i];
                    if (src != j) {
                        const uint32_t t0 = Wg[(int64_t)j * n + c];
                        Wg[(int64_t)j * n + c] = Wg[(int64_t)src * n + c];
                        Wg[(int64_t)src * n + c] = t0;
                    }
                }
            }
            __syncthreads();                                      // the swapped words are read below by other threads
        }
        if (tid < NB_BASE * NB_BASE) {
            const int i = tid / NB_BASE, t = tid % NB_BASE;
            s_l11[i][t] = t < i ? Wg[(int64_t)(kp + i) * n + kp + t] : 0u;
            s_u[i][t] = t < nb ? Wg[(int64_t)(kp + i) * n + k0 + t] : 0u;
        }
        __syncthreads();
        if (tid < nb) {                                           // column tid of U12 (same arithmetic as k_trsm32)
            uint32_t u[NB_BASE];
            u[0] = s_u[0][tid];
#pragma unroll
            for (int i = 1; i < NB_BASE; ++i) {
                uint64_t acc = (uint64_t)s_u[i][tid] << 32;
#pragma unroll
                for (int t = 0; t < i; ++t) acc = mac_lazy(acc, s_l11[i][t], u[t], p);
                u[i] = mont_redc(acc, p, pinv);
                s_u[i][tid] = u[i];
                Wg[(int64_t)(kp + i) * n + k0 + tid] = u[i];
            }
        }
        __syncthreads();
    }
    uint32_t v[LU8_RPT][NB_BASE];
#pragma unroll
    for (int i = 0; i < LU8_RPT; ++i) {
        const int r = k0 + tid + i * LU8_T;
        if (r < n) {
            const uint32_t* src = Wg + (int64_t)r * n + k0;
            if (vec) {
                const uint4 x = *reinterpret_cast<const uint4*>(src), y = *reinterpret_cast<const uint4*>(src + 4);
                v[i][0] = x.x, v[i][1] = x.y, v[i][2] = x.z, v[i][3] = x.w;
                v[i][4] = y.x, v[i][5] = y.y, v[i][6] = y.z, v[i][7] = y.w;
            } else {
#pragma unroll
                for (int c = 0; c < NB_BASE; ++c) v[i][c] = c < nb ? src[c] : 0u;
            }
        } else {
#pragma unroll
            for (int c = 0; c < NB_BASE; ++c) v[i][c] = 0u;
        }
    }
    if (fuse_prev) {
        // A22 -= L21 U12: the row's eight multipliers of the previous panel come from global memory (just written: L2)
        const int kp = k0 - NB_BASE;
#pragma unroll
        for (int i = 0; i < LU8_RPT; ++i) {
            const int r = k0 + tid + i * LU8_T;
            if (r < n) {
                const uint32_t* lsrc = Wg + (int64_t)r * n + kp;
                uint32_t l[NB_BASE];
                if ((n & 3) == 0 && (kp & 3) == 0) {
                    const uint4 x = *reinterpret_cast<const uint4*>(lsrc), y = *reinterpret_cast<const uint4*>(lsrc + 4);
                    l[0] = x.x, l[1] = x.y, l[2] = x.z, l[3] = x.w, l[4] = y.x, l[5] = y.y, l[6] = y.z, l[7] = y.w;
                } else {
#pragma unroll
                    for (int k = 0; k < NB_BASE; ++k) l[k] = lsrc[k];
                }
                uint64_t acc[NB_BASE];
#pragma unroll
                for (int c = 0; c < NB_BASE; ++c) acc[c] = (uint64_t)v[i][c] << 32;
#pragma unroll
                for (int k = 0; k < NB_BASE; ++k) {
                    const uint4 ua = *reinterpret_cast<const uint4*>(&s_u[k][0]), ub = *reinterpret_cast<const uint4*>(&s_u[k][4]);
                    const uint32_t uk[NB_BASE] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
#pragma unroll
                    for (int c = 0; c < NB_BASE; ++c) acc[c] = mac_lazy(acc[c], l[k], uk[c], p);
                }
#pragma unroll
                for (int c = 0; c < NB_BASE; ++c) v[i][c] = mont_redc(acc[c], p, pinv);
            }
        }
    }
    int flags = a.flags[g];
    uint32_t pivw[NB_BASE];                      // piv'_j as Montgomery words (R for a column without pivot)
    int srcs[NB_BASE];                           // pivot rows of the panel (uniform over the CTA)
#pragma unroll
    for (int jj = 0; jj < NB_BASE; ++jj) {
        pivw[jj] = P.one;
        srcs[jj] = k0 + jj;
        if (jj < nb) {                                            // uniform
            const int j = k0 + jj, par = jj & 1;
            // ---- common case: row j (thread jj, slot 0) is the pivot row ----
            if (tid == jj) {
#pragma unroll
                for (int c = 0; c < NB_BASE; ++c) rowbuf[par][1][c] = v[0][c];
                s_nz[par] = v[0][jj] != 0u;
            }
            __syncthreads();
            int src = j;
            bool zero_col = false;
            if (!s_nz[par]) {                                     // uniform; rare
                // ---- pivot search: first row >= j with a non-zero entry in column jj ----
                int best = INT32_MAX;
#pragma unroll
                for (int i = LU8_RPT - 1; i >= 0; --i) {
                    const int r = k0 + tid + i * LU8_T;
                    if (r >= j && r < n && v[i][jj] != 0u) best = r;
                }
                for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
                if (lane == 0) red[warp] = best;
                __syncthreads();
                best = lane < LU8_T / 32 ? red[lane] : INT32_MAX;
                for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
                src = best;
                zero_col = src == INT32_MAX;
                if (zero_col) src = j;                            // column is zero below the diagonal: det = 0
                // ---- exchange rows j and src through shared memory; everyone then reads the pivot row ----
#pragma unroll
                for (int i = 0; i < LU8_RPT; ++i) {
                    const int r = k0 + tid + i * LU8_T;
                    if (r == j) {
#pragma unroll
                        for (int c = 0; c < NB_BASE; ++c) rowbuf[par][0][c] = v[i][c];
                    }
                    if (r == src) {
#pragma unroll
                        for (int c = 0; c < NB_BASE; ++c) rowbuf[par][1][c] = v[i][c];
                    }
                }
                __syncthreads();
                if (src != j) {
#pragma unroll
                    for (int i = 0; i < LU8_RPT; ++i) {
                        const int r = k0 + tid + i * LU8_T;
                        if (r == j) {
#pragma unroll
                            for (int c = 0; c < NB_BASE; ++c) v[i][c] = rowbuf[par][1][c];
                        }
                        if (r == src) {
#pragma unroll
                            for (int c = 0; c < NB_BASE; ++c) v[i][c] = rowbuf[par][0][c];
                        }
                    }
                }
            }
            srcs[jj] = src;
            uint32_t prow[NB_BASE];
#pragma unroll
            for (int c = 0; c < NB_BASE; ++c) prow[c] = rowbuf[par][1][c];
            if (tid == 0) {
                a.piv_row[(int64_t)g * n + j] = src;
                if (zero_col) flags |= 2;
                if (src != j) flags ^= 1;
            }
            // fraction-free step: row <- piv * row - f * pivot_row for the rows below (a zero column changes nothing)
            const uint32_t pm = zero_col ? P.one : mont_mul(prow[jj], P.r2, p, pinv);      // word of piv'
            pivw[jj] = pm;
#pragma unroll
            for (int i = 0; i < LU8_RPT; ++i) {
                const int r = k0 + tid + i * LU8_T;
                if (r > j && r < n) {
                    const uint32_t f = v[i][jj];
                    const uint32_t nf = mont_mul(f ? p - f : 0u, P.r2, p, pinv);            // word of -f
#pragma unroll
                    for (int c = jj + 1; c < NB_BASE; ++c) v[i][c] = mont_fma2(pm, v[i][c], nf, prow[c], p, pinv);
                }
            }
            // no barrier here: the next column publishes into the other half of rowbuf / s_nz, and red is only written
            // behind that column's publish barrier
        }
    }
    // ---- one inversion for the whole panel: ip[j] = word of 1 / (piv'_0 ... piv'_j) ----
    uint32_t pre[NB_BASE];                       // pre[j] = word of piv'_0 ... piv'_j
    pre[0] = pivw[0];
#pragma unroll
    for (int j = 1; j < NB_BASE; ++j) pre[j] = mont_mul(pre[j - 1], pivw[j], p, pinv);
    uint32_t inv = P.one;                        // Fermat, not unrolled (code size)
    {
        const uint32_t e = p - 2u, base = pre[NB_BASE - 1];
#pragma unroll 1
        for (int bit = 31; bit >= 0; --bit) {
            inv = mont_mul(inv, inv, p, pinv);
            if ((e >> bit) & 1u) inv = mont_mul(inv, base, p, pinv);
        }
    }
    // cl[j] = (1 / piv'_j) * R^2  (multiplier of column j:  mont_mul(p - f, cl[j]) = (-f / piv'_j) * R)
    // cu[j] = (1 / S_j) * R       (pivot row j and everything right of its pivot: mont_mul(v, cu[j]) = v / S_j)
    uint32_t cl[NB_BASE], cu[NB_BASE];
    uint32_t detM = a.detM[g];
#pragma unroll
    for (int j = NB_BASE - 1; j >= 0; --j) {
        // inv = word of 1 / pre[j]
        const uint32_t sinv = j ? mont_mul(inv, pivw[j], p, pinv) : P.one;          // word of 1 / pre[j-1] = 1 / S_j
        const uint32_t pinvw = j ? mont_mul(inv, pre[j - 1], p, pinv) : inv;        // word of 1 / piv'_j
        cl[j] = mont_mul(pinvw, P.r2, p, pinv);
        cu[j] = sinv;
        if (j < nb && tid == 0 && !(flags & 2)) detM = mont_mul(detM, mont_mul(pivw[j], sinv, p, pinv), p, pinv);
        inv = sinv;
    }
    // ---- rescale: column c of row k0 + q is a multiplier if q > c, else part of pivot row q (or of a row above) ----
#pragma unroll
    for (int i = 0; i < LU8_RPT; ++i) {
        const int q = tid + i * LU8_T;           // row index inside the panel's row range
#pragma unroll
        for (int c = 0; c < NB_BASE; ++c) {
            if (c < nb) {
                if (q > c) {
                    const uint32_t f = v[i][c];
                    v[i][c] = mont_mul(f ? p - f : 0u, cl[c], p, pinv);
                } else {
                    uint32_t cs = cu[0];
#pragma unroll
                    for (int t = 1; t < NB_BASE; ++t)
                        if (t == q) cs = cu[t];
                    v[i][c] = mont_mul(v[i][c], cs, p, pinv);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < LU8_RPT; ++i) {
        const int r = k0 + tid + i * LU8_T;
        if (r < n) {
            uint32_t* dst = Wg + (int64_t)r * n + k0;
            if (vec) {
                *reinterpret_cast<uint4*>(dst) = make_uint4(v[i][0], v[i][1], v[i][2], v[i][3]);
                *reinterpret_cast<uint4*>(dst + 4) = make_uint4(v[i][4], v[i][5], v[i][6], v[i][7]);
            } else {
#pragma unroll
                for (int c = 0; c < NB_BASE; ++c)
                    if (c < nb) dst[c] = v[i][c];
            }
        }
    }
    if (tid == 0) {
        a.detM[g] = detM;
        a.flags[g] = flags;
    }
    // ---- the panel's row swaps on the columns [col_left, k0) (at most NB_OUT - NB_BASE of them: one per thread) ----
    bool any_swap = false;
#pragma unroll
    for (int jj = 0; jj < NB_BASE; ++jj) any_swap |= srcs[jj] != k0 + jj;
    if (any_swap) {                                               // uniform; rare
        for (int c = col_left + tid; c < k0; c += LU8_T) {
#pragma unroll
            for (int jj = 0; jj < NB_BASE; ++jj) {
                const int j = k0 + jj, src = srcs[jj];
                if (jj < nb && src != j) {
                    const uint32_t t0 = Wg[(int64_t)j * n + c];
                    Wg[(int64_t)j * n + c] = Wg[(int64_t)src * n + c];
                    Wg[(int64_t)src * n + c] = t0;
                }
            }
        }
    }
}

// ---- fallback base panel in global memory (more rows than the register panel holds) ----
__global__ void __launch_bounds__(PANEL_T) k_panel_gmem(LargeArgs a, int k0, int nb) {
    __shared__ uint32_t prow[NB_BASE];
    __shared__ int red[PANEL_T / 32];
    __shared__ uint32_t s_c;
    const int n = a.n, g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = PANEL_T / 32;
    const PrimeRec P = a.primes[g];
    const uint32_t p = P.p, pinv = P.pinv;
    uint32_t* Wg = a.W + (int64_t)g * n * n;
    uint32_t detM = a.detM[g];
    int flags = a.flags[g];
    for (int jj = 0; jj < nb; ++jj) {
        const int j = k0 + jj;
        int best = INT32_MAX;
        for (int r = j + tid; r < n; r += PANEL_T)
            if (Wg[(int64_t)r * n + j] != 0u) {
                best = r;
                break;
            }
        for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        if (lane == 0) red[warp] = best;
        __syncthreads();
        best = red[lane];                                  // NW == 32
        for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        int src = best;
        if (src == INT32_MAX) {
            flags |= 2;
            src = j;
        }
        if (src != j) {
            flags ^= 1;
            if (tid < nb) {
                const uint32_t t0 = Wg[(int64_t)j * n + k0 + tid];
                Wg[(int64_t)j * n + k0 + tid] = Wg[(int64_t)src * n + k0 + tid];
                Wg[(int64_t)src * n + k0 + tid] = t0;
            }
        }
        if (tid == 0) a.piv_row[(int64_t)g * n + j] = src;
        __syncthreads();
        if (tid < nb) prow[tid] = Wg[(int64_t)j * n + k0 + tid];
        if (warp == 0) {
            uint32_t pivM;
            const uint32_t c = pivot_scale(Wg[(int64_t)j * n + j], P, &pivM);
            if (lane == 0) s_c = c;
            if (tid == 0 && !(flags & 2)) detM = mont_mul(detM, pivM, p, pinv);
        }
        __syncthreads();
        const uint32_t c = s_c;
        for (int r = j + 1 + warp; r < n; r += NW) {
            uint32_t* row = Wg + (int64_t)r * n;
            uint32_t w = lane == 0 ? row[j] : 0u;
            w = __shfl_sync(0xffffffffu, w, 0);
            const uint32_t lm = mont_mul(w, c, p, pinv);
            const uint32_t ln = lm ? p - lm : 0u;
            for (int cc = jj + 1 + lane; cc < nb; cc += 32)
                row[k0 + cc] = mont_redc(mac_lazy((uint64_t)row[k0 + cc] << 32, ln, prow[cc], p), p, pinv);
            if (lane == 0) row[j] = ln;
        }
        __syncthreads();
    }
    if (tid == 0) {
        a.detM[g] = detM;
        a.flags[g] = flags;
    }
}

// Row swaps of pivots [ja, jb) applied, in order, to columns [ca, cb).
__global__ void __launch_bounds__(128) k_swap(LargeArgs a, int ja, int jb, int ca, int cb) {
    __shared__ int pr[NB_OUT];
    const int n = a.n, g = blockIdx.y, tid = threadIdx.x;
    uint32_t* Wg = a.W + (int64_t)g * n * n;
    const int c = ca + blockIdx.x * 128 + tid;
    for (int jbase = ja; jbase < jb; jbase += NB_OUT) {
        const int cnt = min(NB_OUT, jb - jbase);
        __syncthreads();
        for (int e = tid; e < cnt; e += 128) pr[e] = a.piv_row[(int64_t)g * n + jbase + e];
        __syncthreads();
        if (c < cb)
            for (int e = 0; e < cnt; ++e) {
                const int j = jbase + e, src = pr[e];
                if (src != j) {
                    const uint32_t t0 = Wg[(int64_t)j * n + c];
                    Wg[(int64_t)j * n + c] = Wg[(int64_t)src * n + c];
                    Wg[(int64_t)src * n + c] = t0;
                }
            }
    }
}

// Columns [ca, cb): solve the unit-lower system of the diagonal block rows/cols [k0, k0 + nb), nb <= 64.
__global__ void __launch_bounds__(128) k_trsm(LargeArgs a, int k0, int nb, int ca, int cb) {
    extern __shared__ __align__(16) uint32_t sm[];
    uint32_t* Ln = sm;                    // [nb][KI] negated Montgomery multipliers of L11
    uint32_t* us = Ln + KI * KI;          // [nb][128]
    const int n = a.n, g = blockIdx.y, tid = threadIdx.x;
    const uint32_t p = a.primes[g].p, pinv = a.primes[g].pinv;
    uint32_t* Wg = a.W + (int64_t)g * n * n;
    for (int e = tid; e < nb * nb; e += 128) {
        const int i = e / nb, t = e % nb;
        Ln[i * KI + t] = Wg[(int64_t)(k0 + i) * n + k0 + t];
    }
    __syncthreads();
    const int c = ca + blockIdx.x * 128 + tid;
    if (c >= cb) return;
    for (int i = 0; i < nb; ++i) {
        uint64_t acc = (uint64_t)Wg[(int64_t)(k0 + i) * n + c] << 32;
        for (int t = 0; t < i; ++t) acc = mac_lazy(acc, Ln[i * KI + t], us[t * 128 + tid], p);
        const uint32_t u = mont_redc(acc, p, pinv);
        us[i * 128 + tid] = u;
        Wg[(int64_t)(k0 + i) * n + c] = u;
    }
}

// Same solve for nb <= TSZ (32 or 64) with the column in REGISTERS: all loads are issued up front, the multipliers
// come from shared memory as broadcast 16-byte loads (4 multiply-adds per load), fully unrolled.  Every row's sum runs
// on TWO accumulators (alternate groups of four terms) that are merged before the reduction: one accumulator is a
// dependent chain of up to TSZ multiply-adds, and with 128 registers per thread there are not enough resident warps
// to hide it (ncu: the 32-wide solve of 3840 columns x 127 primes took 75 us for 28 us of issue slots).
// The 64-wide form replaces the recursion  solve 32 / tensor update of 32 rows / solve 32:  a tensor-core update with 32
// live rows of a 128-row tile costs as much as a full tile (170 us per 3840 columns x 127 primes, plus its plane split).
// with_swap: the row swaps of the pivots [k0, k0 + nb) are applied to the columns first (every thread on its own
// column; a swap is rare, so this replaces a k_swap launch that did nothing most of the time).
constexpr int TS = 32;
constexpr int TS2 = 64;
template <int TSZ>
__global__ void __launch_bounds__(128, TSZ > 32 ? 2 : 4) k_trsm_reg(LargeArgs a, int k0, int nb, int ca, int cb, int with_swap) {
    __shared__ __align__(16) uint32_t Ln[TSZ][TSZ];
    __shared__ int s_piv[TSZ];
    const int n = a.n, g = blockIdx.y, tid = threadIdx.x;
    const uint32_t p = a.primes[g].p, pinv = a.primes[g].pinv;
    uint32_t* Wg = a.W + (int64_t)g * n * n;
    for (int e = tid; e < TSZ * TSZ; e += 128) {
        const int i = e / TSZ, t = e % TSZ;
        Ln[i][t] = (i < nb && t < i) ? Wg[(int64_t)(k0 + i) * n + k0 + t] : 0u;
    }
    if (tid < TSZ) s_piv[tid] = (with_swap && tid < nb) ? a.piv_row[(int64_t)g * n + k0 + tid] : k0 + tid;
    __syncthreads();
    const int c = ca + blockIdx.x * 128 + tid;
    if (c >= cb) return;
    if (with_swap) {
        bool any_swap = false;
        for (int i = 0; i < nb; ++i) any_swap |= s_piv[i] != k0 + i;
        if (any_swap) {                                           // uniform over the CTA; rare
            for (int i = 0; i < nb; ++i) {
                const int j = k0 + i, src = s_piv[i];
                if (src != j) {
                    const uint32_t t0 = Wg[(int64_t)j * n + c];
                    Wg[(int64_t)j * n + c] = Wg[(int64_t)src * n + c];
                    Wg[(int64_t)src * n + c] = t0;
                }
            }
        }
    }
    uint32_t u[TSZ];
#pragma unroll
    for (int i = 0; i < TSZ; ++i) u[i] = i < nb ? Wg[(int64_t)(k0 + i) * n + c] : 0u;
#pragma unroll
    for (int i = 1; i < TSZ; ++i) {
        if (i >= nb) break;                       // uniform: most calls solve an 8- or 16-wide block
        uint64_t acc0 = (uint64_t)u[i] << 32, acc1 = 0ull;
#pragma unroll
        for (int t4 = 0; t4 < i; t4 += 4) {
            const uint4 l = *reinterpret_cast<const uint4*>(&Ln[i][t4]);     // entries at t >= i are zero
            uint64_t& acc = (t4 & 4) ? acc1 : acc0;
            acc = mac_lazy(acc, l.x, u[t4], p);
            if (t4 + 1 < i) acc = mac_lazy(acc, l.y, u[t4 + 1], p);
            if (t4 + 2 < i) acc = mac_lazy(acc, l.z, u[t4 + 2], p);
            if (t4 + 3 < i) acc = mac_lazy(acc, l.w, u[t4 + 3], p);
        }
        if (i > 4) {                              // merge: both below p * 2^32 < 2^63, one conditional subtraction
            acc0 += acc1;
            uint32_t hi = (uint32_t)(acc0 >> 32);
            hi = min(hi, hi - p);
            acc0 = ((uint64_t)hi << 32) | (uint32_t)acc0;
        }
        u[i] = mont_redc(acc0, p, pinv);
    }
#pragma unroll
    for (int i = 1; i < TSZ; ++i)
        if (i < nb) Wg[(int64_t)(k0 + i) * n + c] = u[i];
}

// C[i][j] = redc((C[i][j] << 32) + sum_k Lneg[i][k] * U[k][j])  on rows [r0, r1) x cols [c0, c1),
// Lneg = W[.][k0 .. k0 + K), U = W[k0 .. k0 + K)[.], K <= 64: integer pipe.
__global__ void __launch_bounds__(256) k_gemm_int(LargeArgs a, int r0, int r1, int c0, int c1, int k0, int K) {
    extern __shared__ __align__(16) uint32_t sm[];
    constexpr int LDL = GM + 4;
    uint32_t* Lt = sm;                    // [K][LDL]  (k-major: Lt[k][i])
    uint32_t* Us = Lt + KI * LDL;         // [K][GN]
    const int n = a.n, g = blockIdx.z, tid = threadIdx.x;
    const uint32_t p = a.primes[g].p, pinv = a.primes[g].pinv;
    uint32_t* Wg = a.W + (int64_t)g * n * n;
    const int rb = r0 + blockIdx.y * GM, cbase = c0 + blockIdx.x * GN;
    for (int e = tid; e < GM * K; e += 256) {
        const int i = e / K, k = e % K;
        const int r = rb + i;
        Lt[k * LDL + i] = r < r1 ? Wg[(int64_t)r * n + k0 + k] : 0u;
    }
    for (int e = tid; e < K * GN; e += 256) {
        const int k = e / GN, j = e % GN;
        const int c = cbase + j;
        Us[k * GN + j] = c < c1 ? Wg[(int64_t)(k0 + k) * n + c] : 0u;
    }
    __syncthreads();
    const int tx = tid & 15, ty = tid >> 4;           // 16 x 16 threads: 4 columns x 8 rows each
    uint64_t acc[8][4];
#pragma unroll
    for (int x = 0; x < 8; ++x) {
        const int r = rb + ty * 8 + x;
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int c = cbase + tx * 4 + y;
            acc[x][y] = (r < r1 && c < c1) ? (uint64_t)Wg[(int64_t)r * n + c] << 32 : 0ull;
        }
    }
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const uint4 l0 = *reinterpret_cast<const uint4*>(Lt + k * LDL + ty * 8);
        const uint4 l1 = *reinterpret_cast<const uint4*>(Lt + k * LDL + ty * 8 + 4);
        const uint4 u = *reinterpret_cast<const uint4*>(Us + k * GN + tx * 4);
        const uint32_t lv[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
        const uint32_t uv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int x = 0; x < 8; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y] = mac_lazy(acc[x][y], lv[x], uv[y], p);
    }
#pragma unroll
    for (int x = 0; x < 8; ++x) {
        const int r = rb + ty * 8 + x;
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int c = cbase + tx * 4 + y;
            if (r < r1 && c < c1) Wg[(int64_t)r * n + c] = mont_redc(acc[x][y], p, pinv);
        }
    }
}

// Same update for a tall, narrow region (inside a 32-column panel: K <= 16, at most 16 columns): one thread
// per row keeps its L entries and its C entries in registers, U (K x width) sits in shared memory.
constexpr int NARROW = 16;
__global__ void __launch_bounds__(128) k_gemm_narrow(LargeArgs a, int r0, int r1, int c0, int c1, int k0, int K) {
    __shared__ uint32_t Us[NARROW][NARROW];
    const int n = a.n, g = blockIdx.y, tid = threadIdx.x, wd = c1 - c0;
    const uint32_t p = a.primes[g].p, pinv = a.primes[g].pinv;
    uint32_t* Wg = a.W + (int64_t)g * n * n;
    for (int e = tid; e < NARROW * NARROW; e += 128) {
        const int k = e / NARROW, j = e % NARROW;
        Us[k][j] = (k < K && j < wd) ? Wg[(int64_t)(k0 + k) * n + c0 + j] : 0u;
    }
    __syncthreads();
    const int r = r0 + blockIdx.x * 128 + tid;
    if (r >= r1) return;
    uint32_t* row = Wg + (int64_t)r * n;
    uint32_t l[NARROW];
    uint64_t acc[NARROW];
    const bool vec = (n & 3) == 0 && (k0 & 3) == 0 && (c0 & 3) == 0 && (K & 3) == 0 && (wd & 3) == 0;
    if (vec) {
#pragma unroll
        for (int k = 0; k < NARROW; k += 4)
            if (k < K) {
                const uint4 v = *reinterpret_cast<const uint4*>(row + k0 + k);
                l[k] = v.x, l[k + 1] = v.y, l[k + 2] = v.z, l[k + 3] = v.w;
            } else {
                l[k] = l[k + 1] = l[k + 2] = l[k + 3] = 0u;
            }
#pragma unroll
        for (int j = 0; j < NARROW; j += 4)
            if (j < wd) {
                const uint4 v = *reinterpret_cast<const uint4*>(row + c0 + j);
                acc[j] = (uint64_t)v.x << 32, acc[j + 1] = (uint64_t)v.y << 32, acc[j + 2] = (uint64_t)v.z << 32,
                acc[j + 3] = (uint64_t)v.w << 32;
            } else {
                acc[j] = acc[j + 1] = acc[j + 2] = acc[j + 3] = 0ull;
            }
    } else {
#pragma unroll
        for (int k = 0; k < NARROW; ++k) l[k] = k < K ? row[k0 + k] : 0u;
#pragma unroll
        for (int j = 0; j < NARROW; ++j) acc[j] = j < wd ? (uint64_t)row[c0 + j] << 32 : 0ull;
    }
#pragma unroll
    for (int k = 0; k < NARROW; ++k)
        if (k < K) {
#pragma unroll
            for (int j = 0; j < NARROW; ++j) acc[j] = mac_lazy(acc[j], l[k], Us[k][j], p);
        }
    if (vec) {
#pragma unroll
        for (int j = 0; j < NARROW; j += 4)
            if (j < wd)
                *reinterpret_cast<uint4*>(row + c0 + j) = make_uint4(mont_redc(acc[j], p, pinv), mont_redc(acc[j + 1], p, pinv),
                                                                     mont_redc(acc[j + 2], p, pinv), mont_redc(acc[j + 3], p, pinv));
    } else {
#pragma unroll
        for (int j = 0; j < NARROW; ++j)
            if (j < wd) row[c0 + j] = mont_redc(acc[j], p, pinv);
    }
}

__global__ void k_finish(LargeArgs a, uint32_t* residues) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.G) return;
    const PrimeRec P = a.primes[g];
    uint32_t x = mont_redc((uint64_t)a.detM[g], P.p, P.pinv);    // Montgomery -> plain
    const int f = a.flags[g];
    if (f & 2) x = 0u;
    else if ((f & 1) && x) x = P.p - x;
    residues[g] = x;
}

// ---- host-side recursion over one group of primes -----------------------------------------------------
struct Driver {
    lsx_ctx* ctx;
    cudaStream_t stream;  // the stream this group of primes runs on (ctx->stream or one of its side streams)
    int64_t launches;     // counted here (several drivers run from their own host threads), added to ctx at the end
    bool timing;          // event-time the depth-256 updates (single-stream runs only)
    int max_tiles;        // cap on the tiles one tensor-kernel CTA loops over (0: none); short CTAs free their SM soon
    LargeArgs a;
    uint8_t* AP;          // byte planes for the tensor-core update (lsx_tc.cuh)
    uint8_t* BP;
    bool use_tc;
    bool trsm64;          // 64-wide register solve instead of solve 32 / tensor update / solve 32 (LSX_TRSM64=0: off)
    bool fuse_pair;       // pairs of base panels without the k_trsm32 / k_gemm_narrow launches in between (LSX_LU_PAIR=0: off)
    int n;

    static int left_width(int w) {               // largest power of two below w (w > NB_BASE)
        int h = NB_BASE;
        while (h * 2 < w) h *= 2;
        return h;
    }
    void gemm(int r0, int r1, int c0, int c1, int k0, int K) {
        if (r1 <= r0 || c1 <= c0) return;
        cudaStream_t st = stream;
        if (use_tc && lsx_tc::depth_ok(K)) {
            lsx_tc::Region g{};
            g.n = n, g.r0 = r0, g.r1 = r1, g.c0 = c0, g.c1 = c1, g.k0 = k0, g.K = K;
            g.kc = std::min(K, lsx_tc::KC);
            g.set_tiles();
            // tall and narrow (inside a panel: at most 128 columns): keep the B planes of one 32-column tile
            // resident and stream the row tiles; otherwise keep the A planes of a row tile and stream columns
            g.b_stationary = (g.n_tiles <= 4 && g.m_tiles > 2) ? 1 : 0;
            const int fixed = g.b_stationary ? g.n_tiles : g.m_tiles, looped = g.b_stationary ? g.m_tiles : g.n_tiles;
            const int units = fixed * a.G;
            int groups = std::min(looped, std::max(1, (2 * ctx->sm_count + units - 1) / units));
            g.tiles_per_cta = (looped + groups - 1) / groups;
            if (max_tiles > 0) g.tiles_per_cta = std::min(g.tiles_per_cta, max_tiles);
            groups = (looped + g.tiles_per_cta - 1) / g.tiles_per_cta;
            lsx_tc::launch_split(a.W, AP, BP, g, a.G, st);
            lsx_tc::GemmArgs ga{};
            ga.W = a.W, ga.AP = AP, ga.BP = BP, ga.primes = a.primes, ga.g = g;
            const bool big = timing && K == NB_OUT;
            if (big) lsx_timing_begin(ctx);
            // epilogue mapping, measured on B200 (profiles/r02aa_tc_epilogue_coalesced_vs_rowmap.jsonl): staging the update
            // through shared memory for coalesced C accesses pays at depth 128 (5.80 -> 5.25 ms per 127 primes x 3968^2),
            // where the MMAs of a tile take about as long as its epilogue and both want the L1 data pipe; at depth 256
            // the kernel is MMA bound either way (8.08 vs 8.17 ms) and at depth 32 / 64 the extra hop through shared
            // memory lengthens a latency-bound tile (69 -> 75 us).  LSX_TC_COAL = 0 / 1 forces one mapping.
            static const int env_coal = []() {
                const char* e = getenv("LSX_TC_COAL");
                return e ? atoi(e) : -1;
            }();
            const bool coal = env_coal >= 0 ? env_coal != 0 : K == 128;
            auto kern = coal ? lsx_tc::k_gemm_tc : lsx_tc::k_gemm_tc_rowmap;
            kern<<<dim3(fixed, groups, a.G), lsx_tc::THREADS, lsx_tc::smem_bytes(K, g.b_stationary), st>>>(ga);
            if (big) lsx_timing_end(ctx);
            launches += 2;
            return;
        }
        if (K <= NARROW && c1 - c0 <= NARROW) {   // tall and narrow: one thread per row
            k_gemm_narrow<<<dim3((r1 - r0 + 127) / 128, a.G), 128, 0, st>>>(a, r0, r1, c0, c1, k0, K);
            launches++;
            return;
        }
        for (int kk = 0; kk < K; kk += KI) {      // integer pipe, at most 64 deep per launch
            const int kd = std::min(KI, K - kk);
            const size_t smem = (size_t)(KI * (GM + 4) + KI * GN) * 4;
            k_gemm_int<<<dim3((c1 - c0 + GN - 1) / GN, (r1 - r0 + GM - 1) / GM, a.G), 256, smem, st>>>(a, r0, r1, c0, c1,
                                                                                                        k0 + kk, kd);
            launches++;
        }
    }
    void swap(int ja, int jb, int ca, int cb) {
        if (jb <= ja || cb <= ca) return;
        k_swap<<<dim3((cb - ca + 127) / 128, a.G), 128, 0, stream>>>(a, ja, jb, ca, cb);
        launches++;
    }
    // columns [ca, cb): U = L11^-1 * A for the diagonal block [k0, k0 + w)
    // with_swap (register solves only, w <= 64): the row swaps of the pivots [k0, k0 + w) are applied to the columns by the same launch
    void trsm(int k0, int w, int ca, int cb, bool with_swap = false) {
        if (cb <= ca || w <= 0) return;
        if (w <= TS) {
            k_trsm_reg<TS><<<dim3((cb - ca + 127) / 128, a.G), 128, 0, stream>>>(a, k0, w, ca, cb, with_swap ? 1 : 0);
            launches++;
            return;
        }
        if (w <= TS2 && trsm64) {
            k_trsm_reg<TS2><<<dim3((cb - ca + 127) / 128, a.G), 128, 0, stream>>>(a, k0, w, ca, cb, with_swap ? 1 : 0);
            launches++;
            return;
        }
        if (!use_tc && w <= KI) {
            const size_t smem = (size_t)(KI * KI + KI * 128) * 4;
            k_trsm<<<dim3((cb - ca + 127) / 128, a.G), 128, smem, stream>>>(a, k0, w, ca, cb);
            launches++;
            return;
        }
        const int w1 = left_width(w);
        trsm(k0, w1, ca, cb);
        gemm(k0 + w1, k0 + w, ca, cb, k0, w1);
        trsm(k0 + w1, w - w1, ca, cb);
    }
    // LU of rows [k0, n) x columns [k0, k0 + w); row swaps are applied to columns [cl, k0 + w) only
    // (cl = first column of the enclosing outer block: the L columns to its left are never read again).
    void lu(int k0, int w, int cl) {
        if (w <= NB_BASE) {
            if (n - k0 <= LU8_T * LU8_RPT) {
                k_lu8<<<a.G, LU8_T, 0, stream>>>(a, k0, w, cl, 0);   // swaps the multipliers of earlier panels of the block itself
                launches++;
            } else {
                k_panel_gmem<<<a.G, PANEL_T, 0, stream>>>(a, k0, w);
                launches++;
                swap(k0, k0 + w, cl, k0);                   // multipliers of earlier panels in the same block
            }
            return;
        }
        const int w1 = left_width(w);
        if (w1 == NB_BASE && fuse_pair && n - k0 <= LU8_T * LU8_RPT) {
            // a pair of base panels: the second launch does the solve and the update of its own columns itself
            k_lu8<<<a.G, LU8_T, 0, stream>>>(a, k0, w1, cl, 0);
            k_lu8<<<a.G, LU8_T, 0, stream>>>(a, k0 + w1, w - w1, cl, 1);
            launches += 2;
            return;
        }
        lu(k0, w1, cl);
        // right half of this panel: row swaps, then the solve (one launch when the diagonal block fits k_trsm32)
        if (w1 <= (trsm64 ? TS2 : TS)) {
            trsm(k0, w1, k0 + w1, k0 + w, true);
        } else {
            swap(k0, k0 + w1, k0 + w1, k0 + w);
            trsm(k0, w1, k0 + w1, k0 + w);
        }
        gemm(k0 + w1, n, k0 + w1, k0 + w, k0, w1);
        lu(k0 + w1, w - w1, cl);
    }
    void run() {
        for (int k0 = 0; k0 < n; k0 += NB_OUT) {
            const int w = std::min(NB_OUT, n - k0);
            lu(k0, w, k0);
            const int ce = k0 + w;
            if (ce < n) {
                swap(k0, ce, ce, n);
                trsm(k0, w, ce, n);
                gemm(ce, n, ce, n, k0, w);
            }
        }
    }
};

}  // namespace

// Independent groups of primes run CONCURRENTLY on S streams (default 2, LSX_LARGE_STREAMS), each fed by its own host
// thread: the base panels, the narrow solves and the in-panel updates are latency-bound launches of one CTA per prime
// (or a few CTAs per prime) that leave most of the GPU idle, and a group's chain of ~3500 launches is strictly
// ordered -- but nothing orders one group against another, so the tensor-core updates of one group fill the SMs that
// the panel phase of the other leaves free.  The streams fork from and join into ctx->stream through events, so the
// call stays asynchronous for the caller.  With event timing of the depth-256 updates switched on
// (lsx_timing_enable) everything runs on ctx->stream alone, because a kernel's duration means nothing while it
// shares the GPU.
int lsx_blocked_det_residues(lsx_ctx* ctx, const int32_t* dA, int n, int prime_begin, int count, uint32_t* d_res) {
    static const size_t budget = []() {
        const char* e = getenv("LSX_LARGE_WS_MB");
        size_t mb = e ? (size_t)strtoull(e, nullptr, 10) : 24576;
        return (mb < 64 ? 64 : mb) << 20;
    }();
    static const int env_streams = []() {
        const char* e = getenv("LSX_LARGE_STREAMS");
        const int s = e ? atoi(e) : 2;
        return s < 1 ? 1 : (s > 8 ? 8 : s);
    }();
    static const int env_prio = []() {
        const char* e = getenv("LSX_LARGE_PRIO");
        return e ? atoi(e) : 0;
    }();
    static const int env_max_tiles = []() {
        const char* e = getenv("LSX_TC_MAX_TILES");
        return e ? std::max(0, atoi(e)) : 0;
    }();
    static const int env_group = []() {
        const char* e = getenv("LSX_LARGE_GROUP");
        return e ? std::max(1, atoi(e)) : 0;
    }();
    if (count <= 0) return LSX_OK;
    const bool use_tc = !getenv("LSX_NO_TC");
    const int S = ctx->timing ? 1 : std::min(env_streams, count);
    const size_t rt = (size_t)(n + lsx_tc::TM - 1) / lsx_tc::TM, ct = (size_t)(n + lsx_tc::TN - 1) / lsx_tc::TN;
    const size_t ap_per = rt * lsx_tc::TM * 4 * NB_OUT, bp_per = ct * lsx_tc::TN * 4 * NB_OUT;   // byte planes per prime
    const size_t per = (size_t)n * n * 4 + (size_t)n * 4 + 64 + (use_tc ? ap_per + bp_per : 0);
    // groups of at most one prime per SM (the base-panel kernel is one CTA per prime), evenly sized, within the
    // workspace budget
    const int share = (count + S - 1) / S;                   // primes per stream
    const int cap = env_group ? env_group : ctx->sm_count;   // measured: 2 x 127 beats 2 x 74 and 1 x 145 (profiles/r02m)
    int G = (int)std::min<size_t>((size_t)share, std::max<size_t>(1, budget / S / per));
    if (G > cap) {
        const int ngroups = (share + cap - 1) / cap;
        G = std::min(G, (share + ngroups - 1) / ngroups);
    }
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = (off + bytes + 255) / 256 * 256;
        return o;
    };
    const size_t o_w = take((size_t)G * n * n * 4), o_piv = take((size_t)G * n * 4), o_det = take((size_t)G * 4),
                 o_flag = take((size_t)G * 4), o_ap = take(use_tc ? G * ap_per : 0), o_bp = take(use_tc ? G * bp_per : 0);
    int rc = lsx_ws_reserve(ctx, off * S);
    if (rc != LSX_OK) return rc;
    const size_t smem_trsm = (size_t)(KI * KI + KI * 128) * 4;
    const size_t smem_gemm = (size_t)(KI * (GM + 4) + KI * GN) * 4;
    LSX_CUDA_TRY(ctx, cudaFuncSetAttribute(k_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_trsm));
    LSX_CUDA_TRY(ctx, cudaFuncSetAttribute(k_gemm_int, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_gemm));
    if (use_tc)
    {
        const int smax = (int)std::max(lsx_tc::smem_bytes(lsx_tc::MAX_K, 0), lsx_tc::smem_bytes(lsx_tc::MAX_K, 1));
        LSX_CUDA_TRY(ctx, cudaFuncSetAttribute(lsx_tc::k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smax));
        LSX_CUDA_TRY(ctx, cudaFuncSetAttribute(lsx_tc::k_gemm_tc_rowmap, cudaFuncAttributeMaxDynamicSharedMemorySize, smax));
    }
    // side streams (created once per ctx) start behind everything already enqueued on ctx->stream
    if (S > 1) {
        while ((int)ctx->side_streams.size() < S) {
            cudaStream_t st = nullptr;
            cudaEvent_t ev = nullptr;
            int lo = 0, hi = 0;                              // numerically lower = higher priority
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            const int idx = (int)ctx->side_streams.size();
            const int prio = env_prio == 1 ? (idx & 1 ? lo : hi) : 0;
            LSX_CUDA_TRY(ctx, cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio));
            ctx->side_streams.push_back(st);
            LSX_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            ctx->side_events.push_back(ev);
        }
        if (!ctx->side_fork) LSX_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->side_fork, cudaEventDisableTiming));
        LSX_CUDA_TRY(ctx, cudaEventRecord(ctx->side_fork, ctx->stream));
        for (int s = 0; s < S; ++s) LSX_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->side_streams[s], ctx->side_fork, 0));
    }
    const int ngroups = (count + G - 1) / G;
    std::atomic<int> next_group{0};
    std::vector<cudaError_t> errs(S, cudaSuccess);
    std::vector<int64_t> launched(S, 0);
    auto worker = [&](int s) {
        if (cudaSetDevice(ctx->device) != cudaSuccess) {
            errs[s] = cudaGetLastError();
            return;
        }
        char* base = (char*)ctx->d_ws + (size_t)s * off;
        cudaStream_t st = S > 1 ? ctx->side_streams[s] : ctx->stream;
        for (;;) {
            const int gi = next_group.fetch_add(1);
            if (gi >= ngroups) break;
            const int g0 = gi * G, Gc = std::min(G, count - g0);
            Driver d{};
            d.ctx = ctx;
            d.stream = st;
            d.timing = S == 1 && ctx->timing;
            d.max_tiles = S > 1 ? env_max_tiles : 0;
            d.a.W = (uint32_t*)(base + o_w);
            d.a.primes = ctx->d_primes + prime_begin + g0;
            d.a.piv_row = (int32_t*)(base + o_piv);
            d.a.detM = (uint32_t*)(base + o_det);
            d.a.flags = (int32_t*)(base + o_flag);
            d.a.n = n;
            d.a.G = Gc;
            d.AP = (uint8_t*)(base + o_ap);
            d.BP = (uint8_t*)(base + o_bp);
            d.use_tc = use_tc;
            d.fuse_pair = !getenv("LSX_LU_PAIR") || atoi(getenv("LSX_LU_PAIR")) != 0;
            d.trsm64 = !getenv("LSX_TRSM64") || atoi(getenv("LSX_TRSM64")) != 0;
            d.n = n;
            k_load<<<dim3(ctx->sm_count * 2, Gc), 256, 0, st>>>(dA, d.a);
            d.run();
            k_finish<<<(Gc + 127) / 128, 128, 0, st>>>(d.a, d_res + g0);
            launched[s] += d.launches + 2;
            const cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) {
                errs[s] = e;
                break;
            }
        }
    };
    if (S == 1) {
        worker(0);
    } else {
        std::vector<std::thread> threads;
        for (int s = 0; s < S; ++s) threads.emplace_back(worker, s);
        for (auto& t : threads) t.join();
        for (int s = 0; s < S; ++s) {                         // join, also after an error: ctx->stream must not run ahead
            cudaEventRecord(ctx->side_events[s], ctx->side_streams[s]);
            cudaStreamWaitEvent(ctx->stream, ctx->side_events[s], 0);
        }
    }
    for (int s = 0; s < S; ++s) ctx->launches += launched[s];
    for (int s = 0; s < S; ++s)
        if (errs[s] != cudaSuccess)
            return lsx_fail(ctx, LSX_ERR_CUDA, "blocked LU launch failed on stream %d: %s", s, cudaGetErrorString(errs[s]));
    return LSX_OK;
}
