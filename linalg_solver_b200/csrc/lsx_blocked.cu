// Blocked right-looking modular LU for ONE large integer matrix and MANY primes at once (config 5 of
// BASELINE.json: a 4096 x 4096 determinant, primes sharded over the GPUs).
//
// det(A) mod p = sign * product of the pivots of the forward sweep of reference linalg.py:547-609
// (any non-zero pivot gives the same determinant; the first non-zero at or below the diagonal is
// used, like the reference).  For a group of G primes the residue matrices W_g = A mod p_g live side
// by side in HBM ([G][n][n] words) and every kernel below works on all of them in one launch:
//   k_load      A -> residues
//   k_panel     one CTA per prime: unblocked elimination of the n x NB panel with row pivoting;
//               multipliers are stored as (p - l) * R mod p ("negated Montgomery form") so that every
//               later update is   redc((w << 32) + lneg * u) = w - l * u   with ONE reduction
//   k_swap_trsm row swaps of the panel applied to the trailing columns + unit-lower triangular solve
//   k_gemm      trailing update A22 -= L21 * U12: register-tiled, 64-bit accumulators with a lazy
//               high-word reduction (one IMAD.WIDE + one VIADDMNMX per multiply-add), one REDC at the end
//   k_finish    sign, zero flag, Montgomery -> plain residue
#include <algorithm>

#include "lsx_internal.h"

namespace {

constexpr int NB = 64;            // panel width
constexpr int PANEL_T = 1024;
constexpr int GM = 128, GN = 64;  // trailing-update tile per CTA (256 threads, 8 x 4 outputs each)

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

__global__ void k_load(const int32_t* __restrict__ A, LargeArgs a) {
    const int64_t nn = (int64_t)a.n * a.n;
    const int g = blockIdx.y;
    const uint32_t p = a.primes[g].p;
    uint32_t* Wg = a.W + (int64_t)g * nn;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += (int64_t)gridDim.x * blockDim.x)
        Wg[i] = word_of_int_any(A[i], p);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.detM[g] = a.primes[g].one;
        a.flags[g] = 0;
    }
}

__global__ void __launch_bounds__(PANEL_T) k_panel(LargeArgs a, int k0, int nb) {
    __shared__ uint32_t prow[NB];
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
        // ---- pivot search: first row >= j with a non-zero entry in column j ----
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
        if (src == INT32_MAX) {                            // column is zero below the diagonal: det = 0
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
            // c = piv^-1 * R^2, so that mont_mul(w, c) = (w / piv) * R
            const uint32_t piv = Wg[(int64_t)j * n + j];
            const uint32_t pivM = mont_mul(piv, P.r2, p, pinv);
            const uint32_t invM = mont_pow(pivM, p - 2u, P.one, p, pinv);
            if (lane == 0) s_c = mont_mul(invM, P.r2, p, pinv);
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

// Trailing columns c >= k0 + nb: apply the panel's row swaps, then solve the unit-lower system.
__global__ void __launch_bounds__(128) k_swap_trsm(LargeArgs a, int k0, int nb) {
    extern __shared__ __align__(16) uint32_t sm[];
    uint32_t* Ln = sm;                    // [nb][NB] negated Montgomery multipliers of L11
    uint32_t* us = Ln + NB * NB;          // [nb][128]
    __shared__ int pr[NB];
    const int n = a.n, g = blockIdx.y, tid = threadIdx.x;
    const PrimeRec P = a.primes[g];
    const uint32_t p = P.p, pinv = P.pinv;
    uint32_t* Wg = a.W + (int64_t)g * n * n;
    for (int e = tid; e < nb * nb; e += 128) {
        const int i = e / nb, t = e % nb;
        Ln[i * NB + t] = Wg[(int64_t)(k0 + i) * n + k0 + t];
    }
    if (tid < nb) pr[tid] = a.piv_row[(int64_t)g * n + k0 + tid];
    __syncthreads();
    const int c = k0 + nb + blockIdx.x * 128 + tid;
    if (c >= n) return;
    for (int jj = 0; jj < nb; ++jj) {
        const int j = k0 + jj, src = pr[jj];
        if (src != j) {
            const uint32_t t0 = Wg[(int64_t)j * n + c];
            Wg[(int64_t)j * n + c] = Wg[(int64_t)src * n + c];
            Wg[(int64_t)src * n + c] = t0;
        }
    }
    for (int i = 0; i < nb; ++i) {
        uint64_t acc = (uint64_t)Wg[(int64_t)(k0 + i) * n + c] << 32;
        for (int t = 0; t < i; ++t) acc = mac_lazy(acc, Ln[i * NB + t], us[t * 128 + tid], p);
        const uint32_t u = mont_redc(acc, p, pinv);
        us[i * 128 + tid] = u;
        Wg[(int64_t)(k0 + i) * n + c] = u;
    }
}

// A22[i][j] = redc((A22[i][j] << 32) + sum_k Lneg[i][k] * U[k][j])   for i, j >= k0 + nb
__global__ void __launch_bounds__(256) k_gemm(LargeArgs a, int k0, int nb) {
    extern __shared__ __align__(16) uint32_t sm[];
    constexpr int LDL = GM + 4;
    uint32_t* Lt = sm;                    // [NB][LDL]  (k-major: Lt[k][i])
    uint32_t* Us = Lt + NB * LDL;         // [NB][GN]
    const int n = a.n, g = blockIdx.z, tid = threadIdx.x;
    const PrimeRec P = a.primes[g];
    const uint32_t p = P.p, pinv = P.pinv;
    uint32_t* Wg = a.W + (int64_t)g * n * n;
    const int r0 = k0 + nb + blockIdx.y * GM, c0 = k0 + nb + blockIdx.x * GN;
    // ---- stage L21 tile (transposed) and U12 tile ----
    for (int e = tid; e < GM * NB; e += 256) {
        const int i = e / NB, k = e % NB;
        const int r = r0 + i;
        Lt[k * LDL + i] = (r < n && k < nb) ? Wg[(int64_t)r * n + k0 + k] : 0u;
    }
    for (int e = tid; e < NB * GN; e += 256) {
        const int k = e / GN, j = e % GN;
        const int c = c0 + j;
        Us[k * GN + j] = (c < n && k < nb) ? Wg[(int64_t)(k0 + k) * n + c] : 0u;
    }
    __syncthreads();
    const int tx = tid & 15, ty = tid >> 4;           // 16 x 16 threads: 4 columns x 8 rows each
    uint64_t acc[8][4];
#pragma unroll
    for (int x = 0; x < 8; ++x) {
        const int r = r0 + ty * 8 + x;
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int c = c0 + tx * 4 + y;
            acc[x][y] = (r < n && c < n) ? (uint64_t)Wg[(int64_t)r * n + c] << 32 : 0ull;
        }
    }
#pragma unroll 4
    for (int k = 0; k < NB; ++k) {
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
        const int r = r0 + ty * 8 + x;
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int c = c0 + tx * 4 + y;
            if (r < n && c < n) Wg[(int64_t)r * n + c] = mont_redc(acc[x][y], p, pinv);
        }
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

}  // namespace

int lsx_blocked_det_residues(lsx_ctx* ctx, const int32_t* dA, int n, int prime_begin, int count, uint32_t* d_res) {
    static const size_t budget = []() {
        const char* e = getenv("LSX_LARGE_WS_MB");
        size_t mb = e ? (size_t)strtoull(e, nullptr, 10) : 24576;
        return (mb < 64 ? 64 : mb) << 20;
    }();
    const size_t per = (size_t)n * n * 4 + (size_t)n * 4 + 64;
    int G = (int)std::min<size_t>((size_t)count, std::max<size_t>(1, budget / per));
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = (off + bytes + 255) / 256 * 256;
        return o;
    };
    const size_t o_w = take((size_t)G * n * n * 4), o_piv = take((size_t)G * n * 4), o_det = take((size_t)G * 4),
                 o_flag = take((size_t)G * 4);
    int rc = lsx_ws_reserve(ctx, off);
    if (rc != LSX_OK) return rc;
    char* base = (char*)ctx->d_ws;
    const size_t smem_trsm = (size_t)(NB * NB + NB * 128) * 4;
    const size_t smem_gemm = (size_t)(NB * (GM + 4) + NB * GN) * 4;
    LSX_CUDA_TRY(ctx, cudaFuncSetAttribute(k_swap_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_trsm));
    LSX_CUDA_TRY(ctx, cudaFuncSetAttribute(k_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_gemm));
    for (int g0 = 0; g0 < count; g0 += G) {
        const int Gc = std::min(G, count - g0);
        LargeArgs a{};
        a.W = (uint32_t*)(base + o_w);
        a.primes = ctx->d_primes + prime_begin + g0;
        a.piv_row = (int32_t*)(base + o_piv);
        a.detM = (uint32_t*)(base + o_det);
        a.flags = (int32_t*)(base + o_flag);
        a.n = n;
        a.G = Gc;
        k_load<<<dim3(ctx->sm_count * 2, Gc), 256, 0, ctx->stream>>>(dA, a);
        ctx->launches++;
        for (int k0 = 0; k0 < n; k0 += NB) {
            const int nb = std::min(NB, n - k0);
            k_panel<<<Gc, PANEL_T, 0, ctx->stream>>>(a, k0, nb);
            ctx->launches++;
            const int rest = n - k0 - nb;
            if (rest > 0) {
                k_swap_trsm<<<dim3((rest + 127) / 128, Gc), 128, smem_trsm, ctx->stream>>>(a, k0, nb);
                lsx_timing_begin(ctx);
                k_gemm<<<dim3((rest + GN - 1) / GN, (rest + GM - 1) / GM, Gc), 256, smem_gemm, ctx->stream>>>(a, k0, nb);
                lsx_timing_end(ctx);
                ctx->launches += 2;
            }
        }
        k_finish<<<(Gc + 127) / 128, 128, 0, ctx->stream>>>(a, d_res + g0);
        ctx->launches++;
        LSX_CUDA_TRY(ctx, cudaGetLastError());
    }
    return LSX_OK;
}
