// Fused register-resident kernels for batches of small matrices.
//
// k_inv_tpm<N, HEAD>: inverse + determinant of n x n matrices (n <= 8), ONE launch per batch, one
// thread per matrix.  Input is read once (coalesced, staged through shared memory), the adjugate, the
// determinant and the status word are written once -- the algorithmic byte count of SURVEY.md
// section 8d -- and everything in between lives in registers (mirror of the arithmetic:
// tests/device_model.py::inverse_inplace_v2):
//   * in-place division-free Gauss-Jordan modulo ONE 31-bit prime.  The path is taken only when the
//     Hadamard bound of every minor of A is below the prime, so zero tests modulo p are exact (no bad
//     primes; the pivot row choice equals the reference's, linalg.py:548-567) and the adjugate
//     entries are recovered exactly by the symmetric lift;
//   * the first HEAD pivot steps run on plain int32 (two IMADs per entry, no reduction) while the
//     entries provably fit; the remaining steps use Montgomery words (one two-product REDC per entry);
//   * pivot rows are left unscaled; the per-row factor, the sign and the single modular inversion are
//     folded into the multipliers of the LAST pivot step, so there is no separate scaling pass;
//   * the determinant, which may need one more bit than the prime offers, comes from the exact
//     integer identity det = sum_c A[0][c] * adj[c][0];
//   * singular matrices (no pivot in some column) get LSX_ST_SINGULAR and zeros, which is where the
//     reference returns NoSolution() (linalg.py:725-737).
#include "lsx_internal.h"

namespace {

#ifndef LSX_TPM_THREADS
#define LSX_TPM_THREADS 128
#endif
#ifndef LSX_TPM_MINB
#define LSX_TPM_MINB 4
#endif
constexpr int TPM_THREADS = LSX_TPM_THREADS;

template <int N>
struct TpmSmem {
    static constexpr int E = N * N;
    static constexpr int STRIDE = E | 1;     // odd stride: lane t reads word t*STRIDE + e without bank conflicts
};

__device__ __forceinline__ uint32_t mont_sqn(uint32_t x, int k, uint32_t p, uint32_t pinv) {
    for (int i = 0; i < k; ++i) x = mont_mul(x, x, p, pinv);
    return x;
}
// a^(p-2): addition chain (30 squarings + 8 products) for p = 2^31 - 1, square-and-multiply otherwise
__device__ __forceinline__ uint32_t mont_inverse(uint32_t a, const PrimeRec& P) {
    const uint32_t p = P.p, pinv = P.pinv;
    if (p != 0x7fffffffu) return mont_pow(a, p - 2u, P.one, p, pinv);
    const uint32_t x2 = mont_mul(mont_sqn(a, 1, p, pinv), a, p, pinv);
    const uint32_t x4 = mont_mul(mont_sqn(x2, 2, p, pinv), x2, p, pinv);
    const uint32_t x8 = mont_mul(mont_sqn(x4, 4, p, pinv), x4, p, pinv);
    const uint32_t x16 = mont_mul(mont_sqn(x8, 8, p, pinv), x8, p, pinv);
    const uint32_t x24 = mont_mul(mont_sqn(x16, 8, p, pinv), x8, p, pinv);
    const uint32_t x28 = mont_mul(mont_sqn(x24, 4, p, pinv), x4, p, pinv);
    const uint32_t x29 = mont_mul(mont_sqn(x28, 1, p, pinv), a, p, pinv);
    return mont_mul(mont_sqn(x29, 2, p, pinv), a, p, pinv);
}

template <int N, int HEAD, bool I8>
__global__ void __launch_bounds__(TPM_THREADS, LSX_TPM_MINB)
k_inv_tpm(const int32_t* __restrict__ A, int64_t batch, PrimeRec P, int a_abs_max, int32_t* __restrict__ adj,
          int32_t* __restrict__ det, int32_t* __restrict__ status) {
    constexpr int E = TpmSmem<N>::E, ST = TpmSmem<N>::STRIDE;
    extern __shared__ uint32_t sm[];
    const int tid = threadIdx.x;
    const int64_t tile0 = (int64_t)blockIdx.x * TPM_THREADS;           // first matrix of this block
    const int64_t nmat = min((int64_t)TPM_THREADS, batch - tile0);
    const int64_t nwords = nmat * E;
    const uint32_t p = P.p, pinv = P.pinv;

    // ---- coalesced load of the block's matrices into shared memory ----
    if (I8) {
        // int8 entries: 16 per 16-byte load when the block's bytes allow it
        const int8_t* src = reinterpret_cast<const int8_t*>(A) + tile0 * E;
        if ((E % 16) == 0) {
            const int4* src4 = reinterpret_cast<const int4*>(src);
            const int n16 = (int)(nwords >> 4);
            for (int g = tid; g < n16; g += TPM_THREADS) {
                const int4 v = __ldg(src4 + g);
                const int w = g * 16;
                uint32_t* d = sm + (w / E) * ST + (w % E);
                const int q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 16; ++i) d[i] = (uint32_t)(int32_t)(int8_t)(q[i >> 2] >> (8 * (i & 3)));
            }
        } else {
            for (int w = tid; w < (int)nwords; w += TPM_THREADS) sm[(w / E) * ST + (w % E)] = (uint32_t)(int32_t)__ldg(src + w);
        }
    } else {
        const int32_t* src = A + tile0 * E;
        if ((E % 4) == 0) {
            const int4* src4 = reinterpret_cast<const int4*>(src);
            const int n4 = (int)(nwords >> 2);
#pragma unroll 4
            for (int g = tid; g < n4; g += TPM_THREADS) {
                const int4 v = __ldg(src4 + g);
                const int w = g * 4;
                uint32_t* d = sm + (w / E) * ST + (w % E);
                d[0] = (uint32_t)v.x;
                d[1] = (uint32_t)v.y;
                d[2] = (uint32_t)v.z;
                d[3] = (uint32_t)v.w;
            }
        } else {
            for (int w = tid; w < (int)nwords; w += TPM_THREADS) sm[(w / E) * ST + (w % E)] = (uint32_t)__ldg(src + w);
        }
    }
    __syncthreads();

    const bool active = tid < nmat;
    uint32_t W[N][N];          // head: two's-complement integers; tail: Montgomery words
    int32_t a0[N];
    int amax = 0;
    {
        const uint32_t* mine = sm + tid * ST;
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) {
                const int32_t v = active ? (int32_t)mine[r * N + c] : (r == c ? 1 : 0);
                if (r == 0) a0[c] = v;
                amax = max(amax, v < 0 ? -v : v);
                if (v == INT32_MIN) amax = INT32_MAX;
                W[r][c] = (uint32_t)v;
            }
    }
    const bool bound_bad = amax > a_abs_max;
    if (bound_bad) {           // keep the integer head inside its proven range: compute on the identity instead
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) W[r][c] = r == c ? 1u : 0u;
    }

    uint32_t unit = 0;                      // N fields of 4 bits: unit[r]
#pragma unroll
    for (int r = 0; r < N; ++r) unit |= (uint32_t)r << (4 * r);
    uint32_t outcol = 0;                    // N fields of 4 bits: column of the result held in slot j
    bool neg = false, singular = false;
    uint32_t cw[N];                         // per-row multiplier words (what pivot row k still lacks)
    uint32_t sig = 1u;                      // head: product of the pivots so far (exact integer)
    uint32_t S = 1u, Q = P.one;

#pragma unroll
    for (int j = 0; j < N; ++j) {
        constexpr bool kDummy = false;
        (void)kDummy;
        const bool head = j < HEAD;
        const bool last = j == N - 1;
        if (j == HEAD) {
            // ---- switch to Montgomery words: raw load, S = sigma_h raw, Q = word(prod sigma_k) ----
            uint32_t qh = 1u;               // plain product of the head sigmas modulo p
            uint32_t sk = 1u;
#pragma unroll
            for (int k = 0; k < HEAD; ++k) {
                // cw[k] was stored as the plain integer sigma_k (positive or negative, small)
                const uint32_t sw = word_of_int((int32_t)cw[k], p);
                qh = (uint32_t)(((uint64_t)qh * sw) % p);
                cw[k] = mont_mul(sw, P.r2, p, pinv);
                sk = sw;
            }
            (void)sk;
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int c = 0; c < N; ++c) W[r][c] = word_of_int((int32_t)W[r][c], p);
            S = word_of_int((int32_t)sig, p);
            Q = mont_mul(qh, P.r2, p, pinv);
        }
        if (!singular) {
            int src = -1;
#pragma unroll
            for (int r = N - 1; r >= j; --r)
                if (W[r][j] != 0u) src = r;
            if (src < 0) {
                singular = true;
            } else {
                if (src != j) {
#pragma unroll
                    for (int r = j + 1; r < N; ++r) {
                        const bool sw = r == src;          // selects, not a branch: W must stay in registers
#pragma unroll
                        for (int c = 0; c < N; ++c) {
                            const uint32_t a = W[j][c], b = W[r][c];
                            W[j][c] = sw ? b : a;
                            W[r][c] = sw ? a : b;
                        }
                    }
                    const uint32_t uj = (unit >> (4 * j)) & 15u, us = (unit >> (4 * src)) & 15u;
                    unit &= ~((15u << (4 * j)) | (15u << (4 * src)));
                    unit |= (us << (4 * j)) | (uj << (4 * src));
                    neg = !neg;
                }
                outcol |= ((unit >> (4 * j)) & 15u) << (4 * j);
                const uint32_t piv = W[j][j];
                uint32_t prow[N];
#pragma unroll
                for (int c = 0; c < N; ++c) prow[c] = W[j][c];
                if (head) {
                    // plain two's-complement integers: W[r][c] = piv * W[r][c] - f * prow[c]
#pragma unroll
                    for (int r = 0; r < N; ++r) {
                        if (r == j) continue;
                        const uint32_t f = W[r][j];
#pragma unroll
                        for (int c = 0; c < N; ++c)
                            W[r][c] = (c == j) ? (0u - f * sig) : (piv * W[r][c] - f * prow[c]);
                    }
                    W[j][j] = sig;
                    cw[j] = sig;
                    sig *= piv;
                } else {
                    cw[j] = S;
                    Q = mont_mul(Q, S, p, pinv);
                    uint32_t qinv = 0u;
                    if (last) {
                        qinv = mont_inverse(Q, P);
                        if (neg) qinv = p - qinv;           // Q is a unit, so qinv != 0
                    }
#pragma unroll
                    for (int r = 0; r < N; ++r) {
                        if (r == j) {
                            if (last) {
                                const uint32_t g = mont_mul(qinv, cw[r], p, pinv);
#pragma unroll
                                for (int c = 0; c < N; ++c) W[r][c] = mont_mul(g, c == j ? S : prow[c], p, pinv);
                            } else {
                                W[r][j] = S;
                            }
                        } else {
                            const uint32_t f = W[r][j];
                            uint32_t y = f ? p - f : 0u;
                            uint32_t x = piv;
                            if (last) {
                                const uint32_t g = mont_mul(qinv, cw[r], p, pinv);
                                x = mont_mul(g, x, p, pinv);
                                y = mont_mul(g, y, p, pinv);
                            }
#pragma unroll
                            for (int c = 0; c < N; ++c)
                                W[r][c] = (c == j) ? mont_mul(y, S, p, pinv) : mont_fma2(x, W[r][c], y, prow[c], p, pinv);
                        }
                    }
                    S = mont_mul(S, piv, p, pinv);
                }
            }
        }
    }

    // ---- symmetric lift, undo the column permutation, determinant ----
    const uint32_t half = p >> 1;
    const bool zero_out = singular || bound_bad;
    __syncthreads();                       // everybody has read its input tile: reuse it for the output
    {
        uint32_t* mine = sm + tid * ST;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const int cj = zero_out ? j : (int)((outcol >> (4 * j)) & 15u);   // zeros go to every slot
#pragma unroll
            for (int r = 0; r < N; ++r) {
                const uint32_t v = W[r][j];
                const int32_t s = v > half ? (int32_t)(v - p) : (int32_t)v;
                mine[r * N + cj] = zero_out ? 0u : (uint32_t)s;
            }
        }
        // det = row 0 of A times column 0 of adj (exact; both factors are small)
        long long dsum = 0;
#pragma unroll
        for (int c = 0; c < N; ++c) dsum += (long long)a0[c] * (long long)(int32_t)mine[c * N];
        if (active) {
            det[tile0 + tid] = (int32_t)dsum;
            status[tile0 + tid] = (singular && !bound_bad ? LSX_ST_SINGULAR : 0) | (bound_bad ? LSX_ST_BOUND : 0);
        }
    }
    __syncthreads();
    // ---- coalesced store of the adjugates ----
    {
        int32_t* dst = adj + tile0 * E;
        if ((E % 4) == 0) {
            int4* dst4 = reinterpret_cast<int4*>(dst);
            const int n4 = (int)(nwords >> 2);
#pragma unroll 4
            for (int g = tid; g < n4; g += TPM_THREADS) {
                const int w = g * 4;
                const uint32_t* s = sm + (w / E) * ST + (w % E);
                int4 v;
                v.x = (int)s[0];
                v.y = (int)s[1];
                v.z = (int)s[2];
                v.w = (int)s[3];
                dst4[g] = v;
            }
        } else {
            for (int w = tid; w < (int)nwords; w += TPM_THREADS) dst[w] = (int32_t)sm[(w / E) * ST + (w % E)];
        }
    }
}

// Leading pivot steps that provably stay inside int32: entries grow like B -> 2 B^2 per step.
int head_steps_for(int n, int64_t a_abs_max) {
    int h = 0;
    double B = (double)(a_abs_max < 1 ? 1 : a_abs_max);
    const int hmax = n - 1 < 3 ? n - 1 : 3;
    while (h < hmax) {
        B = 2.0 * B * B;
        if (B >= 2147483648.0) break;
        ++h;
    }
    return h;
}

template <int N, int HEAD, bool I8>
int launch_inv_tpm_h(lsx_ctx* ctx, const ElimJob& job) {
    const size_t smem = (size_t)TPM_THREADS * TpmSmem<N>::STRIDE * 4;
    if (smem > 48 * 1024)
        LSX_CUDA_TRY(ctx, cudaFuncSetAttribute(k_inv_tpm<N, HEAD, I8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((job.batch + TPM_THREADS - 1) / TPM_THREADS);
    const PrimeRec P = lsx_make_prime_rec(ctx->primes[0]);
    lsx_timing_begin(ctx);
    k_inv_tpm<N, HEAD, I8><<<grid, TPM_THREADS, smem, ctx->stream>>>(job.A, job.batch, P, (int)job.a_abs_max,
                                                                 (int32_t*)job.num, (int32_t*)job.den, job.status);
    lsx_timing_end(ctx);
    ctx->launches++;
    LSX_CUDA_TRY(ctx, cudaGetLastError());
    return LSX_OK;
}

// Two instantiations per size: the full integer head when the declared magnitudes allow it, none otherwise.
template <int N>
int launch_inv_tpm(lsx_ctx* ctx, const ElimJob& job) {
    constexpr int HMAX = N - 1 < 3 ? N - 1 : 3;
    const char* e = getenv("LSX_TPM_HEAD");
    const int want = e ? atoi(e) : HMAX;
    const bool head = HMAX > 0 && want >= HMAX && head_steps_for(N, job.a_abs_max) >= HMAX;
    if (job.in_i8) return head ? launch_inv_tpm_h<N, HMAX, true>(ctx, job) : launch_inv_tpm_h<N, 0, true>(ctx, job);
    return head ? launch_inv_tpm_h<N, HMAX, false>(ctx, job) : launch_inv_tpm_h<N, 0, false>(ctx, job);
}

}  // namespace

// The single-prime fused inverse applies when every minor of A is below the first table prime
// (exact zero tests), the adjugate entries fit the symmetric range of that prime and the
// determinant fits one limb.
static bool inv_tpm_applies(const lsx_ctx* ctx, const ElimJob& job) {
    if (job.op != LSX_OP_INVERSE || job.m > 8 || job.L != 1 || !job.right_identity) return false;
    if (getenv("LSX_DISABLE_SMALL")) return false;
    const double pbits = std::log2((double)ctx->primes[0]);
    const double det_bits = lsx_log2_minor_bound(job.m, job.m, false, job.a_abs_max, 0, false, job.m);
    const double adj_bits = job.m > 1 ? lsx_log2_minor_bound(job.m - 1, job.m - 1, false, job.a_abs_max, 0, false, job.m - 1) : 0.0;
    return det_bits < pbits - 1e-6 && adj_bits + 1.0 < pbits - 1e-6 && det_bits + 1.0 <= 32.0;
}

int lsx_run_small(lsx_ctx* ctx, const ElimJob& job, int* handled) {
    *handled = 0;
    if (!inv_tpm_applies(ctx, job)) return LSX_OK;
    int rc;
    switch (job.m) {
        case 1: rc = launch_inv_tpm<1>(ctx, job); break;
        case 2: rc = launch_inv_tpm<2>(ctx, job); break;
        case 3: rc = launch_inv_tpm<3>(ctx, job); break;
        case 4: rc = launch_inv_tpm<4>(ctx, job); break;
        case 5: rc = launch_inv_tpm<5>(ctx, job); break;
        case 6: rc = launch_inv_tpm<6>(ctx, job); break;
        case 7: rc = launch_inv_tpm<7>(ctx, job); break;
        default: rc = launch_inv_tpm<8>(ctx, job); break;
    }
    if (rc == LSX_OK) *handled = 1;
    return rc;
}
