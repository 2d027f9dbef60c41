// Fused register-resident kernels for batches of small matrices.
//
// k_inv_tpm<N>: inverse + determinant of n x n matrices (n <= 8), ONE launch per batch, one thread
// per matrix.  Input is read once (coalesced, staged through shared memory), the adjugate, the
// determinant and the status word are written once -- the algorithmic byte count of SURVEY.md
// section 8d -- and everything in between lives in registers:
//   * in-place uniform-scale Gauss-Jordan on Montgomery words modulo ONE 31-bit prime (mirror:
//     tests/device_model.py::inverse_inplace_words).  The path is taken only when the Hadamard bound
//     of every minor of A is below the prime, so zero tests modulo p are exact (no bad primes, the
//     pivot row choice equals the reference's, linalg.py:548-567) and the adjugate entries are
//     recovered exactly by the symmetric lift;
//   * the determinant, which may need one more bit than the prime offers, comes from the exact
//     integer identity det = sum_c A[0][c] * adj[c][0];
//   * singular matrices (no pivot in some column) get LSX_ST_SINGULAR and zeros, which is where the
//     reference returns NoSolution() (linalg.py:725-737).
#include "lsx_internal.h"

namespace {

constexpr int TPM_THREADS = 128;

template <int N>
struct TpmSmem {
    static constexpr int E = N * N;
    static constexpr int STRIDE = E | 1;     // odd stride: lane t reads word t*STRIDE + e without bank conflicts
};

template <int N>
__global__ void __launch_bounds__(TPM_THREADS, 3)
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
    {
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
    uint32_t W[N][N];
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
                amax = max(amax, v < 0 ? -v : v);     // INT_MIN stays negative and trips the check below
                if (v == INT32_MIN) amax = INT32_MAX;
                W[r][c] = word_of_int(v, p);
            }
    }
    const bool bound_bad = amax > a_abs_max;

    uint32_t S = P.one, Q = P.one, X = 1u, D = 1u;
    uint32_t unit = 0;                      // N fields of 4 bits: unit[r]
#pragma unroll
    for (int r = 0; r < N; ++r) unit |= (uint32_t)r << (4 * r);
    uint32_t outcol = 0;                    // N fields of 4 bits: column of the result held in slot j
    bool neg = false, singular = false;

#pragma unroll
    for (int j = 0; j < N; ++j) {
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
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    if (r == j) {
#pragma unroll
                        for (int c = 0; c < N; ++c)
                            W[r][c] = mont_mul(S, c == j ? D : prow[c], p, pinv);
                    } else {
                        const uint32_t f = W[r][j];
                        const uint32_t y = f ? p - f : 0u;
#pragma unroll
                        for (int c = 0; c < N; ++c)
                            W[r][c] = (c == j) ? mont_mul(y, D, p, pinv) : mont_fma2(piv, W[r][c], y, prow[c], p, pinv);
                    }
                }
                Q = mont_mul(Q, S, p, pinv);
                S = mont_mul(S, piv, p, pinv);
                D = mont_mul(D, piv, p, pinv);
                X = mont_mul(X, P.r2, p, pinv);
            }
        }
    }

    // ---- one inversion, scale to adj = det * A^-1, symmetric lift, undo the column permutation ----
    uint32_t Gw = mont_mul(mont_pow(Q, p - 2u, P.one, p, pinv), X, p, pinv);
    if (neg && Gw) Gw = p - Gw;
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
                const uint32_t v = mont_mul(Gw, W[r][j], p, pinv);
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

template <int N>
int launch_inv_tpm(lsx_ctx* ctx, const ElimJob& job) {
    const size_t smem = (size_t)TPM_THREADS * TpmSmem<N>::STRIDE * 4;
    if (smem > 48 * 1024)
        LSX_CUDA_TRY(ctx, cudaFuncSetAttribute(k_inv_tpm<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((job.batch + TPM_THREADS - 1) / TPM_THREADS);
    const PrimeRec P = lsx_make_prime_rec(ctx->primes[0]);
    lsx_timing_begin(ctx);
    k_inv_tpm<N><<<grid, TPM_THREADS, smem, ctx->stream>>>(job.A, job.batch, P, (int)job.a_abs_max, (int32_t*)job.num,
                                                           (int32_t*)job.den, job.status);
    lsx_timing_end(ctx);
    ctx->launches++;
    LSX_CUDA_TRY(ctx, cudaGetLastError());
    return LSX_OK;
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
