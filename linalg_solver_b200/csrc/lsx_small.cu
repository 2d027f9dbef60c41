// Fused register-resident kernels for batches of small matrices: the launcher side of
// lsx_inv_small.cuh (k_inv_tpm<N, HEAD, I8>: inverse + determinant of n x n matrices, n <= 8, one launch per
// batch, one thread per matrix; see the header for the algorithm).  Singular matrices get LSX_ST_SINGULAR and
// zeros, which is where the reference returns NoSolution() (linalg.py:725-737).
#include <string.h>

#include "lsx_inv_small.cuh"

namespace {

using namespace lsx_inv_small;

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

template <int N, int HEAD, bool I8, int XS = 0>
int launch_inv_tpm_h(lsx_ctx* ctx, const ElimJob& job) {
    const size_t smem = TpmTile<N>::BYTES;
    if (smem > 48 * 1024)
        LSX_CUDA_TRY(ctx, cudaFuncSetAttribute(k_inv_tpm<N, HEAD, I8, XS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((job.batch + TPM_THREADS - 1) / TPM_THREADS);
    const PrimeRec P = lsx_make_prime_rec(ctx->primes[0]);
    // 16-byte accesses need 16-byte aligned caller pointers (a device view with an odd element offset is legal)
    const int vec_ok = (((uintptr_t)job.A | (uintptr_t)job.num) & 15u) == 0;
    lsx_timing_begin(ctx);
    k_inv_tpm<N, HEAD, I8, XS><<<grid, TPM_THREADS, smem, ctx->stream>>>(job.A, job.batch, P, (int)job.a_abs_max, vec_ok,
                                                                 (int32_t*)job.num, (int32_t*)job.den, job.status);
    lsx_timing_end(ctx);
    ctx->launches++;
    LSX_CUDA_TRY(ctx, cudaGetLastError());
    return LSX_OK;
}

// Leading pivot steps of the integer (Bareiss) kernel whose t = piv * w - f * prow fits 32 bits: step j combines minors
// of order j + 1, so it needs 2 M^2 < 2^31 for their Hadamard bound M.
int h32_steps_for(int n, int64_t a_abs_max) {
    int h = 0;
    for (int j = 0; j < n; ++j) {
        const double bits = lsx_log2_minor_bound(j + 1, j + 1, false, a_abs_max, 0, false, j + 1);
        if (1.0 + 2.0 * bits >= 31.0 - 1e-6) break;
        h = j + 1;
    }
    return h;
}

// Default: the elimination over the integers (tpm_eliminate_bareiss; 126 us per 2^20 8 x 8 matrices on B200 against
// 198 us for the residue kernel, profiles/r02n_inv8_lab_bareiss.jsonl), with 32-bit arithmetic in the first min(N, 4)
// pivot steps when the declared magnitudes allow it and 64-bit products from the start otherwise.
// LSX_TPM_ALGO=mont selects the single-prime residue kernel (round-2 v3; kept for A/B measurements and as a second
// implementation the tests compare word for word): the full integer head when the magnitudes allow it, none otherwise.
template <int N>
int launch_inv_tpm(lsx_ctx* ctx, const ElimJob& job) {
    const char* algo = getenv("LSX_TPM_ALGO");
    const bool mont = algo && strcmp(algo, "mont") == 0;
    if (!mont) {
        constexpr int HB = N < 4 ? N : 4;
        const bool h32 = h32_steps_for(N, job.a_abs_max) >= HB;
        if (job.in_i8) return h32 ? launch_inv_tpm_h<N, HB, true, 3>(ctx, job) : launch_inv_tpm_h<N, 0, true, 3>(ctx, job);
        return h32 ? launch_inv_tpm_h<N, HB, false, 3>(ctx, job) : launch_inv_tpm_h<N, 0, false, 3>(ctx, job);
    }
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
