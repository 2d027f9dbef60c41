// Fused register-resident kernels for batches of small matrices: the launcher side of
// lsx_inv_small.cuh (k_inv_tpm<N, HEAD, I8>: inverse + determinant of n x n matrices, n <= 8, one launch per
// batch, one thread per matrix; see the header for the algorithm).  Singular matrices get LSX_ST_SINGULAR and
// zeros, which is where the reference returns NoSolution() (linalg.py:725-737).
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

// The FP64 step (lsx_inv_small.cuh) computes piv * w - f * prow exactly in a double: the entries after h integer
// steps are bounded by B_h (B -> 2 B^2 per step), so it needs 2 B_h^2 < 2^53.
bool f64_step_ok(int64_t a_abs_max, int h) {
    double B = (double)(a_abs_max < 1 ? 1 : a_abs_max);
    for (int i = 0; i < h; ++i) B = 2.0 * B * B;
    return 2.0 * B * B < 9007199254740992.0;
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

// Two instantiations per size: the full integer head when the declared magnitudes allow it, none otherwise.
template <int N>
int launch_inv_tpm(lsx_ctx* ctx, const ElimJob& job) {
    constexpr int HMAX = N - 1 < 3 ? N - 1 : 3;
    const char* e = getenv("LSX_TPM_HEAD");
    const int want = e ? atoi(e) : HMAX;
    const bool head = HMAX > 0 && want >= HMAX && head_steps_for(N, job.a_abs_max) >= HMAX;
    if constexpr (N == 8) {
        // the benchmark shape: pivot step 3 is still exact integer arithmetic when 2 B^2 < 2^53 for the bound B after
        // the integer head, and can be done without a Montgomery reduction.  LSX_TPM_XS = 2: 64-bit integers folded
        // modulo the Mersenne prime 2^31 - 1 (56 IMAD + 56 IMAD.HI less on the fmaheavy pipe); 1: the same step on the
        // FP64 pipe.  Both are bit-exact (tests) and both measured NEUTRAL on B200 (198.4 / 199.4 vs 198.1 us per 2^20
        // matrices, profiles/r02k_inv8_lab_mersenne_step.jsonl): with 4 warps per scheduler the kernel is bound by
        // its dependency chains and the exposed tile load (ncu: 19 % of the stall samples sit on the first STS.128
        // after the global loads), not by the count of fmaheavy instructions.  Default 0: the plain kernel.
        const char* xs_env = getenv("LSX_TPM_XS");
        const int xs = xs_env ? atoi(xs_env) : 0;
        if (head && xs != 0 && f64_step_ok(job.a_abs_max, HMAX)) {
            if (xs == 2 && ctx->primes[0] == 0x7fffffffu)
                return job.in_i8 ? launch_inv_tpm_h<N, HMAX, true, 2>(ctx, job) : launch_inv_tpm_h<N, HMAX, false, 2>(ctx, job);
            if (xs == 1)
                return job.in_i8 ? launch_inv_tpm_h<N, HMAX, true, 1>(ctx, job) : launch_inv_tpm_h<N, HMAX, false, 1>(ctx, job);
        }
    }
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
