// Lab harness for the fused small-matrix inverse kernel (linalg_solver_b200/csrc/lsx_inv_small.cuh):
// runs the kernel on 2^20 random 8x8 matrices (plus planted singular / zero-heavy cases), compares every output
// word with the round-1 library (tools/lab/liblsx_r1.so through the C-ABI) and prints one JSON line with the
// CUDA-event time of both.  Variants are compile-time (-DLSX_TPM_THREADS=, -DLSX_TPM_MINB=, -DLSX_TPM_WINDOW=).
// Build: see tools/lab/run_inv8_lab.sh
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <random>
#include <vector>

#include "../../linalg_solver_b200/csrc/lsx_inv_small.cuh"

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e__ = (x);                                                                 \
        if (e__ != cudaSuccess) {                                                              \
            fprintf(stderr, "%s failed: %s (line %d)\n", #x, cudaGetErrorString(e__), __LINE__); \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

static PrimeRec make_rec(uint32_t p) {
    PrimeRec r;
    r.p = p;
    uint32_t inv = 1;
    for (int i = 0; i < 6; ++i) inv *= 2u - p * inv;   // p^{-1} mod 2^32
    r.pinv = 0u - inv;
    r.one = (uint32_t)((1ull << 32) % p);
    r.r2 = (uint32_t)(((unsigned __int128)1 << 64) % p);
    return r;
}

typedef int (*create_t)(int, void**);
typedef int (*plan_inv_t)(int, int64_t, lsx_plan*);
typedef int (*inv_t)(void*, const lsx_plan*, const int32_t*, int64_t, int, uint32_t*, uint32_t*, int32_t*);
typedef int (*inv8_t)(void*, const lsx_plan*, const int8_t*, int64_t, int, uint32_t*, uint32_t*, int32_t*);
typedef int (*setstream_t)(void*, void*);

template <int N, int HEAD, bool I8, int XS = 0>
float time_new(const void* dA, int64_t batch, PrimeRec P, int amax, int32_t* adj, int32_t* det, int32_t* st, int reps,
               cudaStream_t s) {
    using namespace lsx_inv_small;
    const size_t smem = TpmTile<N>::BYTES;
    CK(cudaFuncSetAttribute(k_inv_tpm<N, HEAD, I8, XS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((batch + TPM_THREADS - 1) / TPM_THREADS);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) k_inv_tpm<N, HEAD, I8, XS><<<grid, TPM_THREADS, smem, s>>>(dA, batch, P, amax, 1, adj, det, st);
    CK(cudaStreamSynchronize(s));
    CK(cudaEventRecord(e0, s));
    for (int i = 0; i < reps; ++i) k_inv_tpm<N, HEAD, I8, XS><<<grid, TPM_THREADS, smem, s>>>(dA, batch, P, amax, 1, adj, det, st);
    CK(cudaEventRecord(e1, s));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

template <int N, int HEAD, bool I8, int XS = 0>
float time_stream(const void* dA, int64_t batch, PrimeRec P, int amax, int32_t* adj, int32_t* det, int32_t* st, int reps,
                  cudaStream_t s, int ctas_per_sm) {
    using namespace lsx_inv_small;
    const size_t smem = I8 ? (size_t)TPM_THREADS * TpmTile<N>::STB : TpmTile<N>::BYTES;
    CK(cudaFuncSetAttribute(k_inv_tpm_stream<N, HEAD, I8, XS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (batch + TPM_THREADS - 1) / TPM_THREADS;
    const unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)148 * ctas_per_sm);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) k_inv_tpm_stream<N, HEAD, I8, XS><<<grid, TPM_THREADS, smem, s>>>(dA, batch, P, amax, adj, det, st);
    CK(cudaStreamSynchronize(s));
    CK(cudaEventRecord(e0, s));
    for (int i = 0; i < reps; ++i) k_inv_tpm_stream<N, HEAD, I8, XS><<<grid, TPM_THREADS, smem, s>>>(dA, batch, P, amax, adj, det, st);
    CK(cudaEventRecord(e1, s));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main(int argc, char** argv) {
    const int reps = argc > 1 ? atoi(argv[1]) : 20;
    const char* tag = argc > 2 ? argv[2] : "default";
    const int N = 8, E = 64;
    const int64_t batch = 1 << 20;
    std::vector<int32_t> A((size_t)batch * E);
    std::mt19937_64 rng(12345);
    for (auto& v : A) v = (int)(rng() % 11) - 5;
    // planted cases
    for (int c = 0; c < 8; ++c) A[(size_t)3 * E + 5 * 8 + c] = A[(size_t)3 * E + 2 * 8 + c];   // singular: equal rows
    memset(&A[(size_t)7 * E], 0, E * 4);                                                        // zero matrix
    for (int i = 0; i < 4096; ++i) {                                                            // zero-heavy: far pivots
        int32_t* M = &A[(size_t)(1000 + 37 * i) * E];
        for (int r = 0; r < 8; ++r)
            for (int c = 0; c < 8; ++c)
                if ((rng() & 3) != 0) M[r * 8 + c] = 0;
    }
    for (int i = 0; i < 512; ++i) {                                                             // permutation-like
        int32_t* M = &A[(size_t)(500000 + 11 * i) * E];
        int perm[8] = {0, 1, 2, 3, 4, 5, 6, 7};
        for (int k = 7; k > 0; --k) std::swap(perm[k], perm[rng() % (k + 1)]);
        memset(M, 0, E * 4);
        for (int r = 0; r < 8; ++r) M[r * 8 + perm[r]] = (int)(rng() % 5) + 1;
    }
    A[(size_t)9 * E + 17] = 77;                                                                 // out of the declared bound
    std::vector<int8_t> A8(A.size());
    for (size_t i = 0; i < A.size(); ++i) A8[i] = (int8_t)A[i];

    int32_t *dA, *adj0, *det0, *st0, *adj1, *det1, *st1;
    int8_t* dA8;
    CK(cudaMalloc(&dA, A.size() * 4));
    CK(cudaMalloc(&dA8, A.size()));
    CK(cudaMalloc(&adj0, A.size() * 4));
    CK(cudaMalloc(&adj1, A.size() * 4));
    CK(cudaMalloc(&det0, batch * 4));
    CK(cudaMalloc(&det1, batch * 4));
    CK(cudaMalloc(&st0, batch * 4));
    CK(cudaMalloc(&st1, batch * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dA8, A8.data(), A.size(), cudaMemcpyHostToDevice));
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));

    // ---- reference: the round-1 library through the C-ABI ----
    float ref_ms = -1.f;
    void* h = dlopen("tools/lab/liblsx_r1.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        fprintf(stderr, "dlopen: %s\n", dlerror());
        return 1;
    }
    void* ctx = nullptr;
    lsx_plan plan;
    if (((create_t)dlsym(h, "lsx_create"))(0, &ctx) != 0) return 2;
    if (((plan_inv_t)dlsym(h, "lsx_plan_inverse"))(N, 5, &plan) != 0) return 3;
    ((setstream_t)dlsym(h, "lsx_set_stream"))(ctx, (void*)s);
    inv_t inv = (inv_t)dlsym(h, "lsx_inverse_batch");
    {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        for (int i = 0; i < 3; ++i)
            if (inv(ctx, &plan, dA, batch, 1, (uint32_t*)adj0, (uint32_t*)det0, st0) != 0) return 4;
        CK(cudaStreamSynchronize(s));
        CK(cudaEventRecord(e0, s));
        for (int i = 0; i < reps; ++i) inv(ctx, &plan, dA, batch, 1, (uint32_t*)adj0, (uint32_t*)det0, st0);
        CK(cudaEventRecord(e1, s));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ref_ms, e0, e1));
        ref_ms /= reps;
    }
    std::vector<int32_t> h_adj0(A.size()), h_det0(batch), h_st0(batch), h_adj1(A.size()), h_det1(batch), h_st1(batch);
    CK(cudaMemcpy(h_adj0.data(), adj0, A.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_det0.data(), det0, batch * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_st0.data(), st0, batch * 4, cudaMemcpyDeviceToHost));

    const PrimeRec P = make_rec(0x7fffffffu);
    auto compare = [&](const char* what) {
        CK(cudaMemcpy(h_adj1.data(), adj1, A.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h_det1.data(), det1, batch * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h_st1.data(), st1, batch * 4, cudaMemcpyDeviceToHost));
        long bad = 0, first = -1;
        for (int64_t i = 0; i < batch; ++i) {
            bool ok = h_det0[i] == h_det1[i] && h_st0[i] == h_st1[i] &&
                      memcmp(&h_adj0[(size_t)i * E], &h_adj1[(size_t)i * E], E * 4) == 0;
            if (!ok) {
                if (first < 0) first = i;
                ++bad;
            }
        }
        if (bad)
            fprintf(stderr, "%s: %ld matrices differ, first %ld (det %d vs %d, status %d vs %d)\n", what, bad, first,
                    h_det0[first], h_det1[first], h_st0[first], h_st1[first]);
        return bad;
    };
    long bad = 0;
    CK(cudaMemset(adj1, 0xff, A.size() * 4));
    const float ms32 = time_new<8, 3, false>(dA, batch, P, 5, adj1, det1, st1, reps, s);
    bad += compare("int32 head3");
    CK(cudaMemset(adj1, 0xff, A.size() * 4));
    const float ms8 = time_new<8, 3, true>(dA8, batch, P, 5, adj1, det1, st1, reps, s);
    bad += compare("int8 head3");
    CK(cudaMemset(adj1, 0xff, A.size() * 4));
    const float ms32h0 = time_new<8, 0, false>(dA, batch, P, 5, adj1, det1, st1, reps, s);
    bad += compare("int32 head0");
    CK(cudaMemset(adj1, 0xff, A.size() * 4));
    const float ms32f = time_new<8, 3, false, 1>(dA, batch, P, 5, adj1, det1, st1, reps, s);
    bad += compare("int32 head3 + fp64 step");
    CK(cudaMemset(adj1, 0xff, A.size() * 4));
    const float ms8f = time_new<8, 3, true, 1>(dA8, batch, P, 5, adj1, det1, st1, reps, s);
    bad += compare("int8 head3 + fp64 step");
    CK(cudaMemset(adj1, 0xff, A.size() * 4));
    const float ms32m = time_new<8, 3, false, 2>(dA, batch, P, 5, adj1, det1, st1, reps, s);
    bad += compare("int32 head3 + mersenne step");
    CK(cudaMemset(adj1, 0xff, A.size() * 4));
    const float ms8m = time_new<8, 3, true, 2>(dA8, batch, P, 5, adj1, det1, st1, reps, s);
    bad += compare("int8 head3 + mersenne step");
    CK(cudaMemset(adj1, 0xff, A.size() * 4));
    const float ms32b = time_new<8, 4, false, 3>(dA, batch, P, 5, adj1, det1, st1, reps, s);
    bad += compare("int32 bareiss h32=4");
    CK(cudaMemset(adj1, 0xff, A.size() * 4));
    const float ms8b = time_new<8, 4, true, 3>(dA8, batch, P, 5, adj1, det1, st1, reps, s);
    bad += compare("int8 bareiss h32=4");
    CK(cudaMemset(adj1, 0xff, A.size() * 4));
    const float ms32b0 = time_new<8, 0, false, 3>(dA, batch, P, 5, adj1, det1, st1, reps, s);
    bad += compare("int32 bareiss h32=0");
    CK(cudaMemset(adj1, 0xff, A.size() * 4));
    const float ms32bf = time_new<8, 4, false, 4>(dA, batch, P, 5, adj1, det1, st1, reps, s);
    bad += compare("int32 bareiss h32=4 + 2 fp64 steps");
    CK(cudaMemset(adj1, 0xff, A.size() * 4));
    const float ms8bf = time_new<8, 4, true, 4>(dA8, batch, P, 5, adj1, det1, st1, reps, s);
    bad += compare("int8 bareiss h32=4 + 2 fp64 steps");
    printf("{\"tag\": \"%s\", \"bareiss_f64_i32_ms\": %.4f, \"bareiss_f64_i8_ms\": %.4f}\n", tag, ms32bf, ms8bf);
    printf("{\"tag\": \"%s\", \"bareiss_i32_ms\": %.4f, \"bareiss_i8_ms\": %.4f, \"bareiss_i32_h0_ms\": %.4f}\n", tag, ms32b, ms8b, ms32b0);
    CK(cudaMemset(adj1, 0xff, A.size() * 4));
    const float ms32s = time_stream<8, 3, false>(dA, batch, P, 5, adj1, det1, st1, reps, s, LSX_TPM_MINB);
    bad += compare("int32 stream");
    CK(cudaMemset(adj1, 0xff, A.size() * 4));
    const float ms8s = time_stream<8, 3, true>(dA8, batch, P, 5, adj1, det1, st1, reps, s, LSX_TPM_MINB);
    bad += compare("int8 stream");
    CK(cudaMemset(adj1, 0xff, A.size() * 4));
    time_stream<8, 3, false>(dA, 1000, P, 5, adj1, det1, st1, 1, s, LSX_TPM_MINB);
    {
        CK(cudaMemcpy(h_adj1.data(), adj1, (size_t)1001 * E * 4, cudaMemcpyDeviceToHost));
        if (memcmp(h_adj0.data(), h_adj1.data(), (size_t)1000 * E * 4) != 0 || h_adj1[(size_t)1000 * E] != -1) {
            fprintf(stderr, "stream: short batch differs or wrote past the end\n");
            ++bad;
        }
    }
    printf("{\"tag\": \"%s\", \"stream_i32_ms\": %.4f, \"stream_i8_ms\": %.4f}\n", tag, ms32s, ms8s);
    // short batch (last block partly filled)
    CK(cudaMemset(adj1, 0xff, A.size() * 4));
    time_new<8, 3, false>(dA, 1000, P, 5, adj1, det1, st1, 1, s);
    {
        CK(cudaMemcpy(h_adj1.data(), adj1, (size_t)1001 * E * 4, cudaMemcpyDeviceToHost));
        if (memcmp(h_adj0.data(), h_adj1.data(), (size_t)1000 * E * 4) != 0 || h_adj1[(size_t)1000 * E] != -1) {
            fprintf(stderr, "short batch differs or wrote past the end\n");
            ++bad;
        }
    }
    printf("{\"tag\": \"%s\", \"threads\": %d, \"minb\": %d, \"window\": %d, \"ref_r1_ms\": %.4f, \"new_i32_ms\": %.4f, "
           "\"new_i8_ms\": %.4f, \"new_i32_head0_ms\": %.4f, \"new_i32_f64_ms\": %.4f, \"new_i8_f64_ms\": %.4f, \"new_i32_mers_ms\": %.4f, \"new_i8_mers_ms\": %.4f, \"mismatches\": %ld}\n",
           tag, LSX_TPM_THREADS, LSX_TPM_MINB, LSX_TPM_WINDOW, ref_ms, ms32, ms8, ms32h0, ms32f, ms8f, ms32m, ms8m, bad);
    return bad ? 5 : 0;
}
