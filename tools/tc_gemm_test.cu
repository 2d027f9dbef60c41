// Stand-alone check + timing of the tensor-core trailing update (lsx_tc.cuh) against a naive integer kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I linalg_solver_b200/csrc \
//        tools/tc_gemm_test.cu linalg_solver_b200/csrc/lsx_primes.cpp -o tools/tc_gemm_test
//   tools/tc_gemm_test n G k0 K [dbg] [reps] [b_stationary] [cols] [rows]
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "lsx_tc.cuh"

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e = (x);                                                               \
        if (e != cudaSuccess) {                                                            \
            printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            return 2;                                                                      \
        }                                                                                  \
    } while (0)

__global__ void k_ref(uint32_t* W, const PrimeRec* primes, lsx_tc::Region g) {
    const int c = g.c0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int r = g.r0 + blockIdx.y;
    const int prime = blockIdx.z;
    if (c >= g.c1 || r >= g.r1) return;
    const PrimeRec P = primes[prime];
    uint32_t* Wg = W + (int64_t)prime * g.n * g.n;
    uint64_t acc = (uint64_t)Wg[(int64_t)r * g.n + c] << 32;
    for (int k = 0; k < g.K; ++k) {
        acc += (uint64_t)Wg[(int64_t)r * g.n + g.k0 + k] * Wg[(int64_t)(g.k0 + k) * g.n + c];
        uint32_t hi = (uint32_t)(acc >> 32);
        hi = min(hi, hi - P.p);
        acc = ((uint64_t)hi << 32) | (uint32_t)acc;
    }
    Wg[(int64_t)r * g.n + c] = mont_redc(acc, P.p, P.pinv);
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 512;
    const int G = argc > 2 ? atoi(argv[2]) : 2;
    const int k0 = argc > 3 ? atoi(argv[3]) : 0;
    const int K = argc > 4 ? atoi(argv[4]) : 64;
    const int swap = argc > 5 ? atoi(argv[5]) : 0;
    const int reps = argc > 6 ? atoi(argv[6]) : 3;
    const int bst = argc > 7 ? atoi(argv[7]) : 0;       // 1: B-stationary CTAs
    const int cols = argc > 8 ? atoi(argv[8]) : 0;      // restrict the region to this many columns (0: all)
    const int rows = argc > 9 ? atoi(argv[9]) : 0;      // restrict the region to this many rows (0: all)
    using namespace lsx_tc;
    std::vector<uint32_t> tab;
    lsx_fill_prime_table(tab, G);
    std::vector<PrimeRec> recs(G);
    for (int i = 0; i < G; ++i) recs[i] = lsx_make_prime_rec(tab[i]);
    Region g{};
    g.n = n;
    g.r0 = g.c0 = k0 + K;
    g.r1 = g.c1 = n;
    g.k0 = k0;
    g.K = K;
    g.kc = K < KC ? K : KC;
    if (cols > 0 && g.c0 + cols < g.c1) g.c1 = g.c0 + cols;
    if (rows > 0 && g.r0 + rows < g.r1) g.r1 = g.r0 + rows;
    g.set_tiles();
    g.b_stationary = bst;
    g.tiles_per_cta = bst ? g.m_tiles : g.n_tiles;
    if (g.m_tiles <= 0 || g.n_tiles <= 0) {
        printf("{\"error\": \"empty region\"}\n");
        return 2;
    }
    const size_t words = (size_t)G * n * n;
    std::vector<uint32_t> h(words);
    std::mt19937_64 rng(12345);
    for (int pg = 0; pg < G; ++pg)
        for (size_t i = 0; i < (size_t)n * n; ++i) h[(size_t)pg * n * n + i] = (uint32_t)(rng() % recs[pg].p);
    uint32_t *dW, *dRef;
    PrimeRec* dP;
    uint8_t *dA, *dB;
    CK(cudaMalloc(&dW, words * 4));
    CK(cudaMalloc(&dRef, words * 4));
    CK(cudaMalloc(&dP, G * sizeof(PrimeRec)));
    CK(cudaMalloc(&dA, a_plane_bytes(g) * G));
    CK(cudaMalloc(&dB, b_plane_bytes(g) * G));
    CK(cudaMemcpy(dP, recs.data(), G * sizeof(PrimeRec), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dRef, h.data(), words * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float ms_ref = 0, ms_split = 0, ms_tc = 0;
    CK(cudaEventRecord(e0));
    k_ref<<<dim3((g.c1 - g.c0 + 127) / 128, g.r1 - g.r0, G), 128>>>(dRef, dP, g);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms_ref, e0, e1));
    const size_t smem = smem_bytes(K, bst);
    void (*kern)(GemmArgs) = k_gemm_tc_t<0, 1>;  // epilogue with coalesced C accesses
    if (swap == 1) kern = k_gemm_tc_t<0, 0>;     // epilogue in the TMEM row mapping
    if (swap == 2) kern = k_gemm_tc_t<2, 0>;     // no epilogue arithmetic
    if (swap == 6) kern = k_gemm_tc_t<6, 0>;     // no TMEM loads, no arithmetic: the MMA pipeline alone
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GemmArgs a{};
    a.W = dW;
    a.AP = dA;
    a.BP = dB;
    a.primes = dP;
    a.g = g;
    for (int rep = 0; rep < reps; ++rep) {
        CK(cudaMemcpy(dW, h.data(), words * 4, cudaMemcpyHostToDevice));
        CK(cudaEventRecord(e0));
        launch_split(dW, dA, dB, g, G, 0);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms_split, e0, e1));
        CK(cudaEventRecord(e0));
        kern<<<dim3(bst ? g.n_tiles : g.m_tiles, 1, G), THREADS, smem>>>(a);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms_tc, e0, e1));
    }
    std::vector<uint32_t> got(words), ref(words);
    CK(cudaMemcpy(got.data(), dW, words * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ref.data(), dRef, words * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0, first_bad = 0;
    for (size_t i = 0; i < words; ++i)
        if (got[i] != ref[i]) {
            if (!bad) first_bad = i;
            ++bad;
        }
    const double macs = (double)G * (g.r1 - g.r0) * (double)(g.c1 - g.c0) * K;
    printf("{\"bst\": %d, \"rows\": %d, \"cols\": %d, ", bst, g.r1 - g.r0, g.c1 - g.c0);
    printf("\"n\": %d, \"G\": %d, \"k0\": %d, \"K\": %d, \"swap\": %d, \"mismatches\": %zu, \"first_bad\": %zu, "
           "\"first_bad_rc\": [%zu, %zu], \"got\": %u, \"ref\": %u, \"ms_ref\": %.3f, \"ms_split\": %.3f, \"ms_tc\": %.3f, "
           "\"mod_mac_per_s\": %.4g, \"int8_tops\": %.1f}\n",
           n, G, k0, K, swap, bad, first_bad, (first_bad % ((size_t)n * n)) / n, first_bad % n, got[first_bad], ref[first_bad],
           ms_ref, ms_split, ms_tc, macs / (ms_tc * 1e-3), macs * 32 / (ms_tc * 1e-3) / 1e12);
    return bad ? 1 : 0;
}
