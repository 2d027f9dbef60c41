// Integer-pipe micro-benchmark for the roofline of the modular kernels (the
// int32-IMAD peak is not in MEASURED_PEAKS.json).  Measures, per SM and clock, the sustained rate
// of: 32-bit IMAD, IMAD.WIDE (32x32+64), the Montgomery multiply and the two-product Montgomery
// update used by the elimination kernels, and FP64 FMA for comparison.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_int ubench_int.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t redc(uint64_t t, uint32_t p, uint32_t pinv) {
    uint32_t m = (uint32_t)t * pinv;
    uint64_t u = t + (uint64_t)m * p;
    uint32_t r = (uint32_t)(u >> 32);
    uint32_t s = r - p;
    return r < s ? r : s;
}

template <int MODE, int ILP>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, int iters, uint32_t p, uint32_t pinv) {
    uint32_t a[ILP], b[ILP];
    uint64_t w[ILP];
    double d[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        a[i] = seed + threadIdx.x * 7 + i;
        b[i] = seed * 3 + i * 5 + blockIdx.x;
        w[i] = a[i];
        d[i] = a[i] * 1e-3;
    }
    const uint32_t x = seed | 1, y = (seed * 77) | 3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) a[i] = a[i] * x + y;                                  // IMAD
            if (MODE == 1) w[i] = (uint64_t)(uint32_t)w[i] * x + w[i];           // IMAD.WIDE
            if (MODE == 2) a[i] = redc((uint64_t)a[i] * x, p, pinv);             // mont_mul
            if (MODE == 3) a[i] = redc((uint64_t)a[i] * x + (uint64_t)b[i] * y, p, pinv);   // mont_fma2
            if (MODE == 4) d[i] = fma(d[i], 1.0000001, 0.5);                     // DFMA
            if (MODE == 5) a[i] = redc(((uint64_t)a[i] << 32) + (uint64_t)b[i] * y, p, pinv); // mulsub form
            if (MODE == 6) a[i] = __umulhi(a[i], x) + b[i];                      // IMAD.HI (32-bit addend)
            if (MODE == 7) { const uint64_t t = (uint64_t)a[i] * x; a[i] = (uint32_t)t ^ (uint32_t)(t >> 32); }  // IMAD.WIDE (no addend) + LOP3
            if (MODE == 8) a[i] = min(a[i] + x, a[i] ^ y);                       // ALU pipe: LOP3 + VIADDMNMX
            if (MODE == 9) { a[i] = a[i] * x + y; b[i] = min(b[i] + x, b[i]) ^ a[i]; }   // IMAD + 2 ALU ops per iteration (dual pipe)
            if (MODE == 10) { const uint64_t t = (uint64_t)a[i] * x + (uint64_t)b[i] * y; a[i] = (uint32_t)(t & 0x7fffffffu) + (uint32_t)(t >> 31); a[i] = min(a[i], a[i] - p); }  // Mersenne-style fold (not exact: rate probe)
            if (MODE == 11) { a[i] = a[i] > x ? b[i] : a[i]; b[i] = b[i] > y ? a[i] : b[i]; }    // ISETP + SEL pairs
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc += a[i] + (uint32_t)w[i] + (uint32_t)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, uint32_t* out, int sms, double clk_ghz) {
    constexpr int ILP = 8;
    const int iters = 4096, blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE, ILP><<<blocks, threads>>>(out, 12345u, 64, 2147483647u, 2147483649u);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k<MODE, ILP><<<blocks, threads>>>(out, 12345u, iters, 2147483647u, 2147483649u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double ops = (double)blocks * threads * iters * ILP;
    const double rate = ops / (best * 1e-3);
    printf("{\"op\": \"%s\", \"gops\": %.1f, \"per_sm_per_clk_at_%.3fGHz\": %.2f, \"ms\": %.3f}\n", name, rate * 1e-9,
           clk_ghz, rate / sms / (clk_ghz * 1e9), best);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ghz = clk_khz * 1e-6;
    uint32_t* out;
    cudaMalloc(&out, (size_t)sms * 8 * 256 * 4);
    printf("{\"device\": \"%s\", \"sms\": %d, \"max_clock_ghz\": %.3f}\n", prop.name, sms, ghz);
    run<0>("imad32", out, sms, ghz);
    run<1>("imad_wide", out, sms, ghz);
    run<2>("mont_mul", out, sms, ghz);
    run<3>("mont_fma2", out, sms, ghz);
    run<5>("mont_mulsub_shift", out, sms, ghz);
    run<6>("imad_hi_plus_add", out, sms, ghz);
    run<7>("imad_wide_rz_plus_lop", out, sms, ghz);
    run<8>("alu_lop_viaddmnmx_pairs", out, sms, ghz);
    run<9>("imad_plus_2alu", out, sms, ghz);
    run<10>("fma2_mersenne_fold", out, sms, ghz);
    run<11>("setp_sel_pairs", out, sms, ghz);
    run<4>("dfma", out, sms, ghz);
    cudaFree(out);
    return 0;
}
