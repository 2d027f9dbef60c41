// Internal declarations shared by the liblsx translation units (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <cmath>
#include <string>
#include <vector>

#include "../../include/lsx.h"

// One table prime with its Montgomery constants (R = 2^32).
struct PrimeRec {
    uint32_t p;     // odd prime < 2^31
    uint32_t pinv;  // -p^{-1} mod 2^32
    uint32_t one;   // R mod p      (word whose value is 1)
    uint32_t r2;    // R^2 mod p    (word whose value is R)
};

constexpr int LSX_TABLE_PRIMES = 2048;   // primes generated per ctx (descending from 2^31-1)
constexpr int LSX_GARNER_DIM = 40;       // Garner inverse table covers the first 40 primes
constexpr int LSX_MAX_BATCH_PRIMES = 32; // K limit of the batched operations
constexpr int LSX_RETRY_EXTRA = 4;       // replacement primes tried for a flagged matrix
constexpr int LSX_RETRY_CAP = 4096;      // flagged matrices handled per chunk
constexpr int LSX_PROF_SKIP = 255;       // profile byte: column has no pivot
constexpr int LSX_ST_INTERNAL_RETRY = 1 << 30;  // internal status bit, cleared before return

struct lsx_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;       // stream work is enqueued on
    cudaStream_t own_stream = nullptr;   // created by lsx_create
    cudaStream_t copy_streams[2] = {nullptr, nullptr};
    cudaEvent_t events[8] = {};
    std::string err;
    int64_t launches = 0;
    std::vector<uint32_t> primes;        // host copy of the table
    PrimeRec* d_primes = nullptr;        // [LSX_TABLE_PRIMES]
    uint32_t* d_garner = nullptr;        // [GARNER_DIM][GARNER_DIM]: (p_i^{-1} * R) mod p_j
    void* d_ws = nullptr;                // grow-only device workspace
    size_t ws_bytes = 0;
    void* d_io = nullptr;                // grow-only staging for LSX_MEM_HOST calls
    size_t io_bytes = 0;
    // optional device timing of the dominant kernel of each call (lsx_timing_enable)
    bool timing = false;
    int tev_used = 0;
    std::vector<cudaEvent_t> tev;        // pairs: [2*i] before, [2*i+1] after
    std::vector<cudaEvent_t> pev;        // events of the host-call pipeline
    // side streams of the blocked LU (independent prime groups run concurrently, lsx_blocked.cu) with one join
    // event each; side_fork orders them behind the work already on `stream`
    std::vector<cudaStream_t> side_streams;
    std::vector<cudaEvent_t> side_events;
    cudaEvent_t side_fork = nullptr;
    // lsx_create_multi: the other GPUs of this context (each a full single-device ctx, driven by its own host
    // thread inside a call); empty for a single-device ctx.  nccl: communicators of all devices (lsx_multi.cpp)
    std::vector<lsx_ctx*> peers;
    void* nccl = nullptr;
    // device word pair of the last generic pass: [1] = primes its data needed (lsx_last_prime_count), or NULL
    const int32_t* last_kword = nullptr;
};
// Record CUDA events around the dominant kernel of an operation when timing is enabled.
void lsx_timing_begin(lsx_ctx* ctx);
void lsx_timing_end(lsx_ctx* ctx);

// ---- Montgomery arithmetic on 31-bit primes ------------------------------------------------
#ifdef __CUDACC__
#define LSX_HD __host__ __device__ __forceinline__
#else
#define LSX_HD inline
#endif

// 32 x 32 -> 64 products and the high word of t + m * p, written as PTX on the device.  From plain C the front end
// sometimes hoists the zero extension of a loop-invariant factor (the prime) and emits a 64 x 64 multiply, which
// ptxas lowers to IMAD.WIDE plus an add of a zero high word per reduction (`IADD3 R, R, UR(=0)` after every
// reduction of the unrolled kernels).  The sequences below are the ones ptxas folds into IMAD.WIDE / IMAD.HI.
LSX_HD uint64_t mul_wide(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint64_t r;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
#else
    return (uint64_t)a * b;
#endif
}
// high 32 bits of t + m * p
LSX_HD uint32_t mad_wide_hi(uint32_t m, uint32_t p, uint64_t t) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm("{\n\t.reg .b64 u;\n\t.reg .b32 lo;\n\tmul.wide.u32 u, %1, %2;\n\tadd.s64 u, u, %3;\n\tmov.b64 {lo, %0}, u;\n\t}"
        : "=r"(r)
        : "r"(m), "r"(p), "l"(t));
    return r;
#else
    return (uint32_t)(((uint64_t)m * p + t) >> 32);
#endif
}

// REDC of t < 2 p^2 (fits 63 bits): returns t / R mod p in [0, p).
LSX_HD uint32_t mont_redc(uint64_t t, uint32_t p, uint32_t pinv) {
    uint32_t m = (uint32_t)t * pinv;
    uint32_t r = mad_wide_hi(m, p, t);
    uint32_t s = r - p;
    return r < s ? r : s;               // r in [0, 2p): subtract p once if needed
}
LSX_HD uint32_t mont_mul(uint32_t a, uint32_t b, uint32_t p, uint32_t pinv) {
    return mont_redc(mul_wide(a, b), p, pinv);
}
// x*a + y*b (one reduction for two products)
LSX_HD uint32_t mont_fma2(uint32_t x, uint32_t a, uint32_t y, uint32_t b, uint32_t p, uint32_t pinv) {
    return mont_redc(mul_wide(x, a) + mul_wide(y, b), p, pinv);
}
LSX_HD uint32_t mont_pow(uint32_t a, uint32_t e, uint32_t one, uint32_t p, uint32_t pinv) {
    uint32_t acc = one;
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (int bit = 31; bit >= 0; --bit) {
        acc = mont_mul(acc, acc, p, pinv);
        if ((e >> bit) & 1u) acc = mont_mul(acc, a, p, pinv);
    }
    return acc;
}
// raw word of a signed input with |a| < p
LSX_HD uint32_t word_of_int(int32_t a, uint32_t p) { return a < 0 ? (uint32_t)a + p : (uint32_t)a; }
// raw word of any int32 (entries at or above p, tiny test primes)
LSX_HD uint32_t word_of_int_any(int32_t a, uint32_t p) {
    int64_t r = (int64_t)a % (int64_t)p;
    return (uint32_t)(r < 0 ? r + (int64_t)p : r);
}

// ---- host helpers (lsx_primes.cpp) -----------------------------------------------------------
bool lsx_is_prime_u32(uint32_t n);
void lsx_fill_prime_table(std::vector<uint32_t>& out, int count);
PrimeRec lsx_make_prime_rec(uint32_t p);
uint32_t lsx_inv_mod(uint32_t a, uint32_t p);   // a^{-1} mod p (plain)
// log2 of the Hadamard bound for the outputs of an elimination
double lsx_log2_minor_bound(int m, int bar, bool has_right, int64_t a_abs, int64_t b_abs,
                            bool right_identity, int max_rank);
void lsx_bits_to_plan(double log2_bound, int* n_primes, int* limbs);

// ---- kernels' host launchers -----------------------------------------------------------------
struct ElimJob {
    // input
    const int32_t* A = nullptr;     // [batch][m][n_in]
    const int32_t* bvec = nullptr;  // [batch][m] appended as last column (solve) or NULL
    int64_t batch = 0;
    int m = 0, n_in = 0, n = 0, bar = 0;
    int right_identity = 0;         // columns n_in..n-1 are the identity
    int in_i8 = 0;                  // A holds int8 entries (fused small inverse only)
    int64_t a_abs_max = 0, b_abs_max = 0;
    int max_rank = 0;
    int op = 0;
    int K = 0, L = 0, gen_cap = 0;
    // outputs (device pointers; which ones are used depends on op)
    uint32_t* num = nullptr;        // rref: [batch][m*n][L]; inverse: [batch][m*m][L]
    uint32_t* den = nullptr;        // [batch][L]
    uint32_t* particular = nullptr; // solve
    uint32_t* generators = nullptr; // solve
    int32_t* pivot_col = nullptr;   // [batch][pivot_slots] or NULL
    int32_t* rank = nullptr;        // [batch] or NULL
    int32_t* status = nullptr;      // [batch]
};

// Generic path: one CTA per (matrix, prime), residues to scratch, then verify + CRT/assemble.
// list == NULL: all matrices of the job.  Otherwise only the matrices in list[0..*list_count).
int lsx_run_generic(lsx_ctx* ctx, const ElimJob& job, const int32_t* list, const int32_t* list_count,
                    int list_cap, size_t ws_offset);
size_t lsx_generic_ws_bytes(const ElimJob& job, int64_t batch);
// kword[0..1] (device) <- primes per matrix the row norms of the job's own matrices need, at most job.K (lsx_tile.cu)
int lsx_row_bound_primes(lsx_ctx* ctx, const ElimJob& job, int32_t* kword);
// Fused register-resident kernels for small shapes; *handled = 1 if the job was covered.
int lsx_run_small(lsx_ctx* ctx, const ElimJob& job, int* handled);
// Fused sub-warp kernel (row per lane, all primes + CRT in one launch) for m <= 32, n <= 33.
int lsx_run_subwarp(lsx_ctx* ctx, const ElimJob& job, int* handled);
// det(A) mod p for table primes [prime_begin, prime_begin + count) through the tile kernel
bool lsx_tile_fits(int m, int n);
int lsx_tile_det_residues(lsx_ctx* ctx, const int32_t* dA, int n, int prime_begin, int count, uint32_t* d_res);

// blocked modular LU in global memory for n beyond the shared-memory tile
int lsx_blocked_det_residues(lsx_ctx* ctx, const int32_t* dA, int n, int prime_begin, int count, uint32_t* d_res);

int lsx_ws_reserve(lsx_ctx* ctx, size_t bytes);
void lsx_multi_release(lsx_ctx* ctx);   // frees the NCCL state of a multi-device ctx (lsx_multi.cu)
int lsx_fail(lsx_ctx* ctx, int code, const char* fmt, ...);
#define LSX_CUDA_TRY(ctx, expr)                                                              \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return lsx_fail((ctx), LSX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,             \
                            cudaGetErrorString(e__), __FILE__, __LINE__);                    \
    } while (0)
