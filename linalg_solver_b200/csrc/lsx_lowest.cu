// Lowest terms on the device (the second half of "CRT plus rational reconstruction", BASELINE.json north star).
//
// The elimination kernels return every rational result as an integer numerator over ONE common denominator per
// matrix (the determinant of the pivot minor).  The reference hands back reduced rationals -- true division on
// `Rational` entries (linalg.py:574) and `sympy.Matrix.inv()` (linalg.py:698-699) -- so the last step of a
// drop-in is p / q = num / den with gcd(p, q) = 1 and q > 0 for every entry.  k_lowest_terms does that for
// multi-limb integers, one thread per entry:
//   * magnitudes and signs are split, powers of two are counted (trailing zeros of both operands);
//   * the odd parts go through the binary gcd (subtract the smaller from the larger, strip the trailing zeros of
//     the difference) -- branch-free per step: compare, select, subtract, funnel shift on LM register limbs, so a
//     warp only diverges in its trip count;
//   * both operands are divided EXACTLY by the odd gcd with Jebelean's low-to-high exact division (one 32-bit
//     inverse modulo 2^32 by Newton's iteration, then L^2 / 2 multiply-subtracts), after a right shift by the
//     common power of two; no long division, no quotient estimates.
// Integers cross the ABI as `limbs` little-endian 32-bit words in two's complement (include/lsx.h).
#include "lsx_internal.h"

namespace {

constexpr int LT_THREADS = 128;

template <int LM>
struct Limbs {
    uint32_t w[LM];
};

template <int LM>
__device__ __forceinline__ bool lt_is_zero(const Limbs<LM>& a) {
    uint32_t acc = 0u;
#pragma unroll
    for (int i = 0; i < LM; ++i) acc |= a.w[i];
    return acc == 0u;
}
template <int LM>
__device__ __forceinline__ bool lt_equal(const Limbs<LM>& a, const Limbs<LM>& b) {
    uint32_t acc = 0u;
#pragma unroll
    for (int i = 0; i < LM; ++i) acc |= a.w[i] ^ b.w[i];
    return acc == 0u;
}
// a < b (unsigned): borrow out of a - b
template <int LM>
__device__ __forceinline__ bool lt_less(const Limbs<LM>& a, const Limbs<LM>& b) {
    uint32_t borrow = 0u;
#pragma unroll
    for (int i = 0; i < LM; ++i) {
        const uint64_t t = (uint64_t)a.w[i] - b.w[i] - borrow;
        borrow = (uint32_t)(t >> 63);
    }
    return borrow != 0u;
}
template <int LM>
__device__ __forceinline__ void lt_sub(Limbs<LM>& a, const Limbs<LM>& b) {
    uint32_t borrow = 0u;
#pragma unroll
    for (int i = 0; i < LM; ++i) {
        const uint64_t t = (uint64_t)a.w[i] - b.w[i] - borrow;
        a.w[i] = (uint32_t)t;
        borrow = (uint32_t)(t >> 63);
    }
}
template <int LM>
__device__ __forceinline__ void lt_negate(Limbs<LM>& a) {
    uint32_t carry = 1u;
#pragma unroll
    for (int i = 0; i < LM; ++i) {
        const uint64_t t = (uint64_t)(~a.w[i]) + carry;
        a.w[i] = (uint32_t)t;
        carry = (uint32_t)(t >> 32);
    }
}
// one step of a right shift: by ctz of the low limb, or by a whole limb when the low limb is zero; returns the bits
// shifted out.  The value must be non-zero.
template <int LM>
__device__ __forceinline__ int lt_strip_step(Limbs<LM>& a) {
    const uint32_t lo = a.w[0];
    if (lo == 0u) {
#pragma unroll
        for (int i = 0; i + 1 < LM; ++i) a.w[i] = a.w[i + 1];
        a.w[LM - 1] = 0u;
        return 32;
    }
    const int s = __ffs((int)lo) - 1;
#pragma unroll
    for (int i = 0; i < LM; ++i) a.w[i] = __funnelshift_r(a.w[i], i + 1 < LM ? a.w[i + 1] : 0u, s);
    return s;
}
template <int LM>
__device__ __forceinline__ int lt_strip(Limbs<LM>& a) {      // all trailing zeros of a non-zero value
    int total = 0;
    while ((a.w[0] & 1u) == 0u) total += lt_strip_step(a);
    return total;
}
template <int LM>
__device__ __forceinline__ void lt_shr(Limbs<LM>& a, int s) {  // any s >= 0
    while (s >= 32) {
#pragma unroll
        for (int i = 0; i + 1 < LM; ++i) a.w[i] = a.w[i + 1];
        a.w[LM - 1] = 0u;
        s -= 32;
    }
    if (s) {
#pragma unroll
        for (int i = 0; i < LM; ++i) a.w[i] = __funnelshift_r(a.w[i], i + 1 < LM ? a.w[i + 1] : 0u, s);
    }
}
// n / g for odd g that divides n exactly (Jebelean): quotient limbs from the low end, each fixed by the inverse of
// g modulo 2^32; n is consumed.
template <int LM>
__device__ __forceinline__ void lt_exact_div(Limbs<LM>& n, const Limbs<LM>& g, Limbs<LM>& q) {
    uint32_t inv = g.w[0];                       // g * inv == 1 mod 2^32 after five Newton steps (3 correct bits to start)
#pragma unroll
    for (int i = 0; i < 5; ++i) inv *= 2u - g.w[0] * inv;
#pragma unroll
    for (int i = 0; i < LM; ++i) {
        const uint32_t qi = n.w[i] * inv;
        q.w[i] = qi;
        uint64_t carry = 0;                      // n -= qi * g << (32 i), limbs >= i only
#pragma unroll
        for (int j = i; j < LM; ++j) {
            const uint64_t prod = (uint64_t)qi * g.w[j - i] + carry;
            const uint32_t lo = (uint32_t)prod;
            carry = (prod >> 32) + (n.w[j] < lo ? 1u : 0u);
            n.w[j] -= lo;
        }
    }
}

template <int LM>
__device__ __forceinline__ void lt_load(const uint32_t* src, int L, Limbs<LM>& a, bool* neg) {
    const uint32_t ext = (src[L - 1] >> 31) ? 0xffffffffu : 0u;
#pragma unroll
    for (int i = 0; i < LM; ++i) a.w[i] = i < L ? src[i] : ext;
    *neg = ext != 0u;
    if (*neg) lt_negate(a);
}
template <int LM>
__device__ __forceinline__ void lt_store(uint32_t* dst, int L, Limbs<LM> a, bool neg) {
    if (neg) lt_negate(a);
#pragma unroll
    for (int i = 0; i < LM; ++i)
        if (i < L) dst[i] = a.w[i];
}

template <int LM>
__global__ void __launch_bounds__(LT_THREADS)
k_lowest_terms(const uint32_t* __restrict__ num, const uint32_t* __restrict__ den, int64_t total, int count, int L,
               uint32_t* __restrict__ p_out, uint32_t* __restrict__ q_out) {
    const int64_t idx = (int64_t)blockIdx.x * LT_THREADS + threadIdx.x;
    if (idx >= total) return;
    const int64_t mat = idx / count;
    Limbs<LM> n, d;
    bool nneg, dneg;
    lt_load<LM>(num + idx * L, L, n, &nneg);
    lt_load<LM>(den + mat * L, L, d, &dneg);
    uint32_t* po = p_out + idx * L;
    uint32_t* qo = q_out + idx * L;
    Limbs<LM> one;
#pragma unroll
    for (int i = 0; i < LM; ++i) one.w[i] = i == 0 ? 1u : 0u;
    if (lt_is_zero(d)) {                         // no denominator (singular / inconsistent matrix): 0 / 0
        Limbs<LM> z;
#pragma unroll
        for (int i = 0; i < LM; ++i) z.w[i] = 0u;
        lt_store<LM>(po, L, z, false);
        lt_store<LM>(qo, L, z, false);
        return;
    }
    if (lt_is_zero(n)) {                         // 0 / d = 0 / 1
        lt_store<LM>(po, L, n, false);
        lt_store<LM>(qo, L, one, false);
        return;
    }
    // odd parts and the common power of two
    Limbs<LM> a = n, b = d;
    const int za = lt_strip(a), zb = lt_strip(b);
    const int k = za < zb ? za : zb;
    while (!lt_equal(a, b)) {                    // both odd: the difference is even and non-zero
        const bool swap = lt_less(a, b);
#pragma unroll
        for (int i = 0; i < LM; ++i) {
            const uint32_t x = a.w[i], y = b.w[i];
            a.w[i] = swap ? y : x;
            b.w[i] = swap ? x : y;
        }
        lt_sub(a, b);
        lt_strip(a);
    }
    // gcd = a << k.  p = (n >> k) / a, q = (d >> k) / a, the sign on p.
    lt_shr(n, k);
    lt_shr(d, k);
    if (!lt_equal(a, one)) {
        Limbs<LM> q;
        lt_exact_div(n, a, q);
        n = q;
        lt_exact_div(d, a, q);
        d = q;
    }
    lt_store<LM>(po, L, n, nneg != dneg);
    lt_store<LM>(qo, L, d, false);
}

template <int LM>
void launch(lsx_ctx* ctx, const uint32_t* num, const uint32_t* den, int64_t total, int count, int L, uint32_t* p,
            uint32_t* q) {
    const unsigned grid = (unsigned)((total + LT_THREADS - 1) / LT_THREADS);
    lsx_timing_begin(ctx);
    k_lowest_terms<LM><<<grid, LT_THREADS, 0, ctx->stream>>>(num, den, total, count, L, p, q);
    lsx_timing_end(ctx);
    ctx->launches++;
}

}  // namespace

extern "C" int lsx_lowest_terms(lsx_ctx* ctx, const uint32_t* num, const uint32_t* den, int64_t batch, int count,
                                int limbs, int mem, uint32_t* p, uint32_t* q) {
    if (!ctx) return LSX_ERR_NULL;
    if (batch < 0 || count < 1 || limbs < 1 || limbs > LSX_MAX_BATCH_PRIMES)
        return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "lowest_terms: needs batch >= 0, count >= 1, 1 <= limbs <= %d",
                        LSX_MAX_BATCH_PRIMES);
    if (mem != LSX_MEM_HOST && mem != LSX_MEM_DEVICE) return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "bad mem flag %d", mem);
    if (batch == 0) return LSX_OK;
    if (!num || !den || !p || !q) return lsx_fail(ctx, LSX_ERR_NULL, "lowest_terms: NULL buffer");
    const int64_t total = batch * count;
    if (total > (int64_t)0x7fffffff * LT_THREADS) return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "lowest_terms: too many entries");
    LSX_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const uint32_t *dn = num, *dd = den;
    uint32_t *dp = p, *dq = q;
    const size_t nb = (size_t)total * limbs * 4, db = (size_t)batch * limbs * 4;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    if (mem == LSX_MEM_HOST) {
        int rc = lsx_ws_reserve(ctx, 3 * up(nb) + up(db));
        if (rc != LSX_OK) return rc;
        char* base = (char*)ctx->d_ws;
        dn = (const uint32_t*)base;
        dp = (uint32_t*)(base + up(nb));
        dq = (uint32_t*)(base + 2 * up(nb));
        dd = (const uint32_t*)(base + 3 * up(nb));
        LSX_CUDA_TRY(ctx, cudaMemcpyAsync((void*)dn, num, nb, cudaMemcpyHostToDevice, ctx->stream));
        LSX_CUDA_TRY(ctx, cudaMemcpyAsync((void*)dd, den, db, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (limbs <= 2) launch<2>(ctx, dn, dd, total, count, limbs, dp, dq);
    else if (limbs <= 4) launch<4>(ctx, dn, dd, total, count, limbs, dp, dq);
    else if (limbs <= 8) launch<8>(ctx, dn, dd, total, count, limbs, dp, dq);
    else if (limbs <= 12) launch<12>(ctx, dn, dd, total, count, limbs, dp, dq);
    else if (limbs <= 20) launch<20>(ctx, dn, dd, total, count, limbs, dp, dq);
    else launch<32>(ctx, dn, dd, total, count, limbs, dp, dq);
    LSX_CUDA_TRY(ctx, cudaGetLastError());
    if (mem == LSX_MEM_HOST) {
        LSX_CUDA_TRY(ctx, cudaMemcpyAsync(p, dp, nb, cudaMemcpyDeviceToHost, ctx->stream));
        LSX_CUDA_TRY(ctx, cudaMemcpyAsync(q, dq, nb, cudaMemcpyDeviceToHost, ctx->stream));
        LSX_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return LSX_OK;
}
