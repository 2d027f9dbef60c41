// Garner CRT of up to KT residues to a signed multi-limb integer (shared by the assemble kernels).
#pragma once
#include "lsx_internal.h"

template <int KT>
__device__ __forceinline__ void crt_limbs(const uint32_t (&r)[KT], const uint8_t* sel, int K,
                                          const PrimeRec* primes, const uint32_t* garner,
                                          uint32_t (&acc)[KT], bool* is_zero) {
    uint32_t v[KT];
    uint32_t pp[KT];
    bool zero = true;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
        v[j] = 0;
        pp[j] = 0;
        if (j < K) {
            const int sj = sel[j];
            const PrimeRec P = primes[sj];
            pp[j] = P.p;
            uint32_t t = r[j];
            zero &= t == 0u;
#pragma unroll
            for (int i = 0; i < j; ++i) {
                uint32_t vi = v[i];
                if (vi >= P.p) vi -= P.p;
                t = t >= vi ? t - vi : t + P.p - vi;
                t = mont_mul(t, garner[(int)sel[i] * LSX_GARNER_DIM + sj], P.p, P.pinv);
            }
            v[j] = t;
        }
    }
    *is_zero = zero;
    // sign: X > (M-1)/2  <=>  mixed-radix digits compare above ((p_i - 1)/2)_i from the top
    bool negv = false, decided = false;
#pragma unroll
    for (int i = KT - 1; i >= 0; --i) {
        if (i < K && !decided) {
            uint32_t h = (pp[i] - 1u) >> 1;
            if (v[i] != h) {
                negv = v[i] > h;
                decided = true;
            }
        }
    }
    // negative: X - M = -(Y + 1) with Y = sum (p_i - 1 - v_i) P_i, so the result is ~Y
    if (negv) {
#pragma unroll
        for (int i = 0; i < KT; ++i)
            if (i < K) v[i] = pp[i] - 1u - v[i];
    }
#pragma unroll
    for (int l = 0; l < KT; ++l) acc[l] = 0u;
#pragma unroll
    for (int i = KT - 1; i >= 0; --i) {
        if (i < K) {
            uint64_t carry = v[i];
#pragma unroll
            for (int l = 0; l < KT - i; ++l) {     // acc < 2^(32 (KT - 1 - i)) before this digit: the limbs above are zero
                uint64_t t = (uint64_t)acc[l] * pp[i] + carry;
                acc[l] = (uint32_t)t;
                carry = t >> 32;
            }
        }
    }
    if (negv) {
#pragma unroll
        for (int l = 0; l < KT; ++l) acc[l] = ~acc[l];
    }
}

template <int KT>
__device__ __forceinline__ void store_limbs(uint32_t* dst, const uint32_t (&acc)[KT], int L, bool negate,
                                            bool zero_out) {
    if (zero_out) {
        for (int l = 0; l < L; ++l) dst[l] = 0u;
        return;
    }
    if (!negate) {
#pragma unroll
        for (int l = 0; l < KT; ++l)
            if (l < L) dst[l] = acc[l];
    } else {
        uint32_t carry = 1u;
#pragma unroll
        for (int l = 0; l < KT; ++l) {
            if (l < L) {
                uint32_t x = ~acc[l];
                uint32_t y = x + carry;
                carry = (y < x) ? 1u : 0u;
                dst[l] = y;
            }
        }
    }
}

