// k_inv_tpm<N, HEAD, I8>: fused inverse + determinant of n x n matrices (n <= 8), ONE launch per batch, one
// thread per matrix, the matrix resident in N*N registers (reference semantics: linalg.py:682-743 through
// [A|I] row_reduce(bar_col = n), results as adjugate + determinant).
//
// TWO arithmetics share the kernel frame (tile staging, transpose, pivot window, output convention):
//   XS == 3 (the default dispatch): exact integers, one-step fraction-free (Bareiss) Gauss-Jordan -- see
//            tpm_eliminate_bareiss below; 126 us per 2^20 8 x 8 matrices on B200, 0.70 of the measured HBM peak.
//   XS <= 2 (LSX_TPM_ALGO=mont): residues modulo one prime, described next; 198 us.
//
// Arithmetic (mirror: tests/device_model.py::inverse_inplace_v2): in-place division-free Gauss-Jordan modulo ONE
// 31-bit prime.  The launcher takes this path only when the Hadamard bound of every minor of A is below the
// prime, so zero tests modulo p are exact and the adjugate entries come back exactly by the symmetric lift.
// The first HEAD pivot steps run on plain int32 (two IMADs per entry) while the entries provably fit; pivot
// rows stay unscaled and the per-row factor, the sign and the single modular inversion are folded into the
// multipliers of the LAST pivot step.
//
// What this version does to the instruction stream (round 2; ncu of the previous version: 31 % of the executed
// instructions were selects / moves / compares, profiles/r01b_ncu_full_k_inv_tpm8_v2.txt):
//   * The adjugate and the determinant do not depend on the pivot order, so the kernel does not reproduce the
//     reference's "first non-zero row" search with a full physical row swap per step.  It tests the pivot
//     position and the row below it (a 2-row window: 16 selects per step) and only when a whole warp vote finds
//     a lane whose window is all zero does it fall into the full search (a rarely taken, warp-uniform branch).
//   * It eliminates on the TRANSPOSE (held in registers: a renaming at load time).  A row exchange of the
//     transpose permutes the ROWS of the result, so every result row is written with 16-byte shared-memory
//     stores at a data-dependent row address instead of 4-byte scatters at data-dependent columns, and the
//     determinant identity det = sum_c adj[i][c] A[c][i] uses the first pivot row, which is already in registers.
//   * Tiles are staged with 16-byte shared-memory accesses on a layout whose 16-byte chunk stride per matrix is
//     odd (conflict-free for the per-matrix vector reads and for the cooperative copies).
//   * The declared-magnitude check is a min/max scan (3-input min/max), singular and out-of-bound matrices are
//     zeroed in a rarely taken branch instead of per-element selects, the negated multiplier needs no zero
//     test (y = p - f is a valid operand of the two-product reduction), and the determinant, which fits int32 by
//     the launcher's bound, is accumulated with wrapping 32-bit multiply-adds against a column of A that is read
//     back from the input tile (8 registers less while the elimination runs).
// Measured (profiles/r02b_*): 3 063 warp instructions per 32 matrices (was 4 760), 202 us per 2^20 matrices (was
// 263); the kernel is now bound by the fmaheavy pipe (IMAD / IMAD.WIDE / IMAD.HI), 75 % busy.
#pragma once
#include "lsx_internal.h"

namespace lsx_inv_small {

#ifndef LSX_TPM_THREADS
#define LSX_TPM_THREADS 128
#endif
#ifndef LSX_TPM_MINB
#define LSX_TPM_MINB 4
#endif
#ifndef LSX_TPM_WINDOW
#define LSX_TPM_WINDOW 1          // rows below the pivot position tested before the full search
#endif
#ifndef LSX_TPM_X
#define LSX_TPM_X 0               // experiment switch of the lab harness (1: inversion in the last step, as before)
#endif
constexpr int TPM_THREADS = LSX_TPM_THREADS;

// Shared-memory tile: one matrix = ST 32-bit words.  With E = N*N a multiple of 4 the matrix is E/4 chunks of
// 16 bytes and ST/4 is odd (lanes t .. t+7 of a quarter warp then hit 8 different 16-byte bank groups);
// otherwise ST is odd and all accesses are scalar.
template <int N>
struct TpmTile {
    static constexpr int E = N * N;
    static constexpr bool VEC = (E % 4) == 0;
    static constexpr int C = E / 4;
    static constexpr int ST = VEC ? 4 * (C | 1) : (E | 1);
    // int8 input tile: bytes per matrix, 16-byte chunks with an odd chunk stride when E is a multiple of 16
    static constexpr bool VEC8 = (E % 16) == 0;
    static constexpr int C8 = E / 16;
    static constexpr int STB = VEC8 ? 16 * (C8 | 1) : E;
    static constexpr size_t BYTES = (size_t)TPM_THREADS * ST * 4;
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

// residue word of a signed value with |a| < p: a >= 0 -> a, a < 0 -> a + p (one add-min instruction)
__device__ __forceinline__ uint32_t word_of_small(uint32_t a, uint32_t p) { return min(a, a + p); }

// conditional exchange of two register rows (selects: the rows must stay in registers)
template <int N>
__device__ __forceinline__ void cswap_rows(bool sw, uint32_t (&x)[N], uint32_t (&y)[N], uint32_t& px, uint32_t& py) {
#pragma unroll
    for (int c = 0; c < N; ++c) {
        const uint32_t a = x[c], b = y[c];
        x[c] = sw ? b : a;
        y[c] = sw ? a : b;
    }
    const uint32_t a = px, b = py;
    px = sw ? b : a;
    py = sw ? a : b;
}

// exact integer (as double) -> residue in [0, p): q = rint(t / p) by a fused multiply-add against 1.5 * 2^52, r = t - q p
// exactly (|t| < 2^53, q < 2^22), integer conversion by the same constant, sign fix by one add-min
__device__ __forceinline__ uint32_t residue_of_double(double t, double pd, double pinvd, uint32_t p) {
    const double MAGIC = 6755399441055744.0;
    const double q = __fma_rn(t, pinvd, MAGIC) - MAGIC;
    const double r = __fma_rn(-q, pd, t);                   // |r| <= p / 2 (+ p when q is off by one): below p
    const uint32_t ri = (uint32_t)__double2loint(r + MAGIC);
    return min(ri, ri + p);
}
// small signed integer held in a 32-bit register -> double, without the conversion pipe
__device__ __forceinline__ double double_of_int(uint32_t w) {
    return __hiloint2double(0x43300000, (int)(w ^ 0x80000000u)) - 4503601774854144.0;   // 2^52 + 2^31
}

// exact 64-bit integer (made non-negative by a multiple of p) -> residue in [0, p) for p = 2^31 - 1: 2^31 = 1 mod p, so
// t = hi * 2^31 + lo folds to hi + lo; with t < 2^55 the sum is below 2 p and one add-min finishes.  No multiply.
__device__ __forceinline__ uint32_t mersenne31_of_u64(uint64_t t) {
    const uint32_t lo = (uint32_t)t & 0x7fffffffu;
    const uint32_t hi = (uint32_t)(t >> 31);
    const uint32_t s = lo + hi;
    return min(s, s - 0x7fffffffu);
}

// The elimination proper, on a matrix (its TRANSPOSE) that sits in registers: in-place division-free Gauss-Jordan as
// described at the top of this file.  On return W holds the residues of the transposed adjugate with its columns in
// pivot order (slot j = row perm[j] of the adjugate of A), `singular` says whether a pivot was missing, and with
// KEEP_A0 (needs HEAD > 0) a0 is the first pivot row as loaded, i.e. column perm[0] of A.
template <int N, int HEAD, int XS, bool KEEP_A0>
__device__ __forceinline__ void tpm_eliminate(uint32_t (&W)[N][N], uint32_t (&perm)[N], uint32_t (&a0)[N], bool& singular,
                                              const PrimeRec& P) {
    static_assert(!KEEP_A0 || HEAD > 0, "the first pivot row is only an integer row when the head has a step");
    const uint32_t p = P.p, pinv = P.pinv;
#pragma unroll
    for (int r = 0; r < N; ++r) perm[r] = (uint32_t)r;
    bool neg = false;
    singular = false;
    uint32_t cw[N];                         // per-row multiplier words (what pivot row k still lacks)
    uint32_t sig = 1u;                      // head: product of the pivots so far (exact integer)
    uint32_t S = 1u, Q = P.one;
    // F64: pivot step HEAD is still exact integer arithmetic, done on the otherwise idle FP64 pipe (the launcher
    // guarantees 2 B^2 < 2^53 for the bound B of the entries after HEAD integer steps); its results are reduced to
    // residue words on the way out, so the Montgomery steps start one step later: HEADX steps carry no factor R^-1.
    constexpr bool F64 = XS != 0;                       // one more exact step before the Montgomery words start
    constexpr int HEADX = HEAD + (F64 ? 1 : 0);
    static_assert(!F64 || (HEAD >= 1 && HEADX <= N - 1), "the FP64 step needs an integer head before and a last step after it");
    constexpr bool EARLY_INV = LSX_TPM_X != 1 && N >= 2 && N - 2 >= HEADX;   // step N-2 runs on residue words
    uint32_t qinv_pre = 0u;

#pragma unroll
    for (int j = 0; j < N; ++j) {
        const bool head = j < HEAD;
        const bool last = j == N - 1;
        if (j == HEADX) {
            // ---- switch to residue words: S = sigma_h, Q = word(prod sigma_k), cw[k] = word(sigma_k) ----
            // cw[k] was stored as the plain integer sigma_k: to Montgomery form, and Q = their product
            uint32_t qh = P.one;
#pragma unroll
            for (int k = 0; k < HEADX; ++k) {
                cw[k] = mont_mul(word_of_small(cw[k], p), P.r2, p, pinv);
                qh = k == 0 ? cw[0] : mont_mul(qh, cw[k], p, pinv);
            }
            if (F64) {
                // every row but the pivot row of the FP64 step already holds residue words; S was set there
#pragma unroll
                for (int c = 0; c < N; ++c) W[HEAD][c] = word_of_small(W[HEAD][c], p);
            } else {
#pragma unroll
                for (int r = 0; r < N; ++r)
#pragma unroll
                    for (int c = 0; c < N; ++c) W[r][c] = word_of_small(W[r][c], p);
                S = word_of_small(sig, p);
            }
            Q = qh;
        }
        // ---- pivot: position j, else a row of the window below it, else (rare, warp vote) any lower row ----
        if (j + 1 < N) {
            bool found = W[j][j] != 0u;
            const int WEND = j + LSX_TPM_WINDOW < N - 1 ? j + LSX_TPM_WINDOW : N - 1;       // last row of the window
            bool swn[LSX_TPM_WINDOW > 0 ? LSX_TPM_WINDOW : 1];
#pragma unroll
            for (int r = j + 1; r <= WEND; ++r) {
                swn[r - j - 1] = !found && W[r][j] != 0u;
                found = found || swn[r - j - 1];
            }
            if (WEND < N - 1) {
                if (__any_sync(0xffffffffu, !found)) {
                    int src = -1;
#pragma unroll
                    for (int r = N - 1; r > WEND; --r)
                        if (W[r][j] != 0u) src = r;
                    const bool far = !found && src >= 0;
#pragma unroll
                    for (int r = WEND + 1; r < N; ++r) cswap_rows<N>(far && r == src, W[j], W[r], perm[j], perm[r]);
                    neg = neg != far;
                }
            }
#pragma unroll
            for (int r = j + 1; r <= WEND; ++r) {
                cswap_rows<N>(swn[r - j - 1], W[j], W[r], perm[j], perm[r]);
                neg = neg != swn[r - j - 1];
            }
        }
        const uint32_t piv = W[j][j];
        singular = singular || piv == 0u;   // keep going on garbage: every operation below is total
        uint32_t prow[N];
#pragma unroll
        for (int c = 0; c < N; ++c) prow[c] = W[j][c];
        if (KEEP_A0 && j == 0) {               // the first pivot row as loaded = column perm[0] of A (HEAD > 0: integers)
#pragma unroll
            for (int c = 0; c < N; ++c) a0[c] = prow[c];
        }
        if (XS == 2 && j == HEAD) {
            // exact 64-bit integers: t = piv * W[r][c] - f * prow[c] (|t| < 2^53 by the launcher's bound) plus 2^23 p to
            // make it non-negative, folded modulo p = 2^31 - 1 without a multiply: the step costs two IMAD.WIDE per
            // entry on the fmaheavy pipe instead of two IMAD.WIDE + IMAD + IMAD.HI, and carries no factor R^-1
            const int64_t off = (int64_t)0x7fffffff << 23;
            const int32_t pivs = (int32_t)piv, sigs = (int32_t)sig;
#pragma unroll
            for (int r = 0; r < N; ++r) {
                if (r == j) continue;
                const int32_t nf = -(int32_t)W[r][j];
#pragma unroll
                for (int c = 0; c < N; ++c) {
                    const int64_t t = (c == j) ? (int64_t)nf * sigs + off
                                               : (int64_t)nf * (int32_t)prow[c] + ((int64_t)pivs * (int32_t)W[r][c] + off);
                    W[r][c] = mersenne31_of_u64((uint64_t)t);
                }
            }
            W[j][j] = sig;                                       // the pivot row stays integer until the switch
            cw[j] = sig;
            S = mersenne31_of_u64((uint64_t)((int64_t)sigs * pivs + off));   // sigma_{HEAD+1} as a residue word
        } else if (XS == 1 && j == HEAD) {
            // exact integers through the FP64 pipe: t = piv * W[r][c] - f * prow[c] (|t| < 2^53), then t mod p
            const double pd = (double)p, pinvd = 1.0 / pd;
            const double pivd = double_of_int(piv), sigd = double_of_int(sig);
            double prd[N];
#pragma unroll
            for (int c = 0; c < N; ++c) prd[c] = double_of_int(prow[c]);
#pragma unroll
            for (int r = 0; r < N; ++r) {
                if (r == j) continue;
                const double nf = -double_of_int(W[r][j]);
#pragma unroll
                for (int c = 0; c < N; ++c) {
                    const double t = (c == j) ? nf * sigd : __fma_rn(nf, prd[c], pivd * double_of_int(W[r][c]));
                    W[r][c] = residue_of_double(t, pd, pinvd, p);
                }
            }
            W[j][j] = sig;                                       // the pivot row stays integer until the switch
            cw[j] = sig;
            S = residue_of_double(sigd * pivd, pd, pinvd, p);    // sigma_{HEAD+1} as a residue word
        } else if (head) {
            // plain two's-complement integers: W[r][c] = piv * W[r][c] - f * prow[c]
            const uint32_t nsig = 0u - sig;
#pragma unroll
            for (int r = 0; r < N; ++r) {
                if (r == j) continue;
                const uint32_t f = W[r][j];
#pragma unroll
                for (int c = 0; c < N; ++c) W[r][c] = (c == j) ? f * nsig : (piv * W[r][c] - f * prow[c]);
            }
            W[j][j] = sig;
            cw[j] = sig;
            sig *= piv;
        } else {
            cw[j] = S;
            Q = mont_mul(Q, S, p, pinv);
            // The one modular inversion is a chain of 38 dependent products.  Its argument, the product of all
            // sigma_k, only needs the pivots up to step N-2, and the sign cannot change any more either (the last
            // step has no row below it to exchange with): started HERE, in the second-to-last step, the chain
            // interleaves with that step's 56 independent row updates instead of stalling the warp on its own.
            if (EARLY_INV && j == N - 2) {
                const uint32_t s_next = mont_mul(S, piv, p, pinv);            // sigma_{N-1}
                qinv_pre = mont_inverse(mont_mul(Q, s_next, p, pinv), P);
                if (neg) qinv_pre = p - qinv_pre;
            }
            uint32_t qinv = 0u;
            if (last) {
                qinv = EARLY_INV ? qinv_pre : mont_inverse(Q, P);
                if (!EARLY_INV && neg) qinv = p - qinv;   // Q is a unit unless the matrix is singular (discarded)
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
                    uint32_t y = p - W[r][j];           // in [1, p]: a valid operand of the two-product reduction
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

// ---- the same elimination over the INTEGERS: one-step fraction-free (Bareiss) Gauss-Jordan, in place ---------------
// Every entry the fraction-free Gauss-Jordan of [M | I] ever holds is a minor of M (Bareiss), and the launcher takes
// the fused path only when every minor of A is below 2^31 -- so the whole elimination fits two's-complement 32-bit
// registers and needs no prime at all.  Step j with pivot piv and previous pivot d (d = 1 before step 0):
//     rows r != j:   W[r][c] <- (piv * W[r][c] - W[r][j] * W[j][c]) / d   (c != j),   W[r][j] <- -W[r][j]
//     pivot row j:   W[j][j] <- d, the rest stays
// and after the last step W is the adjugate (of the row-permuted matrix) and the last pivot its determinant: no
// Montgomery form, no symmetric lift, no modular inversion (the 38 dependent products of the residue kernel), and
// the determinant is free.  The division is EXACT, so it is a multiplication modulo 2^32: with d = 2^s * o, o odd,
//     q = ((t >> s) * o^-1) mod 2^32        (|q| < 2^31: the low word IS the quotient)
// where o^-1 mod 2^32 comes from three Newton steps (6 dependent IMADs per pivot step, computed one step ahead of
// its use).  t = piv * w - f * prow fits 32 bits in the first H32 steps (launcher: 2 M^2 < 2^31 for the Hadamard
// bound M of the minors that step combines): IMAD, IMAD, SHF, IMAD per entry.  Later steps form t in 64 bits and
// take bits [s, s + 32) with one funnel shift: IMAD.WIDE, IMAD.WIDE, SHF, IMAD -- 5.6 issue slots of the fmaheavy
// pipe per entry where the two-product Montgomery update needs 7.9 (2 IMAD.WIDE + IMAD + IMAD.HI), and 3 in the
// 32-bit steps.  Pivoting (window + rare warp-voted search), the row permutation and the output convention are the
// residue kernel's: slot j = row perm[j] of the adjugate of A, the sign of the exchanges folded into the last step.
__device__ __forceinline__ int64_t mul_wide_s32(uint32_t a, uint32_t b) {
    int64_t r;
    asm("mul.wide.s32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ int64_t mad_wide_s32(uint32_t a, uint32_t b, int64_t c) {
    int64_t r;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
}
// o^-1 modulo 2^32 for odd o: (3 o) xor 2 is right to 5 bits, every Newton step doubles that
__device__ __forceinline__ uint32_t inv_odd_u32(uint32_t o) {
    uint32_t x = (o * 3u) ^ 2u;
    x *= 2u - o * x;
    x *= 2u - o * x;
    x *= 2u - o * x;
    return x;
}

template <int N, int H32, int H53 = H32>
__device__ __forceinline__ void tpm_eliminate_bareiss(uint32_t (&W)[N][N], uint32_t (&perm)[N], bool& singular, uint32_t& det) {
#pragma unroll
    for (int r = 0; r < N; ++r) perm[r] = (uint32_t)r;
    bool neg = false;
    singular = false;
    uint32_t dprev = 1u, dinv = 1u;         // previous pivot d = 2^dsh * o and o^-1 mod 2^32
    int dsh = 0;
    uint32_t piv = 0u;
    // Steps H32 <= j < H53 run on the FP64 pipe, which is idle otherwise: there |piv * w| and |f * prow| stay below 2^53
    // (launcher), so t is exact in a double, and q = t / d comes from ONE fused multiply-add against 1.5 * 2^52 with
    // the reciprocal of d: t * fl(1/d) is within 2^-21 of the integer q (|q| < 2^31), the magic constant rounds it
    // there and leaves q in two's complement in the low word of the sum.  DADD (word -> double), DMUL, DFMA, DFMA per
    // entry instead of IMAD.WIDE, IMAD.WIDE, SHF, IMAD.
    constexpr double MAGIC = 6755399441055744.0;
    double rd = 1.0;                        // fl(1 / d)
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const bool last = j == N - 1;
        // ---- pivot: position j, else a row of the window below it, else (rare, warp vote) any lower row ----
        if (j + 1 < N) {
            bool found = W[j][j] != 0u;
            const int WEND = j + LSX_TPM_WINDOW < N - 1 ? j + LSX_TPM_WINDOW : N - 1;       // last row of the window
            bool swn[LSX_TPM_WINDOW > 0 ? LSX_TPM_WINDOW : 1];
#pragma unroll
            for (int r = j + 1; r <= WEND; ++r) {
                swn[r - j - 1] = !found && W[r][j] != 0u;
                found = found || swn[r - j - 1];
            }
            if (WEND < N - 1) {
                if (__any_sync(0xffffffffu, !found)) {
                    int src = -1;
#pragma unroll
                    for (int r = N - 1; r > WEND; --r)
                        if (W[r][j] != 0u) src = r;
                    const bool far = !found && src >= 0;
#pragma unroll
                    for (int r = WEND + 1; r < N; ++r) cswap_rows<N>(far && r == src, W[j], W[r], perm[j], perm[r]);
                    neg = neg != far;
                }
            }
#pragma unroll
            for (int r = j + 1; r <= WEND; ++r) {
                cswap_rows<N>(swn[r - j - 1], W[j], W[r], perm[j], perm[r]);
                neg = neg != swn[r - j - 1];
            }
        }
        piv = W[j][j];
        singular = singular || piv == 0u;   // keep going on garbage: every operation below is total
        uint32_t prow[N];
#pragma unroll
        for (int c = 0; c < N; ++c) prow[c] = W[j][c];
        const bool flip = last && neg;                          // the sign of the row exchanges, folded into the last step
        const uint32_t minv = flip ? 0u - dinv : dinv;
        const bool f64 = j >= H32 && j < H53 && j > 0 && !last;
        double pivd = 0.0, prd[N];
        if (f64) {
            pivd = double_of_int(piv);
#pragma unroll
            for (int c = 0; c < N; ++c) prd[c] = double_of_int(prow[c]);
        }
#pragma unroll
        for (int r = 0; r < N; ++r) {
            if (r == j) continue;
            const uint32_t f = W[r][j], nf = 0u - f;
            const double nfd = f64 ? double_of_int(nf) : 0.0;
#pragma unroll
            for (int c = 0; c < N; ++c) {
                if (c == j) {
                    W[r][c] = flip ? f : nf;
                } else if (f64) {
                    const double td = __fma_rn(nfd, prd[c], pivd * double_of_int(W[r][c]));
                    W[r][c] = (uint32_t)__double2loint(__fma_rn(td, rd, MAGIC));
                } else if (j == 0) {                            // d = 1: the wrapping 32-bit result is the minor itself
                    const uint32_t t = piv * W[r][c] + nf * prow[c];
                    W[r][c] = flip ? 0u - t : t;
                } else if (j < H32) {                           // t fits 32 bits: exact arithmetic shift, then o^-1
                    const uint32_t t = piv * W[r][c] + nf * prow[c];
                    W[r][c] = (uint32_t)((int32_t)t >> dsh) * minv;
                } else {                                        // t in 64 bits, bits [dsh, dsh + 32) of it, then o^-1
                    const int64_t t = mad_wide_s32(nf, prow[c], mul_wide_s32(piv, W[r][c]));
                    W[r][c] = __funnelshift_r((uint32_t)t, (uint32_t)((uint64_t)t >> 32), dsh) * minv;
                }
            }
        }
        W[j][j] = dprev;
        if (flip) {
#pragma unroll
            for (int c = 0; c < N; ++c) W[j][c] = 0u - W[j][c];
        }
        if (!last) {                                            // the next step divides by this pivot
            dprev = piv;
            dsh = (__ffs((int)piv) - 1) & 31;                   // piv == 0: singular, results are discarded
            dinv = inv_odd_u32((uint32_t)((int32_t)piv >> dsh));
            if (j + 1 >= H32 && j + 1 < H53) rd = 1.0 / (double)(int32_t)piv;
        }
    }
    det = neg ? 0u - piv : piv;
}

// XS: how pivot step HEAD is done when it is still exact integer arithmetic (0: it is an ordinary Montgomery step),
//   1 = on the FP64 pipe, 2 = 64-bit integers folded modulo the Mersenne prime 2^31 - 1 (needs p == 2^31 - 1);
//   3 = the whole elimination over the integers (tpm_eliminate_bareiss, HEAD = its count of 32-bit steps, P unused)
template <int N, int HEAD, bool I8, int XS = 0>
__global__ void __launch_bounds__(TPM_THREADS, LSX_TPM_MINB)
k_inv_tpm(const void* __restrict__ Ain, int64_t batch, PrimeRec P, int a_abs_max, int vec_ok,
          int32_t* __restrict__ adj, int32_t* __restrict__ det, int32_t* __restrict__ status) {
    using T = TpmTile<N>;
    constexpr int E = T::E, ST = T::ST;
    extern __shared__ __align__(16) uint32_t sm[];
    const int tid = threadIdx.x;
    const int64_t tile0 = (int64_t)blockIdx.x * TPM_THREADS;           // first matrix of this block
    const int nmat = (int)min((int64_t)TPM_THREADS, batch - tile0);
    const uint32_t p = P.p, pinv = P.pinv;
    const int me = tid < nmat ? tid : 0;                                // idle lanes of the last block redo matrix 0

    uint32_t W[N][N];          // TRANSPOSE of the matrix.  head: two's-complement integers; tail: residue words
    int vmax = INT32_MIN, vmin = INT32_MAX;

    // ---- coalesced load of the block's matrices into shared memory, then one matrix per thread ----
    if (I8) {
        const int8_t* src = reinterpret_cast<const int8_t*>(Ain) + tile0 * E;
        uint8_t* sm8 = reinterpret_cast<uint8_t*>(sm);
        constexpr int CC8 = T::C8 > 0 ? T::C8 : 1;
        if (T::VEC8 && vec_ok) {
            const int4* src4 = reinterpret_cast<const int4*>(src);
            for (int g = tid; g < nmat * T::C8; g += TPM_THREADS)
                *reinterpret_cast<int4*>(sm8 + (g / CC8) * T::STB + (g % CC8) * 16) = __ldg(src4 + g);
        } else {
            for (int w = tid; w < nmat * E; w += TPM_THREADS) sm8[(w / E) * T::STB + (w % E)] = (uint8_t)__ldg(src + w);
        }
        __syncthreads();
        const uint8_t* mine = sm8 + me * T::STB;
        if constexpr (T::VEC8) {
#pragma unroll
            for (int q = 0; q < T::C8; ++q) {
                const int4 v4 = *reinterpret_cast<const int4*>(mine + 16 * q);
                const int w4[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int e = 16 * q + i;
                    const int v = (int)(int8_t)(w4[i >> 2] >> (8 * (i & 3)));
                    vmax = max(vmax, v);
                    vmin = min(vmin, v);
                    W[e % N][e / N] = (uint32_t)v;
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int v = (int)(int8_t)mine[e];
                vmax = max(vmax, v);
                vmin = min(vmin, v);
                W[e % N][e / N] = (uint32_t)v;
            }
        }
    } else {
        const int32_t* src = reinterpret_cast<const int32_t*>(Ain) + tile0 * E;
        if (T::VEC && vec_ok) {
            constexpr int CC = T::C > 0 ? T::C : 1;
            const int4* src4 = reinterpret_cast<const int4*>(src);
            if (nmat == TPM_THREADS && TPM_THREADS % CC == 0) {
                // full tile: chunk g = tid + k * THREADS lies THREADS / CC matrices further per k, so every address is
                // one base plus a compile-time offset (C loads and stores, no index arithmetic in the loop)
                const int4* s0 = src4 + tid;
                int4* d0 = reinterpret_cast<int4*>(sm + (tid / CC) * ST + (tid % CC) * 4);
#pragma unroll
                for (int k = 0; k < T::C; ++k) d0[k * (TPM_THREADS / CC) * (ST / 4)] = __ldg(s0 + k * TPM_THREADS);
            } else {
#pragma unroll 4
                for (int g = tid; g < nmat * T::C; g += TPM_THREADS)
                    *reinterpret_cast<int4*>(sm + (g / CC) * ST + (g % CC) * 4) = __ldg(src4 + g);
            }
        } else {
            for (int w = tid; w < nmat * E; w += TPM_THREADS) sm[(w / E) * ST + (w % E)] = (uint32_t)__ldg(src + w);
        }
        __syncthreads();
        const uint32_t* mine = sm + me * ST;
        if (T::VEC) {
#pragma unroll
            for (int q = 0; q < T::C; ++q) {
                const int4 v4 = *reinterpret_cast<const int4*>(mine + 4 * q);
                const int w4[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int e = 4 * q + i;
                    vmax = max(vmax, w4[i]);
                    vmin = min(vmin, w4[i]);
                    W[e % N][e / N] = (uint32_t)w4[i];
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int v = (int)mine[e];
                vmax = max(vmax, v);
                vmin = min(vmin, v);
                W[e % N][e / N] = (uint32_t)v;
            }
        }
    }
    // entries outside the declared magnitude: the integer head may wrap and the words may leave [0, p); all of
    // that is well-defined unsigned arithmetic whose results are discarded below (status LSX_ST_BOUND)
    const bool bound_bad = vmax > a_abs_max || vmin < -a_abs_max;

    uint32_t perm[N];                       // perm[r]: original index of the transpose row now at position r
    uint32_t a0[N];
    bool singular;
    uint32_t dsum = 0u;
    if constexpr (XS == 3 || XS == 4) {
        if constexpr (XS == 4) tpm_eliminate_bareiss<N, HEAD, (HEAD + 2 < N - 1 ? HEAD + 2 : N - 1)>(W, perm, singular, dsum);
        else tpm_eliminate_bareiss<N, HEAD>(W, perm, singular, dsum);
        if (singular || bound_bad) {            // rare: zeros (the reference returns NoSolution(), linalg.py:725-737)
            dsum = 0u;
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int c = 0; c < N; ++c) W[r][c] = 0u;
        }
    } else {
    tpm_eliminate<N, HEAD, XS, false>(W, perm, a0, singular, P);
    // Column perm[0] of A (= the first pivot row of the transpose as loaded) for the determinant identity below:
    // read back from the input tile, which is intact until the barrier, instead of living in 8 registers.
    if (I8) {
        const int8_t* mine = reinterpret_cast<const int8_t*>(sm) + me * T::STB;
#pragma unroll
        for (int c = 0; c < N; ++c) a0[c] = (uint32_t)(int)mine[c * N + perm[0]];
    } else {
        const uint32_t* mine = sm + me * ST;
#pragma unroll
        for (int c = 0; c < N; ++c) a0[c] = mine[c * N + perm[0]];
    }

    // ---- symmetric lift, determinant, rows of the adjugate to shared memory at their permuted position ----
    if (singular || bound_bad) {            // rare: zeros (the reference returns NoSolution(), linalg.py:725-737)
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) W[r][c] = 0u;
    }
    const uint32_t half = p >> 1;
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
        for (int c = 0; c < N; ++c) {
            const uint32_t v = W[r][c];
            W[r][c] = v > half ? v - p : v;
        }
    // det = sum_c adj[i][c] * A[c][i] with i = perm[0]: adj[i][c] = W[c][0], A[c][i] = a0[c].  The launcher's bound
    // (|det| < 2^31) makes the wrapping 32-bit sum exact.
#pragma unroll
    for (int c = 0; c < N; ++c) dsum += a0[c] * W[c][0];
    }

    __syncthreads();                        // everybody has read its input tile: reuse it for the output
    {
        uint32_t* mine = sm + me * ST;
        if (tid < nmat) {
#pragma unroll
            for (int j = 0; j < N; ++j) {
                // slot j holds row perm[j] of the adjugate of A: its entries are W[0..N-1][j]
                uint32_t* orow = mine + perm[j] * N;
                if (N % 4 == 0) {
#pragma unroll
                    for (int q = 0; q < N / 4; ++q)
                        *reinterpret_cast<uint4*>(orow + 4 * q) =
                            make_uint4(W[4 * q][j], W[4 * q + 1][j], W[4 * q + 2][j], W[4 * q + 3][j]);
                } else if (N % 2 == 0) {
#pragma unroll
                    for (int q = 0; q < N / 2; ++q)
                        *reinterpret_cast<uint2*>(orow + 2 * q) = make_uint2(W[2 * q][j], W[2 * q + 1][j]);
                } else {
#pragma unroll
                    for (int r = 0; r < N; ++r) orow[r] = W[r][j];
                }
            }
            det[tile0 + tid] = (int32_t)dsum;
            status[tile0 + tid] = (singular && !bound_bad ? LSX_ST_SINGULAR : 0) | (bound_bad ? LSX_ST_BOUND : 0);
        }
    }
    __syncthreads();
    // ---- coalesced store of the adjugates ----
    {
        int32_t* dst = adj + tile0 * E;
        if (T::VEC && vec_ok) {
            constexpr int CC = T::C > 0 ? T::C : 1;
            int4* dst4 = reinterpret_cast<int4*>(dst);
            if (nmat == TPM_THREADS && TPM_THREADS % CC == 0) {
                int4* o0 = dst4 + tid;
                const int4* s0 = reinterpret_cast<const int4*>(sm + (tid / CC) * ST + (tid % CC) * 4);
#pragma unroll
                for (int k = 0; k < T::C; ++k) o0[k * TPM_THREADS] = s0[k * (TPM_THREADS / CC) * (ST / 4)];
            } else {
#pragma unroll 4
                for (int g = tid; g < nmat * T::C; g += TPM_THREADS)
                    dst4[g] = *reinterpret_cast<const int4*>(sm + (g / CC) * ST + (g % CC) * 4);
            }
        } else {
            for (int w = tid; w < nmat * E; w += TPM_THREADS) dst[w] = (int32_t)sm[(w / E) * ST + (w % E)];
        }
    }
}


// ---- streaming form of the same kernel (N * N a multiple of 16 bytes, aligned pointers, HEAD > 0) -------------------
// ncu on k_inv_tpm: a fifth of the stall samples sit on the first shared-memory store behind the tile's global
// loads, i.e. every block waits out the DRAM latency of its own tile with its four warps parked.  Here a block is
// persistent and loops over tiles; the input tile is only needed until every thread holds its matrix in registers, so
// right after that barrier the SAME buffer is refilled with the block's next tile by `cp.async` (LDGSTS: no registers,
// no waiting) while the elimination of the current tile runs -- the load latency disappears behind ~14 us of
// arithmetic without a second buffer.  The price is that the tile buffer cannot double as the result staging area,
// so result rows go straight to global memory as 16-byte stores (two per adjugate row, at the permuted row address;
// the two halves of every 32-byte sector are written by the same warp back to back and merge in L2), and the column
// of A that the determinant identity needs is kept in registers from pivot step 0.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}

template <int N, int HEAD, bool I8, int XS = 0>
__global__ void __launch_bounds__(TPM_THREADS, LSX_TPM_MINB)
k_inv_tpm_stream(const void* __restrict__ Ain, int64_t batch, PrimeRec P, int a_abs_max,
                 int32_t* __restrict__ adj, int32_t* __restrict__ det, int32_t* __restrict__ status) {
    using T = TpmTile<N>;
    constexpr int E = T::E, ST = T::ST;
    static_assert(T::VEC && (!I8 || T::VEC8) && HEAD > 0, "streaming form: vector tiles and an integer head");
    constexpr int CH = I8 ? T::C8 : T::C;                 // 16-byte chunks per matrix in the input container
    extern __shared__ __align__(16) uint32_t sm[];
    const int tid = threadIdx.x;
    const uint32_t p = P.p;
    const int64_t ntiles = (batch + TPM_THREADS - 1) / TPM_THREADS;

    auto issue = [&](int64_t tile) {                      // this thread's share of the tile's chunks, asynchronously
        const int64_t t0 = tile * TPM_THREADS;
        const int nm = (int)min((int64_t)TPM_THREADS, batch - t0);
        const char* src = reinterpret_cast<const char*>(Ain) + t0 * E * (I8 ? 1 : 4);
        char* dst = reinterpret_cast<char*>(sm);
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const int g = tid + k * TPM_THREADS;          // chunk g = matrix g / CH, chunk g % CH of it
            if (g < nm * CH) cp_async16(dst + (g / CH) * (I8 ? T::STB : ST * 4) + (g % CH) * 16, src + (size_t)g * 16);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int64_t tile = blockIdx.x;
    if (tile < ntiles) issue(tile);
    for (; tile < ntiles; tile += gridDim.x) {
        const int64_t tile0 = tile * TPM_THREADS;
        const int nmat = (int)min((int64_t)TPM_THREADS, batch - tile0);
        const int me = tid < nmat ? tid : 0;              // idle lanes of the last tile redo matrix 0
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                  // the tile is in shared memory

        uint32_t W[N][N];                                 // TRANSPOSE of the matrix
        int vmax = INT32_MIN, vmin = INT32_MAX;
        if (I8) {
            const uint8_t* mine = reinterpret_cast<const uint8_t*>(sm) + me * T::STB;
#pragma unroll
            for (int q = 0; q < T::C8; ++q) {
                const int4 v4 = *reinterpret_cast<const int4*>(mine + 16 * q);
                const int w4[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int e = 16 * q + i;
                    const int v = (int)(int8_t)(w4[i >> 2] >> (8 * (i & 3)));
                    vmax = max(vmax, v);
                    vmin = min(vmin, v);
                    W[e % N][e / N] = (uint32_t)v;
                }
            }
        } else {
            const uint32_t* mine = sm + me * ST;
#pragma unroll
            for (int q = 0; q < T::C; ++q) {
                const int4 v4 = *reinterpret_cast<const int4*>(mine + 4 * q);
                const int w4[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int e = 4 * q + i;
                    vmax = max(vmax, w4[i]);
                    vmin = min(vmin, w4[i]);
                    W[e % N][e / N] = (uint32_t)w4[i];
                }
            }
        }
        const bool bound_bad = vmax > a_abs_max || vmin < -a_abs_max;
        __syncthreads();                                  // every thread holds its matrix: the buffer is free again
        if (tile + gridDim.x < ntiles) issue(tile + gridDim.x);

        uint32_t perm[N], a0[N];
        bool singular;
        tpm_eliminate<N, HEAD, XS, true>(W, perm, a0, singular, P);

        if (singular || bound_bad) {                      // rare: zeros (the reference returns NoSolution())
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int c = 0; c < N; ++c) W[r][c] = 0u;
        }
        const uint32_t half = p >> 1;
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) {
                const uint32_t v = W[r][c];
                W[r][c] = v > half ? v - p : v;
            }
        uint32_t dsum = 0u;                               // det = sum_c adj[i][c] * A[c][i], i = perm[0] (wrapping: exact)
#pragma unroll
        for (int c = 0; c < N; ++c) dsum += a0[c] * W[c][0];
        if (tid < nmat) {
            uint32_t* mine = reinterpret_cast<uint32_t*>(adj) + (tile0 + tid) * E;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                uint32_t* orow = mine + perm[j] * N;      // slot j = row perm[j] of the adjugate of A
#pragma unroll
                for (int q = 0; q < N / 4; ++q)
                    *reinterpret_cast<uint4*>(orow + 4 * q) =
                        make_uint4(W[4 * q][j], W[4 * q + 1][j], W[4 * q + 2][j], W[4 * q + 3][j]);
            }
            det[tile0 + tid] = (int32_t)dsum;
            status[tile0 + tid] = (singular && !bound_bad ? LSX_ST_SINGULAR : 0) | (bound_bad ? LSX_ST_BOUND : 0);
        }
    }
}

}  // namespace lsx_inv_small
