// Fused sub-warp kernel for batches of small matrices (m <= 32 rows, n <= 33 columns): ONE launch does
// every prime, the pivot-profile agreement check, the Garner CRT and the scatter into the layout of
// the operation (RREF grid / adjugate / particular solution + generators / determinant / rank).
//
//   * G lanes hold one matrix (G = 4, 8, 16 or 32; one ROW per lane, the row's NC words in registers),
//     32/G matrices per warp.  Because rows are lanes, the pivot ROW index may be data dependent
//     (rank-deficient inputs skip columns) while every register index stays static.
//   * pivot search = ballot over "non-zero at or below the pivot position" + ffs: the LOWEST such row,
//     which is the reference's rule (linalg.py:548-567; not a max-magnitude search); row swaps are
//     physical lane exchanges by shuffle, the pivot row is broadcast by shuffle.
//   * per prime: uniform-scale division-free Gauss-Jordan on Montgomery words with one Fermat
//     inversion (mirror: tests/device_model.py::elim_words); residues N = d * RREF go to shared
//     memory; the profile (source row per column) of every prime is compared with the first prime's.
//   * all K primes agree  =>  the profile is the rational one (their product exceeds the Hadamard
//     bound of every minor) and the matrix is finished here; otherwise it is
//     appended to a retry list and recomputed by the tile path with replacement primes.
// Inputs are read once and outputs written once; residues never touch HBM.
#include <utility>

#include "lsx_crt.cuh"
#include "lsx_internal.h"

namespace {

#ifndef LSX_SW_THREADS
#define LSX_SW_THREADS 128
#endif
#ifndef LSX_SW_SYNC
#define LSX_SW_SYNC 0            // experiment: a CTA barrier every LSX_SW_SYNC pivot steps keeps the warps of a CTA on the same
#endif                           // stretch of the unrolled instruction stream (instruction-cache sharing)
constexpr int SW_THREADS = LSX_SW_THREADS;
constexpr int SW_SKIP = 63;          // profile code of a column without pivot

struct SwArgs {
    const int32_t* A;
    const int32_t* bvec;
    int64_t batch;
    int m, n_in, n, bar, right_identity, op;
    int K, L, gen_cap, pivot_slots, max_rank;
    const int32_t* kword;       // NULL, or kword[1] = primes the batch's own row norms need (lsx_row_bound_primes)
    int a_abs_max, b_abs_max;
    const PrimeRec* primes;
    const uint32_t* garner;
    uint32_t* num;
    uint32_t* den;
    uint32_t* particular;
    uint32_t* generators;
    int32_t* pivot_col;
    int32_t* rank;
    int32_t* status;
    int32_t* retry_list;
    int32_t* retry_count;
    int retry_cap;
};

// Garner CRT of the K residues of one entry (scaled on the fly by the per-prime factor) to a signed
// L-limb integer, stored negated / zeroed on request.  Prime records and Garner inverses come from
// shared memory (loaded once per CTA).  One copy per KT (not inlined: the kernel calls it in a loop).
struct SwTables {
    const PrimeRec* primes;      // [K]            (shared memory)
    const uint32_t* garner;      // [K][K]: (p_i^-1 mod p_j) * R mod p_j
};

template <int KT>
__device__ __noinline__ void crt_entry(const uint32_t* res, int res_stride, const uint32_t* scale, int scale_stride, int K,
                                        int L, SwTables T, uint32_t* dst, bool negate, bool zero_out) {
    uint32_t v[KT], pp[KT], acc[KT];
#pragma unroll
    for (int j = 0; j < KT; ++j) {
        v[j] = 0u;
        pp[j] = 0u;
        if (j < K) {
            const PrimeRec P = T.primes[j];
            pp[j] = P.p;
            uint32_t t = res ? mont_mul(scale[j * scale_stride], res[j * res_stride], P.p, P.pinv) : scale[j * scale_stride];
#pragma unroll
            for (int i = 0; i < j; ++i) {
                uint32_t vi = v[i];
                if (vi >= P.p) vi -= P.p;
                t = t >= vi ? t - vi : t + P.p - vi;
                t = mont_mul(t, T.garner[i * K + j], P.p, P.pinv);
            }
            v[j] = t;
        }
    }
    bool negv = false, decided = false;
#pragma unroll
    for (int i = KT - 1; i >= 0; --i) {
        if (i < K && !decided) {
            const uint32_t h = (pp[i] - 1u) >> 1;
            if (v[i] != h) {
                negv = v[i] > h;
                decided = true;
            }
        }
    }
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
                const uint64_t t = (uint64_t)acc[l] * pp[i] + carry;
                acc[l] = (uint32_t)t;
                carry = t >> 32;
            }
        }
    }
    if (negv) {
#pragma unroll
        for (int l = 0; l < KT; ++l) acc[l] = ~acc[l];
    }
    store_limbs<KT>(dst, acc, L, negate, zero_out);
}

// NOT inlined on purpose: the kernel calls it from ten places and every inlined copy carries three Garner
// instantiations; inlined, k_subwarp<16,17> was 9 500 instructions (152 KB) and ncu's top stall was the
// instruction fetch ("no_instruction").
__device__ __noinline__ void crt_entry_any(const uint32_t* res, int res_stride, const uint32_t* scale, int scale_stride,
                                              int K, int L, SwTables T, uint32_t* dst, bool negate, bool zero_out) {
    // KT covers the limbs as well: with the prime count taken from the data K can be below L (sign extension)
    if (K <= 1 && L <= 1) crt_entry<1>(res, res_stride, scale, scale_stride, K, L, T, dst, negate, zero_out);
    else if (K <= 4 && L <= 4) crt_entry<4>(res, res_stride, scale, scale_stride, K, L, T, dst, negate, zero_out);
    else crt_entry<8>(res, res_stride, scale, scale_stride, K, L, T, dst, negate, zero_out);
}

struct SwState {
    uint32_t S, Q, X;
    uint64_t prof;
    int pi;
    bool neg;
};

// One pivot column J (compile-time, so that every register index is static) for PP primes at once.  The primes of
// a pass are independent dependency chains (ballot -> shuffle of the pivot -> multipliers -> row update -> next
// ballot); a warp that runs them one after the other waits out every link of that chain (ncu, round 1: 12 cycles
// per issued instruction per warp, "wait" and the shuffles' scoreboard on top), interleaved they fill each
// other's latencies.  Every phase of the step is therefore written prime-major.
template <int G, int NC, int PP, int J>
__device__ __forceinline__ void sw_step(uint32_t (&row)[PP][NC], SwState (&st)[PP], const PrimeRec (&P)[PP], int r, int gbase,
                                        unsigned gmask, int m, int bar, bool live) {
    constexpr unsigned FULL = 0xffffffffu;
    if (J >= bar) return;                 // uniform over the grid
    bool has[PP];
    int src[PP], pi[PP];
#pragma unroll
    for (int q = 0; q < PP; ++q) {
        pi[q] = st[q].pi;
        const bool nz = r >= pi[q] && r < m && row[q][J] != 0u;
        const unsigned bal = (__ballot_sync(FULL, nz) >> gbase) & gmask;
        has[q] = bal != 0u && pi[q] < m;
        src[q] = has[q] ? __ffs(bal) - 1 : pi[q];
    }
#pragma unroll
    for (int q = 0; q < PP; ++q) {
        if (__any_sync(FULL, has[q] && src[q] != pi[q])) {
            // physical swap of rows pi and src (lanes of groups without a swap read themselves)
            const int partner = (has[q] && src[q] != pi[q]) ? (r == pi[q] ? src[q] : (r == src[q] ? pi[q] : r)) : r;
#pragma unroll
            for (int c = 0; c < NC; ++c) row[q][c] = __shfl_sync(FULL, row[q][c], gbase + partner);
        }
        st[q].neg ^= has[q] && src[q] != pi[q];
        if (r == J % G) st[q].prof |= (uint64_t)(has[q] ? src[q] : SW_SKIP) << (6 * (J / G));
    }
    bool any_has[PP], fast[PP];
    int pl[PP];
    uint32_t piv[PP], x[PP], y[PP];
#pragma unroll
    for (int q = 0; q < PP; ++q) {
        any_has[q] = __any_sync(FULL, has[q]);            // no group of this warp has a pivot in column J: nothing to do
        pl[q] = gbase + (has[q] ? pi[q] : 0);
        piv[q] = __shfl_sync(FULL, row[q][J], pl[q]);
        const uint32_t f = row[q][J];
        const bool isp = r == pi[q];
        x[q] = has[q] ? (isp ? st[q].S : piv[q]) : P[q].one;
        y[q] = (has[q] && !isp && f) ? P[q].p - f : 0u;
        // While no column has been skipped (pi == J in every group of the warp) the columns left of J are
        // finished pivot columns: they are not maintained any more (their final values are known: d on
        // the pivot row, 0 elsewhere) and only columns > J are updated.
        fast[q] = __all_sync(FULL, !live || (has[q] && pi[q] == J));
    }
    // columns right of J: always live
#pragma unroll
    for (int c = J + 1; c < NC; ++c) {
#pragma unroll
        for (int q = 0; q < PP; ++q) {
            if (any_has[q]) {
                const uint32_t pc = __shfl_sync(FULL, row[q][c], pl[q]);
                row[q][c] = mont_fma2(x[q], row[q][c], y[q], pc, P[q].p, P[q].pinv);
            }
        }
    }
    // columns up to J: only once a column has been skipped somewhere in the warp (rank-deficient input)
#pragma unroll
    for (int q = 0; q < PP; ++q) {
        if (any_has[q] && !fast[q]) {
#pragma unroll
            for (int c = 0; c <= J; ++c) {
                const uint32_t pc = __shfl_sync(FULL, row[q][c], pl[q]);
                row[q][c] = mont_fma2(x[q], row[q][c], y[q], pc, P[q].p, P[q].pinv);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < PP; ++q) {
        if (has[q]) {
            st[q].Q = mont_mul(st[q].Q, st[q].S, P[q].p, P[q].pinv);
            st[q].S = mont_mul(st[q].S, piv[q], P[q].p, P[q].pinv);
            st[q].X = mont_mul(st[q].X, P[q].r2, P[q].p, P[q].pinv);
            ++st[q].pi;
        }
    }
}

template <int G, int NC, int PP, int... Js>
__device__ __forceinline__ void sw_steps(uint32_t (&row)[PP][NC], SwState (&st)[PP], const PrimeRec (&P)[PP], int r, int gbase,
                                         unsigned gmask, int m, int bar, bool live, std::integer_sequence<int, Js...>) {
#if LSX_SW_SYNC > 0
    ((sw_step<G, NC, PP, Js>(row, st, P, r, gbase, gmask, m, bar, live), (Js % LSX_SW_SYNC == 0 ? __syncthreads() : (void)0)), ...);
#else
    (sw_step<G, NC, PP, Js>(row, st, P, r, gbase, gmask, m, bar, live), ...);
#endif
}

// No minimum-CTAs bound on purpose: ptxas takes ~126 registers for the 16 x 17 shape (4 CTAs per SM); forcing 5 CTAs
// (96 registers, 8 bytes of spills) or 6 (80 registers, 40 bytes) measured SLOWER on B200: 2.53 / 2.76 against 2.41 ms
// per 2^18 systems (profiles/r02af_c3_minblocks.txt).
template <int G, int NC, int PP>
__global__ void __launch_bounds__(SW_THREADS) k_subwarp(const SwArgs a) {
    extern __shared__ uint32_t sm[];          // res[K][NC][SW_THREADS] then dres[K][SW_THREADS]
    const int tid = threadIdx.x, lane = tid & 31;
    const int r = lane % G;                   // my row
    const int gbase = lane - r;               // first lane of my group
    const unsigned gmask = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
    constexpr unsigned FULL = 0xffffffffu;
    const int m = a.m, n = a.n, bar = a.bar, K = a.kword ? min(a.K, a.kword[1]) : a.K, L = a.L;
    const int64_t mat = ((int64_t)blockIdx.x * SW_THREADS + tid) / G;
    const bool live = mat < a.batch;          // whole groups are live or not
    const bool rowlive = live && r < m;
    // shared memory: res[K][NC][T] raw words, then per prime Gw / G2w / d / meta [K][4][T] (slot of the
    // group's first lane), then the prime records [K] and the Garner inverses [K][K]
    uint32_t* scal = sm + (size_t)K * NC * SW_THREADS;
    PrimeRec* s_primes = reinterpret_cast<PrimeRec*>(scal + (size_t)K * 4 * SW_THREADS);
    uint32_t* s_garner = reinterpret_cast<uint32_t*>(s_primes + K);
    for (int i = tid; i < K; i += SW_THREADS) s_primes[i] = a.primes[i];
    for (int i = tid; i < K * K; i += SW_THREADS) s_garner[i] = a.garner[(i / K) * LSX_GARNER_DIM + (i % K)];
    __syncthreads();

    // ---- load my row (kept as integers for every prime) ----
    int32_t in[NC];
    bool bad = false;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        int32_t v = 0;
        if (rowlive && c < n) {
            int lim;
            if (c < a.n_in) {
                v = a.A[(mat * m + r) * a.n_in + c];
                lim = c < bar ? a.a_abs_max : a.b_abs_max;
            } else if (a.right_identity) {
                v = (c - a.n_in == r) ? 1 : 0;
                lim = 1;
            } else {
                v = a.bvec[mat * m + r];
                lim = a.b_abs_max;
            }
            bad |= v > lim || v < -lim;
        }
        in[c] = v;
    }
    bad = (__ballot_sync(FULL, bad) >> gbase & gmask) != 0;

    uint64_t prof0 = 0;                       // my share of the first prime's profile
    int rank0 = 0;
    bool mismatch = false;

    for (int k0 = 0; k0 < K; k0 += PP) {
        // PP primes per pass; an odd tail repeats the last prime (its second copy is not stored)
        PrimeRec P[PP];
        uint32_t row[PP][NC];
        SwState st[PP];
#pragma unroll
        for (int q = 0; q < PP; ++q) {
            P[q] = s_primes[k0 + q < K ? k0 + q : K - 1];
#pragma unroll
            for (int c = 0; c < NC; ++c) row[q][c] = word_of_int(in[c], P[q].p);
            st[q] = SwState{P[q].one, P[q].one, 1u, 0ull, 0, false};
        }
        sw_steps<G, NC, PP>(row, st, P, r, gbase, gmask, m, bar, live, std::make_integer_sequence<int, NC>{});
#pragma unroll
        for (int q = 0; q < PP; ++q) {
            const int k = k0 + q;
            if (k >= K) break;
            // ---- raw words to shared memory; the scaling factors follow after the prime loop ----
#pragma unroll
            for (int c = 0; c < NC; ++c) sm[((size_t)k * NC + c) * SW_THREADS + tid] = row[q][c];
            if (r == 0) {
                uint32_t* sc = scal + (size_t)k * 4 * SW_THREADS + tid;
                sc[0] = st[q].Q;
                sc[SW_THREADS] = st[q].X;
                sc[2 * SW_THREADS] = st[q].S;
                sc[3 * SW_THREADS] = (uint32_t)st[q].pi | (st[q].neg ? 256u : 0u);
            }
            if (k == 0) {
                prof0 = st[q].prof;
                rank0 = st[q].pi;
            } else {
                mismatch |= st[q].prof != prof0 || st[q].pi != rank0;
            }
        }
    }
    mismatch = ((__ballot_sync(FULL, mismatch) >> gbase) & gmask) != 0u;
    __syncwarp();

    // ---- one Fermat inversion per (matrix, prime), spread over the lanes of the group ----
    const int gt0 = tid - r;                  // slot of the group's first lane
    for (int k = r; k < K; k += G) {
        const PrimeRec P = s_primes[k];
        const uint32_t p = P.p, pinv = P.pinv;
        uint32_t* sc = scal + (size_t)k * 4 * SW_THREADS + gt0;
        const uint32_t Q = sc[0], X = sc[SW_THREADS], S = sc[2 * SW_THREADS], meta = sc[3 * SW_THREADS];
        uint32_t Gw = mont_mul(mont_pow(Q, p - 2u, P.one, p, pinv), X, p, pinv);
        if ((meta & 256u) && Gw) Gw = p - Gw;
        sc[0] = Gw;                                         // factor of pivot rows:      N = Gw * W
        sc[SW_THREADS] = mont_mul(Gw, P.r2, p, pinv);       // factor of non-pivot rows
        sc[2 * SW_THREADS] = mont_mul(Gw, S, p, pinv);      // d = det of the pivot minor (plain residue)
    }
    __syncwarp();

    // ---- pivot columns from the first prime's profile ----
    unsigned pivmask = 0;                     // bit j: column j holds a pivot (bar <= 32)
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        if (j < bar) {
            const uint64_t share = __shfl_sync(FULL, prof0, gbase + j % G);
            if ((int)((share >> (6 * (j / G))) & 63u) != SW_SKIP) pivmask |= 1u << j;
        }
    }
    const int rank = rank0;
    if (!live) return;
    int st = 0;
    if (bad) st |= LSX_ST_BOUND;
    if (a.max_rank > 0 && rank > a.max_rank) st |= LSX_ST_BOUND;
    if (mismatch && !bad) {
        // a bad prime: hand the matrix to the tile path (replacement primes)
        if (r == 0) {
            const int pos = atomicAdd(a.retry_count, 1);
            if (pos < a.retry_cap) {
                a.retry_list[pos] = (int32_t)mat;
                atomicOr(&a.status[mat], LSX_ST_INTERNAL_RETRY);
            } else {
                atomicOr(&a.status[mat], LSX_ST_NO_GOOD_PRIME);
            }
        }
        return;
    }
    if (r == 0) {
        if (a.rank) a.rank[mat] = rank;
        if (a.pivot_col) {
            int cnt = 0;
            for (int j = 0; j < bar; ++j)
                if (pivmask >> j & 1u) a.pivot_col[mat * a.pivot_slots + cnt++] = j;
            for (; cnt < a.pivot_slots; ++cnt) a.pivot_col[mat * a.pivot_slots + cnt] = -1;
        }
    }
    if (st) {
        if (r == 0) atomicOr(&a.status[mat], st);
        return;
    }
    if (a.op == LSX_OP_RANK) return;

    // ---- CRT + scatter: the entries of a matrix are dealt round-robin to the lanes of its group ----
    const SwTables T{s_primes, s_garner};
    const uint32_t* sc_piv = scal + gt0;                      // [k * 4T]: Gw
    const uint32_t* sc_non = scal + SW_THREADS + gt0;         // G2w
    const uint32_t* sc_d = scal + 2 * SW_THREADS + gt0;       // d
    const int sstr = 4 * SW_THREADS, rstr = NC * SW_THREADS;
    auto res_of = [&](int i, int c) { return sm + (size_t)c * SW_THREADS + gt0 + i; };
    const bool singular = rank < m;
    // Every branch below first picks its operands per lane and then makes ONE convergent call of the Garner routine
    // per round: called from inside the branches, lanes with different kinds of entries (particular solution /
    // generator / denominator copies) ran the 200-instruction routine one kind after the other.
    if (r == 0 && a.op != LSX_OP_SOLVE) {       // a solve folds the denominator into its entry list
        const bool zero_det = (a.op == LSX_OP_INVERSE || a.op == LSX_OP_DET) && singular;
        crt_entry_any(nullptr, 0, sc_d, sstr, K, L, T, a.den + mat * L, false, zero_det);
        if (a.op == LSX_OP_INVERSE && singular) atomicOr(&a.status[mat], LSX_ST_SINGULAR);
    }
    if (a.op == LSX_OP_DET) return;
    if (a.op == LSX_OP_RREF) {
        const int E = m * n;
        for (int e = r; e < E; e += G) {
            const int i = e / n, c = e - i * n;
            uint32_t* dst = a.num + ((mat * m + i) * (int64_t)n + c) * L;
            // finished pivot column: d on its pivot row, 0 elsewhere
            const bool pcol = c < bar && (pivmask >> c & 1u);
            const bool mine = pcol && i < rank && c == (int)__fns(pivmask, 0, i + 1);
            crt_entry_any(pcol ? nullptr : res_of(i, c), rstr, pcol ? sc_d : (i < rank ? sc_piv : sc_non), sstr, K, L, T, dst,
                          false, pcol && !mine);
        }
    } else if (a.op == LSX_OP_INVERSE) {
        const int nn = a.n_in, E = m * nn;
        for (int e = r; e < E; e += G) {
            const int i = e / nn, c = e - i * nn;
            crt_entry_any(res_of(i, nn + c), rstr, i < rank ? sc_piv : sc_non, sstr, K, L, T,
                          a.num + ((mat * m + i) * (int64_t)nn + c) * L, false, singular);
        }
    } else {   // LSX_OP_SOLVE
        const int nvars = n - 1;
        if (r >= rank && r < m) {
            // zero left row: inconsistent iff the rhs is non-zero (linalg.py:913-934); the scale is a unit
            bool zero = true;
            for (int k = 0; k < K; ++k) zero &= sm[((size_t)k * NC + nvars) * SW_THREADS + tid] == 0u;
            if (!zero) atomicOr(&a.status[mat], LSX_ST_INCONSISTENT);
        }
        const unsigned freemask = ~pivmask & (nvars >= 32 ? 0xffffffffu : ((1u << nvars) - 1u));
        const int nfree = nvars - rank;
        const int per = nfree + 1, E = rank * per + nfree;
        // lane t of the group keeps the column of pivot t and the t-th free column (one bit scan each instead of one
        // per entry), G covers both lists twice over when m > G / 2 only for G = 32, where two slots per lane are used
        const int my_pcol0 = r < rank ? (int)__fns(pivmask, 0, r + 1) : 0;
        const int my_fcol0 = r < nfree ? (int)__fns(freemask, 0, r + 1) : 0;
        const int my_fcol1 = r + G < nfree ? (int)__fns(freemask, 0, r + G + 1) : 0;     // nfree can reach 32 > G
        // (i, q) = divmod(e, per) carried along instead of divided out per entry
        const int di = G / per, dq = G - di * per;
        int i = r / per, q = r - i * per;
        const unsigned grp = gmask << gbase;                       // the other group of the warp may have left already
        for (int e0 = 0; e0 <= E; e0 += G) {                       // whole groups iterate together (shuffles inside);
            const int e = e0 + r;                                  // entry E is the common denominator
            const bool on = e < E;
            const bool in_rows = on && e < rank * per;
            const int ii = in_rows ? i : 0, qq = on ? (in_rows ? q : e - rank * per) : 0;
            const int pcol = __shfl_sync(grp, my_pcol0, gbase + (ii < G ? ii : 0));
            const int fsel = qq < nfree ? qq : 0;
            const int fc_lo = __shfl_sync(grp, my_fcol0, gbase + (fsel % G));
            const int fc_hi = __shfl_sync(grp, my_fcol1, gbase + (fsel % G));
            const int fc = fsel < G ? fc_lo : fc_hi;
            const uint32_t* res = nullptr;
            const uint32_t* scl = sc_d;
            uint32_t* dst = nullptr;
            bool negate = false;
            if (in_rows) {
                if (q == nfree) {
                    res = res_of(i, nvars);
                    scl = sc_piv;
                    dst = a.particular + (mat * nvars + pcol) * L;
                } else if (q < a.gen_cap) {
                    res = res_of(i, fc);
                    scl = sc_piv;
                    dst = a.generators + ((mat * nvars + pcol) * (int64_t)a.gen_cap + q) * L;
                    negate = true;
                }
            } else if (on) {
                // generator entries equal to d at the free columns (gen[f] = 1, linalg.py:976)
                if (qq < a.gen_cap) dst = a.generators + ((mat * nvars + fc) * (int64_t)a.gen_cap + qq) * L;
            } else if (e == E) {
                dst = a.den + mat * L;
            }
            if (dst) crt_entry_any(res, rstr, scl, sstr, K, L, T, dst, negate, false);
            i += di;
            q += dq;
            if (q >= per) {
                q -= per;
                ++i;
            }
        }
        if (r == 0 && nfree > a.gen_cap) atomicOr(&a.status[mat], LSX_ST_GEN_TRUNC);
    }
}

size_t sw_smem_bytes(int K, int NC) {
    return ((size_t)K * NC + (size_t)K * 4) * SW_THREADS * 4 + (size_t)K * sizeof(PrimeRec) + (size_t)K * K * 4;
}

template <int G, int NC, int PP>
int launch_sw_pp(lsx_ctx* ctx, const SwArgs& a) {
    const size_t smem = sw_smem_bytes(a.K, NC);
    if (smem > 48 * 1024)
        LSX_CUDA_TRY(ctx, cudaFuncSetAttribute(k_subwarp<G, NC, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t threads = a.batch * G;
    const unsigned grid = (unsigned)((threads + SW_THREADS - 1) / SW_THREADS);
    lsx_timing_begin(ctx);
    k_subwarp<G, NC, PP><<<grid, SW_THREADS, smem, ctx->stream>>>(a);
    lsx_timing_end(ctx);
    ctx->launches++;
    LSX_CUDA_TRY(ctx, cudaGetLastError());
    return LSX_OK;
}

// two primes per pass when the plan has at least two and the row fits the register budget twice
template <int G, int NC>
int launch_sw(lsx_ctx* ctx, const SwArgs& a) {
    static const int pp_env = []() {
        const char* e = getenv("LSX_SW_PP");
        return e ? atoi(e) : 1;   // measured on B200 (profiles/r02e): two interleaved primes are SLOWER (3.37 vs 2.88 ms per 2^18 16x17 systems: 128 registers, fewer resident warps, twice the code per pass)
    }();
    if constexpr (NC <= 24) {
        if (a.K >= 2 && pp_env >= 2) return launch_sw_pp<G, NC, 2>(ctx, a);
    }
    return launch_sw_pp<G, NC, 1>(ctx, a);
}

// Instantiated shapes: G lanes per matrix, NC register columns (n is padded up to NC with zero columns).
struct SwShape { int G, NC; };
constexpr SwShape kShapes[] = {{4, 4}, {4, 5}, {4, 8}, {4, 9}, {8, 9}, {8, 12}, {8, 17}, {16, 17}, {16, 24}, {16, 33}, {32, 33}};

bool pick_shape(int m, int n, SwShape* out) {
    for (const SwShape& s : kShapes)
        if (m <= s.G && n <= s.NC) {
            *out = s;
            return true;
        }
    return false;
}

int launch_shape(lsx_ctx* ctx, const SwArgs& a, SwShape s) {
    switch (s.G * 100 + s.NC) {
        case 404: return launch_sw<4, 4>(ctx, a);
        case 405: return launch_sw<4, 5>(ctx, a);
        case 408: return launch_sw<4, 8>(ctx, a);
        case 409: return launch_sw<4, 9>(ctx, a);
        case 809: return launch_sw<8, 9>(ctx, a);
        case 812: return launch_sw<8, 12>(ctx, a);
        case 817: return launch_sw<8, 17>(ctx, a);
        case 1617: return launch_sw<16, 17>(ctx, a);
        case 1624: return launch_sw<16, 24>(ctx, a);
        case 1633: return launch_sw<16, 33>(ctx, a);
        default: return launch_sw<32, 33>(ctx, a);
    }
}

}  // namespace

// *handled = 1 when the shape/plan is covered; bad-prime matrices are then recomputed by the tile path.
int lsx_run_subwarp(lsx_ctx* ctx, const ElimJob& job, int* handled) {
    *handled = 0;
    if (getenv("LSX_DISABLE_SUBWARP")) return LSX_OK;
    if (job.m > 32 || job.n > 33 || job.bar > 32 || job.K > 8 || job.L > 8) return LSX_OK;
    if (job.a_abs_max >= (1 << 30) || job.b_abs_max >= (1 << 30)) return LSX_OK;
    SwShape shape;
    if (!pick_shape(job.m, job.n, &shape)) return LSX_OK;
    const int NC = shape.NC;
    if (sw_smem_bytes(job.K, NC) > 160 * 1024) return LSX_OK;

    // scratch: retry list + counter, then room for the tile path's retry pass
    const size_t o_list = 0, o_count = (size_t)LSX_RETRY_CAP * 4, o_child = o_count + 256;
    int rc = lsx_ws_reserve(ctx, o_child + lsx_generic_ws_bytes(job, 1) + 4096);
    if (rc != LSX_OK) return rc;
    char* base = (char*)ctx->d_ws;
    int32_t* rlist = (int32_t*)(base + o_list);
    int32_t* rcount = (int32_t*)(base + o_count);
    LSX_CUDA_TRY(ctx, cudaMemsetAsync(rcount, 0, 4, ctx->stream));
    if (job.op == LSX_OP_SOLVE) {
        const int nvars = job.n - 1;
        LSX_CUDA_TRY(ctx, cudaMemsetAsync(job.particular, 0, (size_t)job.batch * nvars * job.L * 4, ctx->stream));
        if (job.generators && job.gen_cap > 0)
            LSX_CUDA_TRY(ctx, cudaMemsetAsync(job.generators, 0, (size_t)job.batch * nvars * job.gen_cap * job.L * 4,
                                              ctx->stream));
    }
    SwArgs a{};
    a.A = job.A;
    a.bvec = job.bvec;
    a.batch = job.batch;
    a.m = job.m;
    a.n_in = job.n_in;
    a.n = job.n;
    a.bar = job.bar;
    a.right_identity = job.right_identity;
    a.op = job.op;
    a.K = job.K;
    a.L = job.L;
    a.gen_cap = job.gen_cap;
    a.pivot_slots = job.m < job.bar ? job.m : job.bar;
    a.max_rank = job.max_rank;
    a.a_abs_max = (int)job.a_abs_max;
    a.b_abs_max = (int)job.b_abs_max;
    a.primes = ctx->d_primes;
    a.garner = ctx->d_garner;
    a.num = job.num;
    a.den = job.den;
    a.particular = job.particular;
    a.generators = job.generators;
    a.pivot_col = job.pivot_col;
    a.rank = job.rank;
    a.status = job.status;
    a.retry_list = rlist;
    a.retry_count = rcount;
    a.retry_cap = LSX_RETRY_CAP;
    if (job.K > 1 && job.op != LSX_OP_RANK && !getenv("LSX_NO_DATA_BOUND")) {
        // prime count from the row norms of the batch (the plan's K covers the declared magnitudes)
        int32_t* kw = (int32_t*)(base + o_count + 64);
        rc = lsx_row_bound_primes(ctx, job, kw);
        if (rc != LSX_OK) return rc;
        a.kword = kw;
    } else {
        ctx->last_kword = nullptr;
    }
    rc = launch_shape(ctx, a, shape);
    if (rc != LSX_OK) return rc;
    // bad-prime matrices (normally none): tile path in list mode, scratch behind the list.  A plan of ONE prime has no
    // such matrices: the prime exceeds twice the Hadamard bound, so no non-zero minor vanishes modulo it and there is
    // no second profile to disagree with -- the three launches of the empty retry pass were half of a config 1 step.
    if (job.K > 1) {
        rc = lsx_run_generic(ctx, job, rlist, rcount, LSX_RETRY_CAP, o_child);
        if (rc != LSX_OK) return rc;
    }
    *handled = 1;
    return LSX_OK;
}
