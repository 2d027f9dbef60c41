// Generic path of the exact elimination engine: one CTA per (matrix, prime) with the residue
// tile resident in shared memory, then profile verification and a Garner-CRT assemble kernel.
//
// This path covers every shape that fits shared memory (e.g. 64x128 [A|I]) and is also the
// correctness backstop of the fused small-matrix kernels: matrices whose primes disagree on the
// pivot profile (a "bad prime" divided an intermediate pivot candidate) are recomputed here with
// replacement primes.
//
// Kernels of this file: k_row_bound / k_bound_to_primes (prime count from the row norms of the pass's matrices),
// k_tile_elim (tile in shared memory, any shape), k_tile_reg (tile in registers, up to 128 x 128 cells), k_tile_inv
// (in-place inverse of a square [A|I] in an m x m register tile), k_verify (pivot profiles of the primes), k_assemble
// (Garner CRT into the layout of the operation).
//
// Algorithm per (matrix, prime) -- mirrors tests/device_model.py::elim_words:
//   uniform-scale division-free Gauss-Jordan on Montgomery words.  At a pivot step with pivot
//   value piv (row pi, column j) and current common scale S every row r != pi becomes
//   piv*row_r - W[r][j]*row_pi and the pivot row becomes S*row_pi, so ALL rows carry the scale
//   S' = S*piv.  No modular inverse is needed inside the loop; one Fermat inversion at the end
//   turns the tile into N = d * RREF with d = sign * prod(true pivots) (the determinant of the
//   pivot minor).  The pivot rule is the reference's (linalg.py:548-567): the entry at the pivot
//   position if non-zero, else the first lower row with a non-zero entry, else skip the column.
#include <stdarg.h>
#include <stdio.h>

#include <type_traits>
#include <utility>

#include "lsx_internal.h"
#include "lsx_crt.cuh"

namespace {

struct TileArgs {
    const int32_t* A;
    const int32_t* bvec;
    const int32_t* list;        // NULL: slot == matrix index
    const int32_t* list_count;  // device pointer (list mode)
    int64_t batch;              // number of slots when list == NULL
    int64_t cap;                // slot stride of the scratch arrays
    int m, n_in, n, bar, right_identity;
    int c0, c1;                 // columns [c0, c1) are stored to res
    int64_t a_abs_max, b_abs_max;
    const PrimeRec* primes;
    uint32_t* res;              // [Ktot][cap][m*(c1-c0)]
    uint32_t* dres;             // [Ktot][cap]
    int32_t* rankk;             // [Ktot][cap]
    uint8_t* prof;              // [Ktot][cap][bar]
    int32_t* status;            // [batch] (indexed by matrix)
    const int32_t* kword;       // NULL, or kword[1] = primes the data of this pass needs (k_row_bound)
    int rhs_out_of_tile;        // k_tile_reg: the declared-zero right-hand side column n - 1 is not kept in the tile
    int k_extra;                // replacement primes on top of that (list mode)
    int ktot;                   // primes of the launch: the grid is grid_x * ktot CTAs, prime index fastest
};

// ---- prime count from the data -------------------------------------------------------------------------------------
// The plan sizes K for the declared magnitudes (every entry at a_abs_max).  Every integer an elimination returns is a
// minor of [A | right part] of order <= r_top, so by Hadamard's inequality over ROWS its magnitude is at most the
// product of the r_top largest row norms of the batch's own matrices -- rows of a minor are sub-rows.  k_row_bound
// takes the maximum of that product over the matrices of a pass, k_bound_to_primes turns it into a prime count
// (never above the plan's K; same rounding as lsx_bits_to_plan), and the CTAs of the primes beyond it leave at once;
// k_verify and k_assemble read the same word.  Random entries in [-5, 5] have row norms of sqrt(640) where the plan
// assumes 40: 10 primes instead of 12 for the 64 x 64 inverse, 15 instead of 21 for the rank-48 kernel bases.
struct BoundArgs {
    const int32_t* A;
    const int32_t* bvec;        // right-hand side of a solve (kept out of the row norms, see below) or NULL
    int64_t batch;
    int m, n_in, r, r_top, right_identity, exact_int;
    int32_t* kword;             // [0] = max over matrices of ceil(256 log2 bound), [1] = prime count
};
// With a right-hand side b the minors are of two kinds.  Those without the column b have order <= r and rows that are
// sub-rows of A: at most T_r = the product of the r largest row norms of A.  Those with it (the particular solution,
// and the entries of the zero-left rows that decide consistency, order r_top) expand along that column into
// sum_i +- b_i M_i with minors M_i of A of order r_top - 1: at most |b|_1 T_(r_top - 1).  Folding b into the row norms
// instead would cost every row the magnitude of b (b = A x makes |b_i| as large as a whole row norm of A).
__device__ __forceinline__ int bound_word(double t_r, double t_s1, double b1, bool has_b) {
    double tot = t_r;
    if (has_b) {
        const double lb = b1 > 1.0 ? log2(b1) * (1.0 + 1e-12) : 0.0;
        tot = fmax(t_r, lb + t_s1);
    }
    return (int)ceil(tot * 256.0) + 1;
}

// One thread per ROW, the rows of a matrix next to each other in the CTA (256 / m matrices per CTA), their logarithms
// exchanged through shared memory.  Declared magnitudes up to 2^27 (every call of the benchmarks): the squares are summed
// exactly in 64-bit integers and the logarithm is a float one taken of the sum rounded UP, plus 1e-4 (log2f is good to
// a few ulp: < 1e-5 on values below 64); beyond that everything in double.  An entry above the declared magnitude can
// wrap the integer sum -- that matrix is flagged LSX_ST_BOUND by the elimination kernel and returns nothing, and a
// smaller contribution of ITS rows does not affect the bound of the others.
__global__ void __launch_bounds__(256) k_row_bound(const BoundArgs a) {
    __shared__ double lg[256];
    __shared__ double part[3][256];
    const int m = a.m, n = a.n_in;
    const int per = 256 / m;                                    // matrices per CTA
    const int q = threadIdx.x / m, r = threadIdx.x - q * m;
    const int64_t mat = (int64_t)blockIdx.x * per + q;
    const bool live = q < per && mat < a.batch;
    double v = 0.0, babs = 0.0;
    if (live) {
        const int32_t* row = a.A + (mat * m + r) * (int64_t)n;
        const bool vec = (n & 3) == 0 && (reinterpret_cast<uintptr_t>(a.A) & 15) == 0;
        if (a.exact_int) {
            unsigned long long s = a.right_identity ? 1ull : 0ull;
            if (vec) {
                const int4* row4 = reinterpret_cast<const int4*>(row);
                for (int c = 0; c < (n >> 2); ++c) {
                    const int4 x = row4[c];
                    s += (unsigned long long)((long long)x.x * x.x) + (unsigned long long)((long long)x.y * x.y) +
                         (unsigned long long)((long long)x.z * x.z) + (unsigned long long)((long long)x.w * x.w);
                }
            } else {
                for (int c = 0; c < n; ++c) s += (unsigned long long)((long long)row[c] * row[c]);
            }
            v = s > 1ull ? (double)(0.5f * log2f(__ull2float_ru(s)) + 1e-4f) : 0.0;
        } else {
            double s = a.right_identity ? 1.0 : 0.0;
            for (int c = 0; c < n; ++c) {
                const double x = (double)row[c];
                s += x * x;
            }
            v = s > 1.0 ? 0.5 * log2(s) * (1.0 + 1e-12) : 0.0;
        }
        if (a.bvec) babs = fabs((double)a.bvec[mat * m + r]);
    }
    lg[threadIdx.x] = v;
    __syncthreads();
    int above = 0;                                              // position of my row in the descending order
    if (live && (a.r < m || a.bvec))
        for (int t = 0; t < m; ++t) {
            const double o = lg[q * m + t];
            above += (o > v || (o == v && t < r)) ? 1 : 0;
        }
    part[0][threadIdx.x] = (live && above < (a.bvec ? a.r : a.r_top)) ? v : 0.0;
    part[1][threadIdx.x] = (live && above < a.r_top - 1) ? v : 0.0;
    part[2][threadIdx.x] = babs;
    __syncthreads();
    if (live && r == 0) {
        double t_r = 0.0, t_s1 = 0.0, b1 = 0.0;
        for (int t = 0; t < m; ++t) {
            t_r += part[0][q * m + t];
            t_s1 += part[1][q * m + t];
            b1 += part[2][q * m + t];
        }
        atomicMax(a.kword, bound_word(t_r, t_s1, b1, a.bvec != nullptr));
    }
}

__global__ void k_bound_to_primes(int32_t* kword, int K) {
    const double need = (double)kword[0] / 256.0 + 1.0 + 1e-6;      // sign bit + slack, as lsx_bits_to_plan
    int k = (int)ceil(need / 30.999);
    kword[1] = k < 1 ? 1 : (k > K ? K : k);
}

template <int T>
__global__ void __launch_bounds__(T) k_tile_elim(const TileArgs a) {
    extern __shared__ uint32_t smem[];
    const int m = a.m, n = a.n, bar = a.bar;
    uint32_t* W = smem;
    uint32_t* prow = W + m * n;
    uint32_t* ycol = prow + n;
    int* sh = (int*)(ycol + m);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = T / 32;
    // one-dimensional grid, the prime index fastest: the CTAs of one matrix are neighbours in launch order, so its
    // input is read from DRAM once and served from L2 to the other primes
    const int kslot = (int)(blockIdx.x % (unsigned)a.ktot), slot0 = (int)(blockIdx.x / (unsigned)a.ktot);
    if (a.kword && kslot >= a.kword[1] + a.k_extra) return;   // more primes than the data of this pass needs
    const PrimeRec P = a.primes[kslot];
    const uint32_t p = P.p, pinv = P.pinv;
    const int64_t nslots = a.list ? min((int64_t)*a.list_count, a.cap) : a.batch;
    const int ncs = a.c1 - a.c0;

    for (int64_t slot = slot0; slot < nslots; slot += gridDim.x / (unsigned)a.ktot) {
        const int64_t mat = a.list ? (int64_t)a.list[slot] : slot;
        // ---- load the tile as raw words (value = a / R, see device_model.elim_words) ----
        bool bad = false;
        for (int r = warp; r < m; r += NW) {
            for (int c = lane; c < n; c += 32) {
                int32_t v;
                int64_t lim;
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
                int64_t av = v < 0 ? -(int64_t)v : (int64_t)v;
                bad |= av > lim;
                W[r * n + c] = word_of_int_any(v, p);
            }
        }
        if (bad) atomicOr(&a.status[mat], LSX_ST_BOUND);
        __syncthreads();

        uint32_t S = P.one, Q = P.one, X = 1u;
        int pi = 0;
        bool neg = false;
        uint8_t* prof = a.prof + ((int64_t)kslot * a.cap + slot) * bar;
        for (int j = 0; j < bar; ++j) {
            if (pi >= m) {
                if (tid == 0) prof[j] = LSX_PROF_SKIP;
                continue;
            }
            // pivot search: first row >= pi with a non-zero entry in column j
            if (warp == 0) {
                int src = -1;
                for (int base = pi; base < m; base += 32) {
                    int r = base + lane;
                    bool nz = r < m && W[r * n + j] != 0u;
                    unsigned bal = __ballot_sync(0xffffffffu, nz);
                    if (bal) {
                        src = base + __ffs(bal) - 1;
                        break;
                    }
                }
                if (lane == 0) sh[0] = src;
            }
            __syncthreads();
            const int src = sh[0];
            if (src < 0) {
                if (tid == 0) prof[j] = LSX_PROF_SKIP;
                __syncthreads();   // sh[0] is rewritten in the next iteration
                continue;
            }
            if (tid == 0) prof[j] = (uint8_t)src;
            if (src != pi) {
                for (int c = tid; c < n; c += T) {
                    uint32_t t0 = W[pi * n + c];
                    W[pi * n + c] = W[src * n + c];
                    W[src * n + c] = t0;
                }
                neg = !neg;
                __syncthreads();
            }
            const uint32_t piv = W[pi * n + j];
            for (int c = tid; c < n; c += T) prow[c] = W[pi * n + c];
            for (int r = tid; r < m; r += T) ycol[r] = (r == pi) ? 0u : p - W[r * n + j];
            __syncthreads();
            for (int r = warp; r < m; r += NW) {
                const uint32_t x = (r == pi) ? S : piv;
                const uint32_t y = ycol[r];
                for (int c = lane; c < n; c += 32)
                    W[r * n + c] = mont_fma2(x, W[r * n + c], y, prow[c], p, pinv);
            }
            __syncthreads();
            Q = mont_mul(Q, S, p, pinv);
            S = mont_mul(S, piv, p, pinv);
            X = mont_mul(X, P.r2, p, pinv);
            ++pi;
        }
        // ---- one inversion, then scale the tile to N = d * RREF (plain residues) ----
        const uint32_t qinv = mont_pow(Q, p - 2u, P.one, p, pinv);
        uint32_t Gw = mont_mul(qinv, X, p, pinv);
        if (neg && Gw) Gw = p - Gw;
        const uint32_t G2w = mont_mul(Gw, P.r2, p, pinv);
        if (ncs > 0) {
            uint32_t* out = a.res + ((int64_t)kslot * a.cap + slot) * ((int64_t)m * ncs);
            for (int r = warp; r < m; r += NW) {
                const uint32_t g = r < pi ? Gw : G2w;
                for (int c = a.c0 + lane; c < a.c1; c += 32)
                    out[r * ncs + (c - a.c0)] = mont_mul(g, W[r * n + c], p, pinv);
            }
        }
        if (tid == 0) {
            a.dres[(int64_t)kslot * a.cap + slot] = mont_mul(Gw, S, p, pinv);
            a.rankk[(int64_t)kslot * a.cap + slot] = pi;
        }
        __syncthreads();   // tile is reloaded by the next slot
    }
}

// ---- register-tiled variant: the residue tile lives in REGISTERS, shared memory only carries the
// pivot column, the pivot row and the row order ------------------------------------------------------
// 256 threads as a 16 x 16 grid; thread (ty, tx) owns the cells (ty + 16 a, tx + 16 b), a < RA, b < CB
// (m <= 16 RA, n <= 16 CB).  Rows never move physically: the row order is a permutation in shared
// memory (perm[logical position] = physical row), so the reference's "swap rows" (linalg.py:548-567)
// costs two bytes.  Per pivot step every cell needs one two-product update from registers, CB pivot-row
// words and RA pivot-column words from shared memory -- the smem traffic of k_tile_elim (3 accesses
// per cell and step) is gone.  While no column has been skipped, 16-column blocks that lie completely
// left of the pivot column are finished pivot columns and are not updated any more.
// TYN = thread rows (16: 256 threads; 8: 128 threads, every thread owns twice the rows -- the per-step bookkeeping
// (column publish, multiplier set-up, block tests) is per thread, so it is spread over twice the cell updates).
// any int32 into [0, p) for p > 2^30 (|v| <= 2^31 < 2 p): conditional corrections, no division
__device__ __forceinline__ uint32_t residue_fast(int32_t v, uint32_t p) {
    int64_t t = v;
    if (t < 0) t += p;
    if (t < 0) t += p;
    if (t >= (int64_t)p) t -= p;
    return (uint32_t)t;
}

template <class F, int... Is>
__device__ __forceinline__ void for_each_block(F& f, std::integer_sequence<int, Is...>) {
    (f(std::integral_constant<int, Is>{}), ...);
}

template <int RA, int CB, int TYN>
__global__ void __launch_bounds__(16 * TYN, TYN == 16 ? ((RA * CB <= 32) ? 3 : 2) : ((RA * CB <= 32) ? 5 : (RA * CB <= 64) ? 4 : 2)) k_tile_reg(const TileArgs a) {
    __shared__ uint32_t prow2[2][16 * CB];   // double buffered by the parity of the column: no barrier at the end of a step
    __shared__ uint32_t colbuf[TYN * RA];
    __shared__ uint8_t perm[TYN * RA];
    __shared__ uint8_t inv[TYN * RA];
    const int m = a.m, n = a.n, bar = a.bar;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31;
    // one-dimensional grid, the prime index fastest: the CTAs of one matrix are neighbours in launch order, so its
    // input is read from DRAM once and served from L2 to the other primes
    const int kslot = (int)(blockIdx.x % (unsigned)a.ktot), slot0 = (int)(blockIdx.x / (unsigned)a.ktot);
    if (a.kword && kslot >= a.kword[1] + a.k_extra) return;   // more primes than the data of this pass needs
    const PrimeRec P = a.primes[kslot];
    const uint32_t p = P.p, pinv = P.pinv;
    const int64_t nslots = a.list ? min((int64_t)*a.list_count, a.cap) : a.batch;
    const int ncs = a.c1 - a.c0;

    for (int64_t slot = slot0; slot < nslots; slot += gridDim.x / (unsigned)a.ktot) {
        const int64_t mat = a.list ? (int64_t)a.list[slot] : slot;
        uint32_t W[RA][CB];
        bool bad = false;
        {
            // declared magnitudes as 32-bit limits (an int32 entry cannot exceed a larger one anyway), one row pointer
            // per register row, and the division only for the tiny primes of the tests
            const int lim_a = (int)(a.a_abs_max < 0x7fffffff ? a.a_abs_max : 0x7fffffff);
            const int lim_b = (int)(a.b_abs_max < 0x7fffffff ? a.b_abs_max : 0x7fffffff);
            const bool big_p = p > (1u << 30);
#pragma unroll
            for (int ia = 0; ia < RA; ++ia) {
                const int r = ty + TYN * ia;
                const bool rowlive = r < m;
                const int32_t* arow = a.A + (mat * m + (rowlive ? r : 0)) * (int64_t)a.n_in;
                const int32_t bv = (rowlive && !a.right_identity && a.bvec && n > a.n_in) ? a.bvec[mat * m + r] : 0;
#pragma unroll
                for (int ib = 0; ib < CB; ++ib) {
                    const int c = tx + 16 * ib;
                    int32_t v = 0;
                    int lim = 0x7fffffff;
                    if (rowlive && c < n) {
                        if (c < a.n_in) {
                            v = arow[c];
                            lim = c < bar ? lim_a : lim_b;
                        } else if (a.right_identity) {
                            v = (c - a.n_in == r) ? 1 : 0;
                        } else {
                            v = bv;
                            lim = lim_b;
                        }
                    }
                    bad |= v > lim || v < -lim;
                    W[ia][ib] = big_p ? residue_fast(v, p) : word_of_int_any(v, p);
                }
            }
        }
        if (a.rhs_out_of_tile)                 // the right-hand side is declared zero and stays out of the tile: check it
            for (int q = tid; q < m; q += 16 * TYN) bad |= a.bvec[mat * m + q] != 0;
        if (bad) atomicOr(&a.status[mat], LSX_ST_BOUND);
        for (int q = tid; q < TYN * RA; q += 16 * TYN) perm[q] = (uint8_t)q;
        __syncthreads();

        uint32_t S = P.one, Q = P.one, X = 1u;
        int pi = 0;
        bool neg = false, tri = true;          // tri: no column skipped so far (pi == j)
        // [A|I]: the identity column of a physical row that has not been a pivot row yet is still a unit column
        // (its one entry carries the common scale S, the rest is zero), so whole 16-column blocks of the right
        // part need no update until one of their rows is used.  rb_on = number of right blocks maintained so far;
        // a block is switched on by writing the current scale into its 16 diagonal cells (the cells are raw words
        // that lose a factor R per step while S is a Montgomery word: the cell value is S / R).
        const bool lazy_id = a.right_identity && (a.n_in & 15) == 0 && n == a.n_in + m;
        const int nleft = a.n_in >> 4;         // 16-column blocks of the left part
        // homogeneous system (kernel(), linalg.py:749-756): the right-hand side is declared zero (entries above the
        // declared magnitude are flagged at the load) and row operations keep it zero; when it has its last 16-column
        // block to itself that block is never updated (64 x 65: 4 live blocks of 5)
        const int cb_live = (!a.rhs_out_of_tile && !a.right_identity && a.bvec && a.b_abs_max == 0 && n == a.n_in + 1 &&
                             (a.n_in & 15) == 0 && n > 16 * (CB - 1)) ? CB - 1 : CB;
        int rb_on = 0;
        uint8_t* prof = a.prof + ((int64_t)kslot * a.cap + slot) * bar;
        // The pivot columns are walked block by block with the 16-column block index JB a COMPILE-TIME constant (the
        // loop body is instantiated once per block): the column publish then reads W[ia][JB] with a static register
        // index.  With a run-time block index it took a select chain over all CB blocks per row and step (a switch is
        // turned into a dynamically indexed access by the compiler, which put the whole tile into LOCAL memory: ncu,
        // round 2, one LDL + one STL per cell update).
        auto block_steps = [&](auto jb_c) {
          constexpr int JB = decltype(jb_c)::value;
          for (int jj = 0; jj < 16; ++jj) {
            const int j = 16 * JB + jj;
            if (j >= bar) break;
            if (pi >= m) {
                if (tid == 0) prof[j] = LSX_PROF_SKIP;
                continue;
            }
            // 1. publish column j (indexed by physical row)
            if (tx == jj) {
#pragma unroll
                for (int ia = 0; ia < RA; ++ia) colbuf[ty + TYN * ia] = W[ia][JB];
            }
            __syncthreads();
            // 2. pivot search over logical positions pi .. m-1: the first non-zero.  Every warp does it
            //    for itself (a few instructions), which saves a barrier; thread 0 records the result.
            int src = -1;
            for (int base = pi; base < m; base += 32) {
                const int q = base + lane;
                const bool nz = q < m && colbuf[perm[q]] != 0u;
                const unsigned bal = __ballot_sync(0xffffffffu, nz);
                if (bal) {
                    src = base + __ffs(bal) - 1;
                    break;
                }
            }
            if (tid == 0) prof[j] = src >= 0 ? (uint8_t)src : (uint8_t)LSX_PROF_SKIP;
            if (src < 0) {
                tri = false;
                __syncthreads();               // colbuf is rewritten by the next column
                continue;
            }
            neg ^= src != pi;
            const int prp = perm[src];         // physical row of the pivot (perm is swapped after the barrier)
            if (lazy_id) {
                // the pivot row's own identity column (n_in + prp) becomes active now, with every block before it
                while (rb_on <= (prp >> 4)) {
                    // the diagonal cell (r, n_in + r) of row r = ty + 16 * rb_on.  Written as selects over the whole
                    // tile: a conditional store at a run-time (ia, ib) is a dynamically indexed store to the compiler,
                    // and that alone moved the tile from registers to local memory (128 bytes of stack per thread).
                    const uint32_t sv = mont_redc((uint64_t)S, p, pinv);
#pragma unroll
                    for (int ia = 0; ia < RA; ++ia)
#pragma unroll
                        for (int ib = 0; ib < CB; ++ib) {
                            const bool hit = ib == nleft + rb_on && ty + TYN * ia == tx + 16 * rb_on && ty + TYN * ia < m;
                            W[ia][ib] = hit ? sv : W[ia][ib];
                        }
                    ++rb_on;
                }
            }
            // 3. publish the pivot row
            uint32_t* prow = prow2[j & 1];
            if (ty == prp % TYN) {
                const int pb = prp / TYN;      // same: selects, not a switch
#pragma unroll
                for (int ib = 0; ib < CB; ++ib) {
                    uint32_t v = W[0][ib];
#pragma unroll
                    for (int ia = 1; ia < RA; ++ia) v = (pb == ia) ? W[ia][ib] : v;
                    prow[tx + 16 * ib] = v;
                }
            }
            const uint32_t piv = colbuf[prp];
            uint32_t xs[RA], ys[RA];
#pragma unroll
            for (int ia = 0; ia < RA; ++ia) {
                const int r = ty + TYN * ia;
                const uint32_t f = colbuf[r];
                const bool isp = r == prp;
                xs[ia] = isp ? S : piv;
                ys[ia] = (isp || f == 0u) ? 0u : p - f;
            }
            __syncthreads();
            // 4. update (register resident)
            const int b0 = tri ? ((j + 1) >> 4) : 0;        // blocks left of the pivot column are finished
            const int b1 = lazy_id ? nleft + rb_on : cb_live;   // right blocks that are still pure unit columns
#pragma unroll
            for (int ib = 0; ib < CB; ++ib) {
                if (ib >= b0 && ib < b1) {
                    const uint32_t pc = prow[tx + 16 * ib];
#pragma unroll
                    for (int ia = 0; ia < RA; ++ia) W[ia][ib] = mont_fma2(xs[ia], W[ia][ib], ys[ia], pc, p, pinv);
                }
            }
            Q = mont_mul(Q, S, p, pinv);
            S = mont_mul(S, piv, p, pinv);
            X = mont_mul(X, P.r2, p, pinv);
            if (tid == 0 && src != pi) {       // every warp has read perm[src] before the barrier above
                const uint8_t t0 = perm[pi];
                perm[pi] = perm[src];
                perm[src] = t0;
            }
            ++pi;
            // No barrier here.  colbuf and perm were last read before barrier 2 of this step; the pivot row is still being
            // read by slower warps, but the next column writes the OTHER pivot-row buffer, and this one is not written again
            // before every warp has passed a barrier of the column in between.
          }
        };
        for_each_block(block_steps, std::make_integer_sequence<int, CB>{});
        if (lazy_id) {
            // rank-deficient input: blocks never switched on still hold the initial ones; give them the final scale
            while (rb_on < ((m + 15) >> 4)) {
                const uint32_t sv = mont_redc((uint64_t)S, p, pinv);
#pragma unroll
                for (int ia = 0; ia < RA; ++ia)
#pragma unroll
                    for (int ib = 0; ib < CB; ++ib) {
                        const bool hit = ib == nleft + rb_on && ty + TYN * ia == tx + 16 * rb_on && ty + TYN * ia < m;
                        W[ia][ib] = hit ? sv : W[ia][ib];
                    }
                ++rb_on;
            }
        }
        // ---- one inversion, scale to N = d * RREF (plain residues), rows in logical order ----
        __syncthreads();                       // thread 0 may have swapped perm in the last step
        for (int q = tid; q < m; q += 16 * TYN) inv[perm[q]] = (uint8_t)q;
        __syncthreads();
        const uint32_t qinv = mont_pow(Q, p - 2u, P.one, p, pinv);
        uint32_t Gw = mont_mul(qinv, X, p, pinv);
        if (neg && Gw) Gw = p - Gw;
        const uint32_t G2w = mont_mul(Gw, P.r2, p, pinv);
        if (ncs > 0) {
            uint32_t* out = a.res + ((int64_t)kslot * a.cap + slot) * ((int64_t)m * ncs);
#pragma unroll
            for (int ia = 0; ia < RA; ++ia) {
                const int r = ty + TYN * ia;
                if (r < m) {
                    const int q = inv[r];
                    const uint32_t g = q < pi ? Gw : G2w;
#pragma unroll
                    for (int ib = 0; ib < CB; ++ib) {
                        const int c = tx + 16 * ib;
                        if (c >= a.c0 && c < a.c1) out[q * ncs + (c - a.c0)] = mont_mul(g, W[ia][ib], p, pinv);
                    }
                }
            }
            if (a.rhs_out_of_tile && n - 1 >= a.c0 && n - 1 < a.c1)
                for (int q = tid; q < m; q += 16 * TYN) out[q * ncs + (n - 1 - a.c0)] = 0u;
        }
        if (tid == 0) {
            a.dres[(int64_t)kslot * a.cap + slot] = mont_mul(Gw, S, p, pinv);
            a.rankk[(int64_t)kslot * a.cap + slot] = pi;
        }
        __syncthreads();
    }
}

// ---- in-place inverse: [A|I] with square A in an m x m register tile ------------------------------------------------
// Gauss-Jordan on [A|I] keeps exactly m "live" columns while no column has been skipped: the left columns right of the
// pivot column, and the identity columns of the rows that have been pivot rows already.  The left column that step j
// eliminates dies in the same step in which the identity column of its pivot row comes alive, so the new column takes
// its slot: the owner of slot j replaces its (published) column by the unit column of the pivot row under the common
// scale (S / R on the pivot row, 0 elsewhere) before the pivot row is published, and the ordinary two-product update of
// ALL slots then yields the same words as the update of the 2m-wide tile in k_tile_reg.  Half the registers of the
// 64 x 128 tile, every update a useful one (k_tile_reg<4,8> carries 5 live blocks of 8 on average), and shapes that are
// no multiple of 16 need no special case.  At the end slot j holds the right column of the row at logical position j
// (perm[j]), which is undone in the store.  After a skipped column (singular input, or a prime that divides a pivot
// candidate) the right part is abandoned -- k_assemble stores zeros for rank < m and k_verify drops a deviating
// prime -- and the elimination goes on over the left columns alone, so rank and pivot profile stay those of k_tile_reg.
// Only LSX_OP_INVERSE takes this kernel (columns [n_in, n_in + m) stored, bar == n_in == m).
template <int RA, int CB, int TYN, int MINB>
__global__ void __launch_bounds__(16 * TYN, MINB) k_tile_inv(const TileArgs a) {
    __shared__ uint32_t prow2[2][16 * CB];   // double buffered by the parity of the column (see k_tile_reg)
    __shared__ uint32_t colbuf[TYN * RA];
    __shared__ uint8_t perm[TYN * RA];
    __shared__ uint8_t inv[TYN * RA];
    const int m = a.m;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31;
    // one-dimensional grid, the prime index fastest: the CTAs of one matrix are neighbours in launch order, so its
    // input is read from DRAM once and served from L2 to the other primes
    const int kslot = (int)(blockIdx.x % (unsigned)a.ktot), slot0 = (int)(blockIdx.x / (unsigned)a.ktot);
    if (a.kword && kslot >= a.kword[1] + a.k_extra) return;   // more primes than the data of this pass needs
    const PrimeRec P = a.primes[kslot];
    const uint32_t p = P.p, pinv = P.pinv;
    const int64_t nslots = a.list ? min((int64_t)*a.list_count, a.cap) : a.batch;

    for (int64_t slot = slot0; slot < nslots; slot += gridDim.x / (unsigned)a.ktot) {
        const int64_t mat = a.list ? (int64_t)a.list[slot] : slot;
        uint32_t W[RA][CB];
        bool bad = false;
        {
            const int lim_a = (int)(a.a_abs_max < 0x7fffffff ? a.a_abs_max : 0x7fffffff);
            const bool big_p = p > (1u << 30);
#pragma unroll
            for (int ia = 0; ia < RA; ++ia) {
                const int r = ty + TYN * ia;
                const bool rowlive = r < m;
                const int32_t* arow = a.A + (mat * m + (rowlive ? r : 0)) * (int64_t)m;
#pragma unroll
                for (int ib = 0; ib < CB; ++ib) {
                    const int c = tx + 16 * ib;
                    const int32_t v = (rowlive && c < m) ? arow[c] : 0;
                    bad |= v > lim_a || v < -lim_a;
                    W[ia][ib] = big_p ? residue_fast(v, p) : word_of_int_any(v, p);
                }
            }
        }
        if (bad) atomicOr(&a.status[mat], LSX_ST_BOUND);
        for (int q = tid; q < TYN * RA; q += 16 * TYN) perm[q] = (uint8_t)q;
        __syncthreads();

        uint32_t S = P.one, Q = P.one, X = 1u;
        int pi = 0;
        bool neg = false, tri = true;          // tri: no column skipped so far (pi == j), the slots hold the live columns
        uint8_t* prof = a.prof + ((int64_t)kslot * a.cap + slot) * a.bar;
        auto block_steps = [&](auto jb_c) {
          constexpr int JB = decltype(jb_c)::value;
          for (int jj = 0; jj < 16; ++jj) {
            const int j = 16 * JB + jj;
            if (j >= m) break;
            if (pi >= m) {
                if (tid == 0) prof[j] = LSX_PROF_SKIP;
                continue;
            }
            // 1. publish column j (indexed by physical row)
            if (tx == jj) {
#pragma unroll
                for (int ia = 0; ia < RA; ++ia) colbuf[ty + TYN * ia] = W[ia][JB];
            }
            __syncthreads();
            // 2. pivot search over logical positions pi .. m-1: the first non-zero (every warp for itself)
            int src = -1;
            for (int base = pi; base < m; base += 32) {
                const int q = base + lane;
                const bool nz = q < m && colbuf[perm[q]] != 0u;
                const unsigned bal = __ballot_sync(0xffffffffu, nz);
                if (bal) {
                    src = base + __ffs(bal) - 1;
                    break;
                }
            }
            if (tid == 0) prof[j] = src >= 0 ? (uint8_t)src : (uint8_t)LSX_PROF_SKIP;
            if (src < 0) {
                tri = false;
                __syncthreads();               // colbuf is rewritten by the next column
                continue;
            }
            neg ^= src != pi;
            const int prp = perm[src];         // physical row of the pivot (perm is swapped after the barrier)
            // 3. slot j dies as a left column and comes alive as the identity column of the pivot row
            if (tri && tx == jj) {
                const uint32_t sv = mont_redc((uint64_t)S, p, pinv);
#pragma unroll
                for (int ia = 0; ia < RA; ++ia) W[ia][JB] = (ty + TYN * ia == prp) ? sv : 0u;
            }
            // 4. publish the pivot row (selects, not a switch: see k_tile_reg)
            uint32_t* prow = prow2[j & 1];
            if (ty == prp % TYN) {
                const int pb = prp / TYN;
#pragma unroll
                for (int ib = 0; ib < CB; ++ib) {
                    uint32_t v = W[0][ib];
#pragma unroll
                    for (int ia = 1; ia < RA; ++ia) v = (pb == ia) ? W[ia][ib] : v;
                    prow[tx + 16 * ib] = v;
                }
            }
            const uint32_t piv = colbuf[prp];
            uint32_t xs[RA], ys[RA];
#pragma unroll
            for (int ia = 0; ia < RA; ++ia) {
                const int r = ty + TYN * ia;
                const uint32_t f = colbuf[r];
                const bool isp = r == prp;
                xs[ia] = isp ? S : piv;
                ys[ia] = (isp || f == 0u) ? 0u : p - f;
            }
            __syncthreads();
            // 5. update every slot
#pragma unroll
            for (int ib = 0; ib < CB; ++ib) {
                const uint32_t pc = prow[tx + 16 * ib];
#pragma unroll
                for (int ia = 0; ia < RA; ++ia) W[ia][ib] = mont_fma2(xs[ia], W[ia][ib], ys[ia], pc, p, pinv);
            }
            Q = mont_mul(Q, S, p, pinv);
            S = mont_mul(S, piv, p, pinv);
            X = mont_mul(X, P.r2, p, pinv);
            if (tid == 0 && src != pi) {       // every warp has read perm[src] before the barrier above
                const uint8_t t0 = perm[pi];
                perm[pi] = perm[src];
                perm[src] = t0;
            }
            ++pi;
            // no barrier here: see k_tile_reg
          }
        };
        for_each_block(block_steps, std::make_integer_sequence<int, CB>{});
        // ---- one inversion, scale to N = d * RREF (plain residues), rows in logical order, columns by pivot row ----
        __syncthreads();                       // the last swap of perm
        for (int q = tid; q < m; q += 16 * TYN) inv[perm[q]] = (uint8_t)q;
        __syncthreads();
        const uint32_t qinv = mont_pow(Q, p - 2u, P.one, p, pinv);
        uint32_t Gw = mont_mul(qinv, X, p, pinv);
        if (neg && Gw) Gw = p - Gw;
        const uint32_t G2w = mont_mul(Gw, P.r2, p, pinv);
        {
            uint32_t* out = a.res + ((int64_t)kslot * a.cap + slot) * ((int64_t)m * m);
            int colof[CB];                     // slot c holds the identity column of the row at logical position c
#pragma unroll
            for (int ib = 0; ib < CB; ++ib) {
                const int c = tx + 16 * ib;
                colof[ib] = c < m ? perm[c] : -1;
            }
#pragma unroll
            for (int ia = 0; ia < RA; ++ia) {
                const int r = ty + TYN * ia;
                if (r < m) {
                    const int q = inv[r];
                    const uint32_t g = q < pi ? Gw : G2w;
#pragma unroll
                    for (int ib = 0; ib < CB; ++ib)
                        if (colof[ib] >= 0) out[q * m + colof[ib]] = mont_mul(g, W[ia][ib], p, pinv);
                }
            }
        }
        if (tid == 0) {
            a.dres[(int64_t)kslot * a.cap + slot] = mont_mul(Gw, S, p, pinv);
            a.rankk[(int64_t)kslot * a.cap + slot] = pi;
        }
        __syncthreads();
    }
}

template <int RA, int CB, int TYN, int MINB>
int launch_tile_inv(lsx_ctx* ctx, const TileArgs& ta, int Ktot, int64_t grid_x) {
    dim3 grid((unsigned)(grid_x * Ktot));      // prime index fastest (TileArgs.ktot == Ktot)
    lsx_timing_begin(ctx);
    k_tile_inv<RA, CB, TYN, MINB><<<grid, 16 * TYN, 0, ctx->stream>>>(ta);
    lsx_timing_end(ctx);
    ctx->launches++;
    return LSX_OK;
}

// In-place inverse kernel for a square [A|I] job whose right part alone is stored, or false.
bool launch_tile_inv_any(lsx_ctx* ctx, const TileArgs& ta, int Ktot, int64_t grid_x, int* rc) {
    if (getenv("LSX_DISABLE_TILE_REG") || getenv("LSX_DISABLE_TILE_INV")) return false;
    const int m = ta.m;
    if (!ta.right_identity || ta.n_in != m || ta.bar != m || ta.n != 2 * m || ta.c0 != m || ta.c1 != 2 * m) return false;
    if (m > 128 || m < 9) return false;
    const char* shp = getenv("LSX_TILE_INV_SHAPE");      // measurement switch for the 64-row tile
    const int s = shp ? atoi(shp) : 0;
    if (m <= 32) *rc = launch_tile_inv<2, 2, 16, 4>(ctx, ta, Ktot, grid_x);
    else if (m <= 64) {
        if (s == 1) *rc = launch_tile_inv<4, 4, 16, 3>(ctx, ta, Ktot, grid_x);
        else if (s == 2) *rc = launch_tile_inv<8, 4, 8, 4>(ctx, ta, Ktot, grid_x);
        else if (s == 3) *rc = launch_tile_inv<8, 4, 8, 6>(ctx, ta, Ktot, grid_x);
        else *rc = launch_tile_inv<8, 4, 8, 5>(ctx, ta, Ktot, grid_x);
    } else *rc = launch_tile_inv<8, 8, 16, 2>(ctx, ta, Ktot, grid_x);
    return true;
}

template <int RA, int CB, int TYN = 16>
int launch_tile_reg(lsx_ctx* ctx, const TileArgs& ta, int Ktot, int64_t grid_x) {
    dim3 grid((unsigned)(grid_x * Ktot));      // prime index fastest (TileArgs.ktot == Ktot)
    lsx_timing_begin(ctx);
    k_tile_reg<RA, CB, TYN><<<grid, 16 * TYN, 0, ctx->stream>>>(ta);
    lsx_timing_end(ctx);
    ctx->launches++;
    return LSX_OK;
}

// Picks a register-tiled instantiation for the shape, or returns false (fall back to k_tile_elim).
bool launch_tile_reg_any(lsx_ctx* ctx, const TileArgs& ta_in, int Ktot, int64_t grid_x, int* rc) {
    if (getenv("LSX_DISABLE_TILE_REG")) return false;
    TileArgs ta = ta_in;
    const int m = ta.m, n = ta.n;
    // homogeneous system whose 64 left columns fill four blocks exactly (kernel() of a 64-column matrix): the zero
    // right-hand side stays out of the tile, 32 cells per thread instead of 40 and 5 CTAs per SM instead of 4
    if (m <= 64 && n == 65 && ta.n_in == 64 && !ta.right_identity && ta.bvec && ta.b_abs_max == 0 && m * n >= 256 &&
        !getenv("LSX_TILE_RHS_IN")) {
        ta.rhs_out_of_tile = 1;
        *rc = launch_tile_reg<8, 4, 8>(ctx, ta, Ktot, grid_x);
        return true;
    }
    if (m > 128 || n > 128 || m * n < 256) return false;
    // 64-row tiles, measured on B200 with the block-static step loop (profiles/r02t_tile.txt, per 4096 matrices):
    //   64 x 65 (kernel basis, 21 primes): 128 threads x 8 rows <8,5,8> 13.3 ms, 256 threads <4,5> 15.5 ms (was 18.5)
    //   64 x 128 (inverse, 12 primes):     256 threads <4,8> 13.5 ms, 128 threads <8,8,8> 16.3 ms (was 16.0; the eight
    //                                      instantiated step bodies of the 64-cell shape no longer share the instruction cache
    //                                      between the CTAs of an SM)
    // LSX_TILE_TY8 = 0 / 1 forces the 256- / 128-thread shapes for both.
    const char* ty8 = getenv("LSX_TILE_TY8");
    const int t8 = ty8 ? (atoi(ty8) != 0 ? 1 : 0) : -1;
    if (m <= 32 && n <= 48) *rc = launch_tile_reg<2, 3>(ctx, ta, Ktot, grid_x);
    else if (m <= 64 && n <= 80) *rc = t8 != 0 ? launch_tile_reg<8, 5, 8>(ctx, ta, Ktot, grid_x) : launch_tile_reg<4, 5>(ctx, ta, Ktot, grid_x);
    else if (m <= 64 && n <= 128) *rc = t8 == 1 ? launch_tile_reg<8, 8, 8>(ctx, ta, Ktot, grid_x) : launch_tile_reg<4, 8>(ctx, ta, Ktot, grid_x);
    else *rc = launch_tile_reg<8, 8>(ctx, ta, Ktot, grid_x);
    return true;
}

// ---- verification: pick the primes that agree with the lexicographically smallest profile ----
struct VerifyArgs {
    const int32_t* list;
    const int32_t* list_count;
    int64_t batch, cap;
    int bar, K, Ktot, max_rank, pivot_slots, allow_retry;
    const int32_t* kword;
    const uint8_t* prof;
    const int32_t* rankk;
    uint8_t* sel;        // [cap][LSX_MAX_BATCH_PRIMES]
    int32_t* rank_ws;    // [cap]
    int32_t* piv_ws;     // [cap][pivot_slots]
    int32_t* status;     // by matrix
    int32_t* rank_out;   // by matrix or NULL
    int32_t* piv_out;    // by matrix or NULL
    int32_t* retry_list;
    int32_t* retry_count;
    int retry_cap;
};

__global__ void k_verify(const VerifyArgs a) {
    const int64_t nslots = a.list ? min((int64_t)*a.list_count, a.cap) : a.batch;
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= nslots) return;
    const int64_t mat = a.list ? (int64_t)a.list[slot] : slot;
    const int bar = a.bar;
    const int K = a.kword ? min(a.K, a.kword[1]) : a.K;       // primes the data needs (k_row_bound)
    const int Ktot = K + (a.Ktot - a.K);
    // A prime whose profile deviates from the rational one picks a LATER row or skips a column
    // at the first deviation (a non-zero candidate looked like zero), so the rational profile is
    // the lexicographic minimum as soon as one prime is good; if the primes agreeing with the
    // minimum have a product above the Hadamard bound the minimum is provably the rational
    // profile.
    // Profiles are compared 16 bytes at a time when the rows allow it (bar a multiple of 16: the rows of the scratch
    // array are then 16-byte aligned); byte by byte this kernel was a chain of 2 K bar dependent loads per thread
    // (53 us per 1351 64-column matrices x 10 primes, 5 % of the config 4 step).
    const bool vec = (bar & 15) == 0 && (reinterpret_cast<uintptr_t>(a.prof) & 15) == 0;
    auto same_rows = [&](const uint8_t* x, const uint8_t* y) {
        bool same = true;
        if (vec) {
            const uint4* x4 = reinterpret_cast<const uint4*>(x);
            const uint4* y4 = reinterpret_cast<const uint4*>(y);
            for (int j = 0; j < (bar >> 4); ++j) {
                const uint4 u = x4[j], v = y4[j];
                same &= u.x == v.x && u.y == v.y && u.z == v.z && u.w == v.w;
            }
        } else {
            for (int j = 0; j < bar; ++j) same &= x[j] == y[j];
        }
        return same;
    };
    int best = 0;
    for (int k = 1; k < Ktot; ++k) {
        const uint8_t* pk = a.prof + ((int64_t)k * a.cap + slot) * bar;
        const uint8_t* pb = a.prof + ((int64_t)best * a.cap + slot) * bar;
        if (same_rows(pk, pb)) continue;
        for (int j = 0; j < bar; ++j) {
            if (pk[j] != pb[j]) {
                if (pk[j] < pb[j]) best = k;
                break;
            }
        }
    }
    const uint8_t* pb = a.prof + ((int64_t)best * a.cap + slot) * bar;
    int cnt = 0;
    uint8_t* sel = a.sel + slot * LSX_MAX_BATCH_PRIMES;
    for (int k = 0; k < Ktot && cnt < K; ++k) {
        const uint8_t* pk = a.prof + ((int64_t)k * a.cap + slot) * bar;
        if (k == best || same_rows(pk, pb)) sel[cnt++] = (uint8_t)k;
    }
    int st = 0;
    if (cnt < K) {
        if (a.allow_retry) {
            int pos = atomicAdd(a.retry_count, 1);
            if (pos < a.retry_cap) {
                a.retry_list[pos] = (int32_t)mat;
                st = LSX_ST_INTERNAL_RETRY;
            } else {
                st = LSX_ST_NO_GOOD_PRIME;
            }
        } else {
            st = LSX_ST_NO_GOOD_PRIME;
        }
    } else if (!a.allow_retry) {
        st = LSX_ST_RETRIED;
    }
    const int rk = a.rankk[(int64_t)best * a.cap + slot];
    if (a.max_rank > 0 && rk > a.max_rank) st |= LSX_ST_BOUND;
    a.rank_ws[slot] = rk;
    int c = 0;
    for (int j = 0; j < bar; ++j) {
        if (pb[j] != LSX_PROF_SKIP) {
            a.piv_ws[slot * a.pivot_slots + c] = j;
            if (a.piv_out) a.piv_out[mat * a.pivot_slots + c] = j;
            ++c;
        }
    }
    for (; c < a.pivot_slots; ++c) {
        a.piv_ws[slot * a.pivot_slots + c] = -1;
        if (a.piv_out) a.piv_out[mat * a.pivot_slots + c] = -1;
    }
    if (a.rank_out) a.rank_out[mat] = rk;
    if (!a.allow_retry) atomicAnd(&a.status[mat], ~LSX_ST_INTERNAL_RETRY);
    if (st) atomicOr(&a.status[mat], st);
}

// ---- assemble: Garner CRT of every needed integer, written in the layout of the operation ----
struct AsmArgs {
    const int32_t* list;
    const int32_t* list_count;
    int64_t batch, cap;
    int op, m, n, n_in, bar, K, L, ncs, c0, pivot_slots, gen_cap;
    const PrimeRec* primes;
    const uint32_t* garner;   // [GD][GD]
    const int32_t* kword;
    const uint32_t* res;
    const uint32_t* dres;
    const uint8_t* sel;
    const int32_t* rank_ws;
    const int32_t* piv_ws;
    int32_t* status;
    uint32_t* num;
    uint32_t* den;
    uint32_t* particular;
    uint32_t* generators;
};

template <int KT>
__global__ void __launch_bounds__(128) k_assemble(const AsmArgs a) {
    const int64_t nslots = a.list ? min((int64_t)*a.list_count, a.cap) : a.batch;
    const int64_t E = (int64_t)a.m * a.ncs;       // stored entries per matrix
    // threads per matrix: one per stored entry + the denominator; a solve keeps at most gen_cap + 1 entries of a row
    // (free columns and the right-hand side), so its threads are dealt over (row, kept slot) instead
    const bool compact = a.op == LSX_OP_SOLVE;
    const int kept = a.gen_cap + 1;
    const int64_t per = (compact ? (int64_t)a.m * kept : E) + 1;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nslots * per) return;
    const int64_t slot = t / per;
    int64_t e = t - slot * per;
    const int64_t mat = a.list ? (int64_t)a.list[slot] : slot;
    const int st = a.status[mat];
    if (st & (LSX_ST_INTERNAL_RETRY | LSX_ST_NO_GOOD_PRIME | LSX_ST_BOUND)) return;
    const uint8_t* sel = a.sel + slot * LSX_MAX_BATCH_PRIMES;
    const int K = a.kword ? min(a.K, a.kword[1]) : a.K, L = a.L;
    const int rank = a.rank_ws[slot];

    // A solve keeps only part of the tile: the right-hand-side column and the FREE columns of the pivot rows (and a
    // zero test of the right-hand side below them).  Decide that before the residues are loaded and reconstructed:
    // for the rank-48 kernel bases of config 4 four fifths of the entries are pivot columns or zero rows.
    int solve_tfree = 0;
    if (compact) {
        if (e == per - 1) {
            e = E;                                // the denominator
        } else {
            const int nvars = a.n - 1;
            const int i = (int)(e / kept), q = (int)(e - (int64_t)i * kept);     // q < gen_cap: q-th free column, else rhs
            int j = nvars;
            if (i >= rank) {
                // zero left row: inconsistent iff the rhs is non-zero (linalg.py:913-934)
                if (q != a.gen_cap) return;
                const int64_t er = (int64_t)i * a.ncs + (nvars - a.c0);
                bool zero = true;
                for (int k = 0; k < K; ++k) zero &= a.res[((int64_t)sel[k] * a.cap + slot) * E + er] == 0u;
                if (!zero) atomicOr(&a.status[mat], LSX_ST_INCONSISTENT);
                return;
            }
            if (q != a.gen_cap) {
                // column of the q-th free variable (none: fewer than q + 1 free columns)
                const int32_t* piv = a.piv_ws + slot * a.pivot_slots;
                int k = 0, tfree = 0;
                j = -1;
                for (int c = 0; c < nvars; ++c) {
                    if (k < rank && piv[k] == c) {
                        ++k;
                    } else {
                        if (tfree == q) {
                            j = c;
                            break;
                        }
                        ++tfree;
                    }
                }
                if (j < 0) return;
                solve_tfree = q;
            }
            e = (int64_t)i * a.ncs + (j - a.c0);
        }
    }

    uint32_t r[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        r[k] = 0;
        if (k < K) {
            const int64_t base = (int64_t)sel[k] * a.cap + slot;
            r[k] = (e < E) ? a.res[base * E + e] : a.dres[base];
        }
    }
    uint32_t acc[KT];
    bool is_zero;
    crt_limbs<KT>(r, sel, K, a.primes, a.garner, acc, &is_zero);

    const int m = a.m;
    if (e == E) {   // denominator / determinant
        switch (a.op) {
            case LSX_OP_RREF:
            case LSX_OP_SOLVE:
                store_limbs<KT>(a.den + mat * L, acc, L, false, false);
                break;
            case LSX_OP_INVERSE:
                store_limbs<KT>(a.den + mat * L, acc, L, false, rank < m);
                if (rank < m) atomicOr(&a.status[mat], LSX_ST_SINGULAR);
                break;
            case LSX_OP_DET:
                store_limbs<KT>(a.den + mat * L, acc, L, false, rank < m);
                break;
            default:
                break;
        }
        if (a.op == LSX_OP_SOLVE) {
            // generator entries equal to d at the free columns (gen[f] = 1, linalg.py:976)
            const int nvars = a.n - 1;
            const int32_t* piv = a.piv_ws + slot * a.pivot_slots;
            int k = 0, tfree = 0;
            for (int c = 0; c < nvars; ++c) {
                if (k < rank && piv[k] == c) {
                    ++k;
                    continue;
                }
                if (tfree < a.gen_cap)
                    store_limbs<KT>(a.generators + ((mat * nvars + c) * a.gen_cap + tfree) * L, acc, L, false,
                                    false);
                ++tfree;
            }
            if (tfree > a.gen_cap) atomicOr(&a.status[mat], LSX_ST_GEN_TRUNC);
        }
        return;
    }
    const int i = (int)(e / a.ncs);
    const int j = a.c0 + (int)(e - (int64_t)i * a.ncs);
    switch (a.op) {
        case LSX_OP_RREF: {
            // pivot columns are not maintained by every elimination kernel: their values are known
            // (d on the pivot row, 0 elsewhere)
            const int32_t* piv = a.piv_ws + slot * a.pivot_slots;
            bool is_pivot_col = false, mine = false;
            for (int k = 0; k < rank; ++k)
                if (piv[k] == j) {
                    is_pivot_col = true;
                    mine = k == i;
                }
            if (is_pivot_col) {
                uint32_t rd[KT], accd[KT];
                bool zd;
#pragma unroll
                for (int k = 0; k < KT; ++k) rd[k] = k < K ? a.dres[(int64_t)sel[k] * a.cap + slot] : 0u;
                crt_limbs<KT>(rd, sel, K, a.primes, a.garner, accd, &zd);
                store_limbs<KT>(a.num + (mat * E + e) * L, accd, L, false, !mine);
            } else {
                store_limbs<KT>(a.num + (mat * E + e) * L, acc, L, false, false);
            }
            break;
        }
        case LSX_OP_INVERSE:
            store_limbs<KT>(a.num + (mat * E + e) * L, acc, L, false, rank < m);
            break;
        case LSX_OP_SOLVE: {
            // only the kept entries get here (see the early exit above): pivot rows, rhs or free column
            const int nvars = a.n - 1;
            const int32_t* piv = a.piv_ws + slot * a.pivot_slots;
            const int ci = piv[i];
            if (j == nvars)
                store_limbs<KT>(a.particular + (mat * nvars + ci) * L, acc, L, false, false);
            else
                store_limbs<KT>(a.generators + ((mat * nvars + ci) * a.gen_cap + solve_tfree) * L, acc, L, true, false);
            break;
        }
        default:
            break;
    }
}

template <int T>
int launch_tile(lsx_ctx* ctx, const TileArgs& ta, int Ktot, int64_t grid_x, size_t smem) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_tile_elim<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return lsx_fail(ctx, LSX_ERR_CUDA, "smem attribute: %s", cudaGetErrorString(e));
    }
    dim3 grid((unsigned)(grid_x * Ktot));      // prime index fastest (TileArgs.ktot == Ktot)
    lsx_timing_begin(ctx);
    k_tile_elim<T><<<grid, T, smem, ctx->stream>>>(ta);
    lsx_timing_end(ctx);
    ctx->launches++;
    return LSX_OK;
}

}  // namespace

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

namespace {
// Scratch layout of one generic pass (Ktot primes, `cap` slots).
struct GenericWs {
    size_t o_res, o_dres, o_rankk, o_prof, o_sel, o_rank, o_piv, o_rlist, o_rcount, o_kword, total;
};
GenericWs generic_ws(int Ktot, int64_t cap, int m, int ncs, int bar, int pivot_slots) {
    GenericWs w{};
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    w.o_res = take((size_t)Ktot * cap * m * ncs * 4);
    w.o_dres = take((size_t)Ktot * cap * 4);
    w.o_rankk = take((size_t)Ktot * cap * 4);
    w.o_prof = take((size_t)Ktot * cap * bar);
    w.o_sel = take((size_t)cap * LSX_MAX_BATCH_PRIMES);
    w.o_rank = take((size_t)cap * 4);
    w.o_piv = take((size_t)cap * pivot_slots * 4);
    w.o_rlist = take((size_t)LSX_RETRY_CAP * 4);
    w.o_rcount = take(256);
    w.o_kword = take(256);
    w.total = off;
    return w;
}
}  // namespace

// Bytes of device workspace lsx_run_generic needs for `batch` matrices of this job (parent pass
// plus the retry pass behind it).
size_t lsx_generic_ws_bytes(const ElimJob& job, int64_t batch) {
    int c0 = 0, c1 = job.n;
    if (job.op == LSX_OP_INVERSE) c0 = job.n_in;
    if (job.op == LSX_OP_DET || job.op == LSX_OP_RANK) c0 = c1 = 0;
    const int pivot_slots = job.m < job.bar ? job.m : job.bar;
    GenericWs a = generic_ws(job.K, batch, job.m, c1 - c0, job.bar, pivot_slots);
    GenericWs b = generic_ws(job.K + LSX_RETRY_EXTRA, LSX_RETRY_CAP, job.m, c1 - c0, job.bar, pivot_slots);
    return a.total + b.total;
}

// Runs the generic path for the matrices of `job` (all, or those in `list`).  When list == NULL
// the K plan primes are used and matrices whose primes disagree are appended to an internal
// retry list and recomputed with K + LSX_RETRY_EXTRA primes.  Scratch is taken from the ctx
// workspace starting at byte ws_offset (the caller has reserved lsx_generic_ws_bytes).
// kword[0..1] <- the prime count the matrices of `job` need (see k_row_bound); enqueued on the ctx's stream.
int lsx_row_bound_primes(lsx_ctx* ctx, const ElimJob& job, int32_t* kword) {
    LSX_CUDA_TRY(ctx, cudaMemsetAsync(kword, 0, 8, ctx->stream));
    const int m = job.m, n = job.n, bar = job.bar;
    BoundArgs ba{};
    ba.A = job.A;
    ba.bvec = (!job.right_identity && n > job.n_in) ? job.bvec : nullptr;
    ba.batch = job.batch;
    ba.m = m;
    ba.n_in = job.n_in;
    ba.right_identity = job.right_identity;
    int r = m < bar ? m : bar;
    if (job.max_rank > 0 && job.max_rank < r) r = job.max_rank;
    ba.r = r;
    ba.r_top = (n > bar && r < m) ? r + 1 : r;              // zero-left rows hold minors of order r + 1
    ba.kword = kword;
    ba.exact_int = (job.a_abs_max <= (1 << 27) && job.b_abs_max <= (1 << 27)) ? 1 : 0;
    const int per = 256 / m;                                // m <= 254 on every path that gets here
    k_row_bound<<<(unsigned)((job.batch + per - 1) / per), 256, 0, ctx->stream>>>(ba);
    k_bound_to_primes<<<1, 1, 0, ctx->stream>>>(kword, job.K);
    ctx->launches += 2;
    ctx->last_kword = kword;
    LSX_CUDA_TRY(ctx, cudaGetLastError());
    return LSX_OK;
}

static int run_generic(lsx_ctx* ctx, const ElimJob& job, const int32_t* list, const int32_t* list_count,
                       int list_cap, size_t ws_offset, const int32_t* kword_parent);

int lsx_run_generic(lsx_ctx* ctx, const ElimJob& job, const int32_t* list, const int32_t* list_count,
                    int list_cap, size_t ws_offset) {
    return run_generic(ctx, job, list, list_count, list_cap, ws_offset, nullptr);
}

static int run_generic(lsx_ctx* ctx, const ElimJob& job, const int32_t* list, const int32_t* list_count,
                       int list_cap, size_t ws_offset, const int32_t* kword_parent) {
    const bool list_mode = list != nullptr;
    const int K = job.K;
    const int Ktot = list_mode ? K + LSX_RETRY_EXTRA : K;
    if (Ktot > LSX_GARNER_DIM || K > LSX_MAX_BATCH_PRIMES)
        return lsx_fail(ctx, LSX_ERR_BOUND, "needs %d primes, limit is %d", K, LSX_MAX_BATCH_PRIMES);
    const int m = job.m, n = job.n, bar = job.bar;
    int c0 = 0, c1 = n;
    if (job.op == LSX_OP_INVERSE) c0 = job.n_in;
    if (job.op == LSX_OP_DET || job.op == LSX_OP_RANK) c0 = c1 = 0;
    const int ncs = c1 - c0;
    const int pivot_slots = m < bar ? m : bar;
    const int64_t cap = list_mode ? list_cap : job.batch;
    if (cap <= 0) return LSX_OK;

    const size_t smem = ((size_t)m * n + n + m + 4) * sizeof(uint32_t);
    if (smem > 200 * 1024)
        return lsx_fail(ctx, LSX_ERR_UNSUPPORTED, "matrix %dx%d does not fit the shared-memory tile", m, n);

    const GenericWs w = generic_ws(Ktot, cap, m, ncs, bar, pivot_slots);
    if (ws_offset + w.total > ctx->ws_bytes)
        return lsx_fail(ctx, LSX_ERR_CUDA, "internal: workspace too small (%zu + %zu > %zu)", ws_offset, w.total,
                        ctx->ws_bytes);
    char* base = (char*)ctx->d_ws + ws_offset;
    const size_t o_res = w.o_res, o_dres = w.o_dres, o_rankk = w.o_rankk, o_prof = w.o_prof, o_sel = w.o_sel,
                 o_rank = w.o_rank, o_piv = w.o_piv, o_rlist = w.o_rlist, o_rcount = w.o_rcount;
    uint32_t* res = (uint32_t*)(base + o_res);
    uint32_t* dres = (uint32_t*)(base + o_dres);
    int32_t* rankk = (int32_t*)(base + o_rankk);
    uint8_t* prof = (uint8_t*)(base + o_prof);
    uint8_t* sel = (uint8_t*)(base + o_sel);
    int32_t* rank_ws = (int32_t*)(base + o_rank);
    int32_t* piv_ws = (int32_t*)(base + o_piv);
    int32_t* rlist = (int32_t*)(base + o_rlist);
    int32_t* rcount = (int32_t*)(base + o_rcount);

    if (!list_mode) LSX_CUDA_TRY(ctx, cudaMemsetAsync(rcount, 0, 4, ctx->stream));

    // prime count from the row norms of this pass's matrices (a retry pass behind it keeps the parent's count; a
    // retry list handed in by the fused kernels has none and runs with the plan's K)
    const int32_t* kword = kword_parent;
    if (!list_mode && K > 1 && job.op != LSX_OP_RANK && !getenv("LSX_NO_DATA_BOUND")) {
        int32_t* kw = (int32_t*)(base + w.o_kword);
        const int brc = lsx_row_bound_primes(ctx, job, kw);
        if (brc != LSX_OK) return brc;
        kword = kw;
    } else if (!list_mode) {
        ctx->last_kword = nullptr;
    }

    TileArgs ta{};
    ta.A = job.A;
    ta.bvec = job.bvec;
    ta.list = list;
    ta.list_count = list_count;
    ta.batch = job.batch;
    ta.cap = cap;
    ta.m = m;
    ta.n_in = job.n_in;
    ta.n = n;
    ta.bar = bar;
    ta.right_identity = job.right_identity;
    ta.c0 = c0;
    ta.c1 = c1;
    ta.a_abs_max = job.a_abs_max;
    ta.b_abs_max = job.b_abs_max;
    ta.primes = ctx->d_primes;
    ta.res = res;
    ta.dres = dres;
    ta.rankk = rankk;
    ta.prof = prof;
    ta.status = job.status;
    ta.kword = kword;
    ta.k_extra = Ktot - K;
    ta.ktot = Ktot;

    int64_t gx = list_mode ? 256 : job.batch;
    const int64_t max_gx = (int64_t)ctx->sm_count * 64;
    if (gx > max_gx) gx = max_gx;
    const int cells = m * n;
    int rc;
    if (job.op == LSX_OP_INVERSE && launch_tile_inv_any(ctx, ta, Ktot, gx, &rc)) {
    } else if (launch_tile_reg_any(ctx, ta, Ktot, gx, &rc)) {
    } else if (cells <= 128)
        rc = launch_tile<32>(ctx, ta, Ktot, gx, smem);
    else if (cells <= 1024)
        rc = launch_tile<64>(ctx, ta, Ktot, gx, smem);
    else if (cells <= 4096)
        rc = launch_tile<128>(ctx, ta, Ktot, gx, smem);
    else
        rc = launch_tile<256>(ctx, ta, Ktot, gx, smem);
    if (rc != LSX_OK) return rc;

    VerifyArgs va{};
    va.list = list;
    va.list_count = list_count;
    va.batch = job.batch;
    va.cap = cap;
    va.bar = bar;
    va.K = K;
    va.Ktot = Ktot;
    va.max_rank = job.max_rank;
    va.pivot_slots = pivot_slots;
    va.allow_retry = list_mode ? 0 : 1;
    va.kword = kword;
    va.prof = prof;
    va.rankk = rankk;
    va.sel = sel;
    va.rank_ws = rank_ws;
    va.piv_ws = piv_ws;
    va.status = job.status;
    va.rank_out = job.rank;
    va.piv_out = job.pivot_col;
    va.retry_list = rlist;
    va.retry_count = rcount;
    va.retry_cap = LSX_RETRY_CAP;
    {
        const int64_t nthreads = cap;
        const int bs = 128;
        k_verify<<<(unsigned)((nthreads + bs - 1) / bs), bs, 0, ctx->stream>>>(va);
        ctx->launches++;
    }

    if (job.op != LSX_OP_RANK) {
        if (job.op == LSX_OP_SOLVE) {
            // unused generator/particular slots are zero
            const int nvars = n - 1;
            if (!list_mode) {
                LSX_CUDA_TRY(ctx, cudaMemsetAsync(job.particular, 0, (size_t)job.batch * nvars * job.L * 4, ctx->stream));
                if (job.generators && job.gen_cap > 0)
                    LSX_CUDA_TRY(ctx, cudaMemsetAsync(job.generators, 0,
                                                      (size_t)job.batch * nvars * job.gen_cap * job.L * 4, ctx->stream));
            }
        }
        AsmArgs aa{};
        aa.list = list;
        aa.list_count = list_count;
        aa.batch = job.batch;
        aa.cap = cap;
        aa.op = job.op;
        aa.m = m;
        aa.n = n;
        aa.n_in = job.n_in;
        aa.bar = bar;
        aa.K = K;
        aa.L = job.L;
        aa.ncs = ncs;
        aa.c0 = c0;
        aa.pivot_slots = pivot_slots;
        aa.gen_cap = job.gen_cap;
        aa.primes = ctx->d_primes;
        aa.garner = ctx->d_garner;
        aa.kword = kword;
        aa.res = res;
        aa.dres = dres;
        aa.sel = sel;
        aa.rank_ws = rank_ws;
        aa.piv_ws = piv_ws;
        aa.status = job.status;
        aa.num = job.num;
        aa.den = job.den;
        aa.particular = job.particular;
        aa.generators = job.generators;
        const int64_t nthreads = cap * ((job.op == LSX_OP_SOLVE ? (int64_t)m * (job.gen_cap + 1) : (int64_t)m * ncs) + 1);
        const int bs = 128;
        const unsigned gridn = (unsigned)((nthreads + bs - 1) / bs);
        if (K <= 4)
            k_assemble<4><<<gridn, bs, 0, ctx->stream>>>(aa);
        else if (K <= 8)
            k_assemble<8><<<gridn, bs, 0, ctx->stream>>>(aa);
        else if (K <= 12)                       // 64 x 64 inverse: 12 primes
            k_assemble<12><<<gridn, bs, 0, ctx->stream>>>(aa);
        else if (K <= 16)
            k_assemble<16><<<gridn, bs, 0, ctx->stream>>>(aa);
        else if (K <= 24)                       // 64 x 64 kernel basis: 21 primes (the Horner loop is KT wide)
            k_assemble<24><<<gridn, bs, 0, ctx->stream>>>(aa);
        else
            k_assemble<32><<<gridn, bs, 0, ctx->stream>>>(aa);
        ctx->launches++;
    }
    LSX_CUDA_TRY(ctx, cudaGetLastError());

    if (!list_mode && K > 1) {
        // recompute flagged matrices with replacement primes (a no-op grid when none was flagged; a single prime
        // above twice the bound cannot be flagged)
        rc = run_generic(ctx, job, rlist, rcount, LSX_RETRY_CAP, ws_offset + w.total, kword);
        if (rc != LSX_OK) return rc;
    }
    return LSX_OK;
}

// ---- det(A) mod p for a range of table primes through the tile kernel (n x n fits shared memory) ----
namespace {
__global__ void k_det_fix(const uint32_t* dres, const int32_t* rankk, int n, int count, uint32_t* residues) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) residues[k] = rankk[k] < n ? 0u : dres[k];
}
}  // namespace

bool lsx_tile_fits(int m, int n) { return ((size_t)m * n + n + m + 4) * sizeof(uint32_t) <= 200 * 1024 && m <= 254; }

int lsx_tile_det_residues(lsx_ctx* ctx, const int32_t* dA, int n, int prime_begin, int count, uint32_t* d_res) {
    const size_t smem = ((size_t)n * n + 2 * n + 4) * sizeof(uint32_t);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    const size_t o_dres = take((size_t)count * 4), o_rank = take((size_t)count * 4), o_prof = take((size_t)count * n),
                 o_st = take(256);
    int rc = lsx_ws_reserve(ctx, off);
    if (rc != LSX_OK) return rc;
    char* base = (char*)ctx->d_ws;
    LSX_CUDA_TRY(ctx, cudaMemsetAsync(base + o_st, 0, 4, ctx->stream));
    TileArgs ta{};
    ta.A = dA;
    ta.batch = 1;
    ta.cap = 1;
    ta.m = ta.n = ta.n_in = ta.bar = n;
    ta.a_abs_max = 0x7fffffffLL;
    ta.b_abs_max = 0;
    ta.primes = ctx->d_primes + prime_begin;
    ta.ktot = count;
    ta.dres = (uint32_t*)(base + o_dres);
    ta.rankk = (int32_t*)(base + o_rank);
    ta.prof = (uint8_t*)(base + o_prof);
    ta.status = (int32_t*)(base + o_st);
    const int cells = n * n;
    if (cells <= 1024)
        rc = launch_tile<64>(ctx, ta, count, 1, smem);
    else if (cells <= 4096)
        rc = launch_tile<128>(ctx, ta, count, 1, smem);
    else
        rc = launch_tile<256>(ctx, ta, count, 1, smem);
    if (rc != LSX_OK) return rc;
    k_det_fix<<<(count + 127) / 128, 128, 0, ctx->stream>>>(ta.dres, ta.rankk, n, count, d_res);
    ctx->launches++;
    LSX_CUDA_TRY(ctx, cudaGetLastError());
    return LSX_OK;
}
