/*
 * lsx.h -- C-ABI of the B200 exact elimination engine (liblsx.so).
 *
 * Drop-in boundary for ONE hot path of koskja/linalg-solver: Gauss-Jordan
 * `Matrix.row_reduce` (reference linalg_solver/linalg.py:534-630) and the calls whose
 * results are functions of it: `determinant` (linalg.py:183-262), `inverse` (682-743),
 * `rank` (745-747), `kernel` (749-756), `find_preimage_of` (632-680, 870-999).
 *
 * The reference has no FFI for this path (its only native module, linalg-helper, is a
 * sparsity-pattern planner: linalg-helper/src/lib.rs:122-143).  The functions below are
 * what a binding for the path would call: from Python through ctypes
 * (linalg_solver_b200/_lib.py) or from the Rust crate through an `extern "C"` block
 * (INTEGRATION.md shows both stubs).
 *
 * Conventions
 *  - Plain pointers and sizes only; no C++/torch types.  Every function returns an int:
 *    LSX_OK or a negative LSX_ERR_* code, never throws, never aborts.
 *    lsx_last_error(ctx) gives a human-readable message for the last failure on ctx.
 *  - Inputs are batches of row-major int32 matrices, `batch` of them back to back.
 *  - Outputs are exact.  Every rational result is given as an integer numerator over ONE
 *    common denominator per matrix (the determinant of the pivot minor), each integer as
 *    `limbs` 32-bit little-endian words in two's complement (limbs == 1 is an int32,
 *    limbs == 2 an int64).  `limbs` and the number of 31-bit primes used come from the
 *    lsx_plan_* query (Hadamard bound of the declared entry magnitudes); the caller
 *    allocates all buffers.
 *  - `mem` says where ALL data pointers of the call live: LSX_MEM_DEVICE (device
 *    pointers on ctx's GPU; the call only enqueues work on ctx's stream) or LSX_MEM_HOST
 *    (host pointers; the call copies in, computes, copies out and returns when done).
 *  - Mathematical failure is a per-matrix status bit, not an error code
 *    (reference: `Matrix.NoSolution()`, linalg.py:524-532).  Shape errors return
 *    LSX_ERR_BAD_SHAPE (reference: ValueError, linalg.py:643, 693).
 *  - A ctx is bound to one GPU and one stream and must not be used from two threads at
 *    once; different ctxs are independent.
 */
#ifndef LSX_H
#define LSX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSX_ABI_VERSION 1

/* return codes */
#define LSX_OK                 0
#define LSX_ERR_BAD_SHAPE     -1   /* dimensions/bar_col/limits invalid for the call          */
#define LSX_ERR_CUDA          -2   /* CUDA runtime failure, see lsx_last_error                */
#define LSX_ERR_BOUND         -3   /* declared magnitudes need more primes than supported     */
#define LSX_ERR_NULL          -4   /* required pointer is NULL                                */
#define LSX_ERR_UNSUPPORTED   -5   /* shape outside what the kernels cover                    */
#define LSX_ERR_NO_DEVICE     -6   /* no usable CUDA device                                   */

/* memory space of the data pointers of a call */
#define LSX_MEM_HOST   0
#define LSX_MEM_DEVICE 1

/* per-matrix status bits */
#define LSX_ST_SINGULAR      1   /* inverse: rank < n  -> reference returns NoSolution()      */
#define LSX_ST_INCONSISTENT  2   /* solve: zero left row with non-zero rhs -> NoSolution()    */
#define LSX_ST_BOUND         4   /* an entry exceeded the declared magnitude or the rank      */
                                 /* exceeded max_rank: outputs for this matrix are not valid  */
#define LSX_ST_NO_GOOD_PRIME 8   /* could not collect enough agreeing primes (never expected) */
#define LSX_ST_GEN_TRUNC    16   /* solve: more free variables than gen_cap columns           */
#define LSX_ST_RETRIED      32   /* informational: a bad prime was detected and replaced      */

/* operations a plan describes */
#define LSX_OP_RREF    1
#define LSX_OP_INVERSE 2
#define LSX_OP_DET     3
#define LSX_OP_RANK    4
#define LSX_OP_SOLVE   5

typedef struct lsx_ctx lsx_ctx;

typedef struct lsx_plan {
    int32_t op;            /* LSX_OP_*                                                        */
    int32_t m, n;          /* rows / columns of the matrix the elimination runs on ([A|B])    */
    int32_t bar_col;       /* pivots are searched in columns [0, bar_col); 1 <= bar_col <= n  */
    int32_t max_rank;      /* declared upper bound on the rank (<= min(m, bar_col))           */
    int32_t n_primes;      /* 31-bit primes used per matrix                                   */
    int32_t limbs;         /* 32-bit words per output integer                                 */
    int32_t pivot_slots;   /* = min(m, bar_col): entries of pivot_col per matrix              */
    int32_t gen_cap;       /* solve: generator columns stored per matrix                      */
    int32_t reserved;
    int64_t a_abs_max;     /* declared max |entry| of the left block                          */
    int64_t b_abs_max;     /* declared max |entry| of the right block                         */
    double  log2_bound;    /* log2 of the Hadamard bound every output integer satisfies       */
} lsx_plan;

/* ---- context ------------------------------------------------------------------------ */
int  lsx_abi_version(void);
int  lsx_create(int device_id, lsx_ctx** out);
void lsx_destroy(lsx_ctx* ctx);
/*
 * One context over several GPUs of this process (distinct device ids).  Batched calls with LSX_MEM_HOST buffers are
 * sharded BY MATRIX over the GPUs (contiguous slices, one host thread per GPU, no collective); lsx_det_large shards
 * BY PRIME.  LSX_MEM_DEVICE calls return LSX_ERR_UNSUPPORTED on such a context (device data belongs to one GPU: keep
 * one single-device context per GPU for it).  Stream, timing and launch-count calls address the first GPU.
 */
int  lsx_create_multi(const int* device_ids, int n_dev, lsx_ctx** out);
int  lsx_device_count(const lsx_ctx* ctx);
/* 1 once the context's NCCL communicators exist (created by the first multi-GPU lsx_det_large), 0 before that or when
 * libnccl could not be loaded and the residues are gathered by peer copies instead. */
int  lsx_multi_uses_nccl(const lsx_ctx* ctx);
const char* lsx_last_error(const lsx_ctx* ctx);
/* Run on the caller's CUDA stream (a cudaStream_t passed as void*; NULL = the ctx's own). */
int  lsx_set_stream(lsx_ctx* ctx, void* cuda_stream);
/* Block until everything enqueued on the ctx's stream has finished. */
int  lsx_synchronize(lsx_ctx* ctx);
/* Number of kernels this ctx has launched so far (for launch accounting in benchmarks). */
int64_t lsx_launch_count(const lsx_ctx* ctx);
/* Primes per matrix the most recent tile-path pass of this ctx really used.  The plan's n_primes covers the DECLARED
 * magnitudes; the tile path bounds the minors of the matrices it is given by the product of their largest row norms
 * (Hadamard) and runs only the primes that bound needs -- never more than the plan's.  *out = 0 when the last call
 * did not take the tile path (fused small kernels) or computes a rank.  Waits for the ctx's stream. */
int  lsx_last_prime_count(lsx_ctx* ctx, int* out);
/* Device timing of the DOMINANT kernel of each following call (the elimination kernel): enable
 * records a CUDA event pair around it on the ctx stream; lsx_timing_read waits for the recorded
 * pairs, writes up to `cap` durations in milliseconds (oldest first), returns how many were
 * recorded in *count and clears the list.  Used by bench.py for the roofline figure. */
int  lsx_timing_enable(lsx_ctx* ctx, int enable);
int  lsx_timing_read(lsx_ctx* ctx, float* ms_out, int cap, int* count);
/* Test hook: replace the first `count` primes of the table (odd primes < 2^31, distinct),
 * e.g. tiny primes to force the bad-prime path.  count == 0 restores the default table. */
int  lsx_debug_set_primes(lsx_ctx* ctx, const uint32_t* primes, int count);
/* Copy the first `count` primes of the table to out (host). */
int  lsx_get_primes(const lsx_ctx* ctx, uint32_t* out, int count);

/* ---- plans: how many primes / limbs a call needs -------------------------------------- */
/* row_reduce of an m x n matrix whose columns >= bar_col only receive the row operations
 * (linalg.py:534-630; the `bar_col or n-1` default of linalg.py:543 is applied by the
 * caller).  max_rank <= 0 means unknown. */
int lsx_plan_rref(int m, int n, int bar_col, int64_t a_abs_max, int64_t b_abs_max,
                  int max_rank, lsx_plan* out);
/* inverse of n x n through [A|I], bar_col = n (linalg.py:704-743) */
int lsx_plan_inverse(int n, int64_t a_abs_max, lsx_plan* out);
/* determinant of n x n: sign * product of the forward-sweep pivots (linalg.py:547-609) */
int lsx_plan_det(int n, int64_t a_abs_max, lsx_plan* out);
/* rank of m x n (linalg.py:745-747) */
int lsx_plan_rank(int m, int n, int64_t a_abs_max, lsx_plan* out);
/* find_preimage_of: A (m x n) x = b through [A|b], bar_col = n (linalg.py:648-680,
 * 913-999); gen_cap = generator columns stored per matrix (<= n). */
int lsx_plan_solve(int m, int n, int64_t a_abs_max, int64_t b_abs_max, int max_rank,
                   int gen_cap, lsx_plan* out);

/* ---- batched operations ---------------------------------------------------------------- */
/*
 * row_reduce.  A: [batch][m][n] int32.
 *   num       [batch][m*n][limbs]        numerators: d * RREF entry
 *   den       [batch][limbs]             d = determinant of the pivot minor (never 0)
 *   pivot_col [batch][pivot_slots] int32 column of pivot k (row k), -1 padded
 *   rank      [batch] int32
 *   status    [batch] int32              LSX_ST_* bits
 * Row choice follows the reference exactly (entry at the pivot position if non-zero,
 * else the first lower non-zero row, linalg.py:548-567), so the right block of a
 * rank-deficient input matches the reference too.
 */
int lsx_rref_batch(lsx_ctx* ctx, const lsx_plan* plan, const int32_t* A, int64_t batch,
                   int mem, uint32_t* num, uint32_t* den, int32_t* pivot_col,
                   int32_t* rank, int32_t* status);

/*
 * inverse + determinant.  A: [batch][n][n] int32.
 *   adj    [batch][n*n][limbs]   A^-1 = adj / det   (zeros when singular)
 *   det    [batch][limbs]        determinant (0 when singular)
 *   status [batch]               LSX_ST_SINGULAR where the reference returns NoSolution()
 */
int lsx_inverse_batch(lsx_ctx* ctx, const lsx_plan* plan, const int32_t* A, int64_t batch,
                      int mem, uint32_t* adj, uint32_t* det, int32_t* status);

/*
 * Same with the input matrices as int8 ([batch][n][n] bytes, for entries in [-127, 127]): a quarter of the
 * host-to-device traffic of the call, which matters because LSX_MEM_HOST calls on small matrices are PCIe bound.
 * Served by the fused register-resident kernel only (n <= 8 and one prime / one limb in the plan); other plans
 * return LSX_ERR_UNSUPPORTED and the caller widens to int32.
 */
int lsx_inverse_batch_i8(lsx_ctx* ctx, const lsx_plan* plan, const int8_t* A, int64_t batch,
                         int mem, uint32_t* adj, uint32_t* det, int32_t* status);

/* determinant (and rank, may be NULL).  A: [batch][n][n].  det: [batch][limbs]. */
int lsx_det_batch(lsx_ctx* ctx, const lsx_plan* plan, const int32_t* A, int64_t batch,
                  int mem, uint32_t* det, int32_t* rank, int32_t* status);

/* rank.  A: [batch][m][n].  rank: [batch]. */
int lsx_rank_batch(lsx_ctx* ctx, const lsx_plan* plan, const int32_t* A, int64_t batch,
                   int mem, int32_t* rank, int32_t* status);

/*
 * find_preimage_of.  A: [batch][m][n], b: [batch][m].
 *   den        [batch][limbs]                common denominator d
 *   particular [batch][n][limbs]             numerators; free variables are 0
 *   generators [batch][n][gen_cap][limbs]    numerators, column t belongs to the t-th free
 *                                            column in ascending order: entry d at the free
 *                                            column, -d*R[i][f] at pivot columns
 *                                            (linalg.py:973-983); unused columns are 0
 *   pivot_col  [batch][pivot_slots], rank [batch], status [batch]
 *   status has LSX_ST_INCONSISTENT where the reference returns NoSolution().
 */
int lsx_solve_batch(lsx_ctx* ctx, const lsx_plan* plan, const int32_t* A, const int32_t* b,
                    int64_t batch, int mem, uint32_t* den, uint32_t* particular,
                    uint32_t* generators, int32_t* pivot_col, int32_t* rank,
                    int32_t* status);

/* ---- lowest terms ------------------------------------------------------------------------------------------ */
/*
 * The batched operations return numerators over ONE common denominator per matrix; the reference returns reduced
 * rationals (true division on Rational entries, linalg.py:574; sympy.Matrix.inv(), linalg.py:698-699).  This call
 * reduces on the device: for every matrix b and entry i < count
 *   p[b][i] / q[b][i] = num[b][i] / den[b],  gcd(p, q) = 1,  q > 0      (binary gcd + exact division, multi-limb)
 *   num [batch][count][limbs], den [batch][limbs]  ->  p, q [batch][count][limbs]   (two's complement words)
 * num == 0 gives 0 / 1; den == 0 (a matrix flagged singular / inconsistent) gives 0 / 0.  limbs <= 32.
 */
int lsx_lowest_terms(lsx_ctx* ctx, const uint32_t* num, const uint32_t* den, int64_t batch, int count,
                     int limbs, int mem, uint32_t* p, uint32_t* q);

/* ---- step trace of row_reduce (reference linalg.py:544-629: intermediate_matrices / intermediate_steps) ------ */
/* Upper bound on the number of recorded steps: per pivot S, N, E (below) and E (above). */
int lsx_rref_trace_max_ops(int m, int n, int bar_col);
/*
 * Replays row_reduce of ONE m x n matrix (m * n <= 4096) modulo the first n_primes table primes in exactly the
 * reference's operation order and records every step the reference records:
 *   ops       [n_primes][max_ops][4] int32   kind (1 = S row swap: a, b = the two rows; 2 = N normalisation of row a;
 *                                            3 = E elimination below the pivot of column a; 4 = E above), a, b, 0
 *   frames    [n_primes][max_ops][m][n]      residues (plain, in [0, p)) of the matrix after each step
 *   n_ops     [n_primes]                     number of recorded steps (identical for all primes unless a prime divides
 *                                            an intermediate value: the caller compares the logs)
 *   pivot_col [n_primes][min(m, bar_col)]    pivot column of row k, -1 padded
 * The caller lifts the residues to rationals (CRT + rational reconstruction); every intermediate entry is a
 * quotient of minors of A, so n_primes follows from twice the Hadamard bound.  Not a throughput path.
 */
int lsx_rref_trace(lsx_ctx* ctx, const int32_t* A, int m, int n, int bar_col, int n_primes, int mem,
                   int32_t* ops, uint32_t* frames, int32_t* n_ops, int32_t* pivot_col);
/*
 * Same for a RATIONAL matrix A / D (row_reduce accepts any Fraction matrix, linalg.py:534-630): A holds the integer
 * numerators over one common denominator D > 0 and den_residues[k] = D mod (k-th table prime) ([n_primes], in `mem`;
 * NULL means D = 1).  The replay runs on the residues of A / D, so "pivot is 1" is tested on the rational entry as
 * the reference does.  A prime that divides D reports n_ops[k] = -1 and is to be dropped by the caller.
 */
int lsx_rref_trace_q(lsx_ctx* ctx, const int32_t* A, const uint32_t* den_residues, int m, int n, int bar_col,
                     int n_primes, int mem, int32_t* ops, uint32_t* frames, int32_t* n_ops, int32_t* pivot_col);

/* ---- one large determinant, shardable by prime ------------------------------------------ */
/* Number of primes the determinant of an n x n matrix with |entries| <= a_abs_max needs. */
int lsx_det_large_prime_count(int n, int64_t a_abs_max, int* n_primes, double* log2_bound);
/*
 * Same, from the matrix itself: Hadamard's bound with the ACTUAL row and column norms,
 * |det A| <= min(prod_i ||row_i||_2, prod_j ||col_j||_2), which for random entries is a few per cent tighter
 * than n * log2(sqrt(n) * a_abs_max) and still rigorous (every rank must use the same count: the bound only
 * depends on A).  A: [n][n] int32 in `mem`.  A zero row or column gives n_primes = 1, log2_bound = 0.
 */
int lsx_det_large_prime_count_for(lsx_ctx* ctx, const int32_t* A, int n, int mem, int* n_primes,
                                  double* log2_bound);
/*
 * det(A) mod p for the primes [prime_begin, prime_begin + prime_count) of the table.
 * A: [n][n] int32.  residues: [prime_count] uint32 (plain residues), primes_out (may be
 * NULL): the primes used.  Ranks shard the prime range and all-gather `residues`.
 */
int lsx_det_large_residues(lsx_ctx* ctx, const int32_t* A, int n, int prime_begin,
                           int prime_count, int mem, uint32_t* residues,
                           uint32_t* primes_out);
/*
 * rank of ONE m x n int32 matrix of any size that fits device memory (reference rank(), linalg.py:745-747, has no
 * size limit; lsx_rank_batch keeps one residue tile per CTA, m <= 254).  Fraction-free elimination in global memory
 * modulo groups of table primes; the answer is the largest rank seen, which is the rank over Q once the primes'
 * product exceeds the Hadamard bound of the minors (full rank modulo one prime ends the loop early).
 * A follows `mem`; rank and primes_used (may be NULL: how many primes were run) are HOST pointers.
 */
int lsx_rank_large(lsx_ctx* ctx, const int32_t* A, int m, int n, int mem, int32_t* rank, int32_t* primes_used);
/*
 * The whole by-prime determinant behind one call (BASELINE.json configs[4]): Hadamard bound from the row / column
 * norms -> prime count K; the primes are sharded over the context's GPUs (each gets its own copy of A), the residue
 * vectors are all-gathered over NVLink (NCCL, loaded at run time; peer copies if it is unavailable), the Garner
 * CRT runs on the first GPU.  A: [n][n] int32 in HOST memory.  det_words: HOST, room for limbs_cap words; the
 * signed determinant comes back as *limbs_out little-endian two's-complement words (LSX_ERR_BOUND if limbs_cap is too
 * small: *limbs_out then says how many are needed).  n_primes_out may be NULL.  Works on a single-device context too.
 */
int lsx_det_large(lsx_ctx* ctx, const int32_t* A, int n, int limbs_cap, uint32_t* det_words, int* limbs_out,
                  int* n_primes_out);
/*
 * CRT of `count` residues (for table primes [0, count)) to a signed integer of `limbs`
 * words (two's complement, little endian).  residues/out follow `mem`.
 */
int lsx_crt_signed(lsx_ctx* ctx, const uint32_t* residues, int count, int limbs, int mem,
                   uint32_t* out);

#ifdef __cplusplus
}
#endif
#endif /* LSX_H */
