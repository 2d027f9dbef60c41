// Tensor-core trailing update of the blocked modular LU (config 5 of BASELINE.json), sm_100a only.
//
//   C[r][c] <- (C[r][c] * 2^32 + sum_k L[r][k] * U[k][c]) / 2^32  mod p      r in [r0,r1), c in [c0,c1), k in [k0,k0+K)
//
// with L = W[.][k0..k0+K) (the negated Montgomery multipliers left by the panel kernels) and
// U = W[k0..k0+K)[.], for G primes side by side (W is [G][n][n] words).  This is the dense contraction of
// the forward sweep of reference linalg.py:587-596 (row_i -= factor * pivot_row) for a whole block of
// pivots at once, and the only genuinely GEMM-shaped piece of the path.
//
// 31-bit residues do not fit a tensor-core operand, so each word is split into four unsigned bytes
// (k_tc_split, written straight in the no-swizzle K-major core-matrix layout tcgen05 wants) and the 16
// byte-plane products La * Ub^T accumulate, by weight t = a + b, into SEVEN s32 TMEM accumulators
// Q_t = sum_{a+b=t} La * Ub^T.  Every Q_t stays below 4 * K * 255^2 < 2^26 for K <= 256, so the s32
// accumulation is exact.  The epilogue reads the seven accumulators with tcgen05.ld, forms
// S = sum_t Q_t * (2^(8t) mod p) < 2^60 in 64 bits and finishes with one Montgomery reduction:
// C + REDC(S) mod p  -- bit-identical to the integer-pipe k_gemm_int.
//
// Instruction shape (the point of this layout; measurements in profiles/r01i_tc_gemm_experiments.md): a
// tcgen05.mma of M = 128, N, K = 32 costs about (4096 + 32 N) B / 110 B/clk of operand fetch, so the 16
// products issued as 16 instructions of N = 64 run at half the tensor peak.  Here the four byte planes of
// the B tile are STACKED along N in shared memory ([k16][plane b][column][16 B]): one instruction per A
// plane a with N = 4 * TN covers the accumulators a .. a+3, which are adjacent in TMEM -- 4 instructions
// per K step instead of 16, each long enough to be MAC bound.  (The very first K step of a tile splits
// three of them so that every accumulator's first write has accumulate = 0.)  With TN = 32 the seven
// accumulators take 224 columns, so TWO tiles fit TMEM and the epilogue of one tile overlaps the MMAs of
// the next.
//
// CTA = (prime g, 128-row tile, a contiguous group of 32-column tiles) or, for tall narrow regions, (prime,
// 32-column tile, group of row tiles).  The planes of the fixed tile stay resident in shared memory for the
// whole CTA; the other operand streams through a ring of stages filled by cp.async.bulk (TMA engine, SASS
// UBLKCP) and released by tcgen05.commit.  Warp 0 = TMEM allocator + copy producer, warp 1 = MMA issuer
// (converged warp, one elected lane), warps 2..9 = epilogue (two warps per TMEM lane quadrant, 16 columns each).
#pragma once
#include "lsx_internal.h"

namespace lsx_tc {

constexpr int TM = 128;            // rows per tile (= TMEM lanes)
constexpr int TN = 32;             // columns per tile; the MMA N is 4 * TN = 128 (four stacked byte planes)
constexpr int KC = 64;             // contraction bytes per smem stage (32 when K == 32)
constexpr int A_CHUNK = 4 * TM * KC;   // 32768 B: four byte planes of a 128 x 64 slice
constexpr int B_CHUNK = 4 * TN * KC;   // 8192 B
constexpr int STAGES_A = 8;        // ring slots when the A planes are stationary (8 KB B slots)
constexpr int STAGES_B = 4;        // ring slots when the B planes are stationary (32 KB A slots)
constexpr int THREADS = 320;
constexpr int ACC_COLS = 7 * TN;   // 224 TMEM columns per tile
constexpr int ACC_STRIDE = 256;    // second accumulator buffer
constexpr int TMEM_COLS = 512;
constexpr int MAX_K = 256;

struct Region {
    int n;            // order of W (row stride in words)
    int r0, r1;       // rows of C updated
    int c0, c1;       // columns of C updated
    int k0, K;        // contraction range; K is 32 or a multiple of 64, at most MAX_K
    int kc;           // contraction bytes per stage: min(K, 64)
    int m_tiles;      // tiles of 128 rows
    int n_tiles;      // tiles of 32 columns
    int tiles_per_cta;  // tiles of the looped dimension handled by one CTA
    int b_stationary;   // 0: CTA keeps the A planes of one row tile and loops over column tiles; 1: the reverse
    void set_tiles() {
        m_tiles = (r1 - r0 + TM - 1) / TM;
        n_tiles = (c1 - c0 + TN - 1) / TN;
    }
};

inline size_t a_plane_bytes(const Region& g) { return (size_t)g.m_tiles * g.K * 4 * TM; }   // per prime
inline size_t b_plane_bytes(const Region& g) { return (size_t)g.n_tiles * g.K * 4 * TN; }   // per prime
constexpr int EPI_ROW_BYTES = 80;  // row pitch of the epilogue's staging blocks (64 data bytes + 16: conflict-free)
constexpr int EPI_STAGE = 8 * 32 * EPI_ROW_BYTES;   // eight epilogue warps x 32 rows
inline size_t smem_bytes(int K, int b_stationary) {
    return (b_stationary ? (size_t)K * 4 * TN + (size_t)STAGES_B * A_CHUNK : (size_t)K * 4 * TM + (size_t)STAGES_A * B_CHUNK) +
           256 + EPI_STAGE;
}
inline bool depth_ok(int K) { return K == 32 || (K % KC == 0 && K >= KC && K <= MAX_K); }

#ifdef __CUDACC__

// ---- byte-plane split ---------------------------------------------------------------------------------
// A planes (rows of L), prime g, row tile ti:      [K/kc][plane a][k16 = kc/16][row = 128][16 B]
// B planes (columns of U), prime g, column tile tj: [K/kc][k16 = kc/16][plane b][col = 32][16 B]
// (both no-swizzle K-major core matrices; the B planes are stacked along N inside each 16-byte K group)
__device__ __forceinline__ void split_store(uint8_t* dst, int plane_stride, const uint32_t (&w)[16]) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            o[i] = ((w[4 * i] >> (8 * a)) & 255u) | (((w[4 * i + 1] >> (8 * a)) & 255u) << 8) |
                   (((w[4 * i + 2] >> (8 * a)) & 255u) << 16) | (((w[4 * i + 3] >> (8 * a)) & 255u) << 24);
        *reinterpret_cast<uint4*>(dst + a * plane_stride) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}
__device__ __forceinline__ void split_a_body(const uint32_t* __restrict__ W, uint8_t* __restrict__ AP, const Region& g,
                                             int64_t block, int prime) {
    const int q_per = g.K / 16;                       // 16-byte k groups
    const int64_t per_prime = (int64_t)g.m_tiles * TM * q_per;
    const int64_t t = block * blockDim.x + threadIdx.x;
    if (t >= per_prime) return;
    const int r = (int)(t % TM);
    const int q = (int)((t / TM) % q_per);
    const int ti = (int)(t / ((int64_t)TM * q_per));
    const int row = g.r0 + ti * TM + r;
    uint32_t w[16];
    if (row < g.r1) {
        const uint32_t* src = W + ((int64_t)prime * g.n + row) * g.n + g.k0 + q * 16;
        if ((g.n & 3) == 0 && (g.k0 & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint4 v = *reinterpret_cast<const uint4*>(src + 4 * i);
                w[4 * i] = v.x, w[4 * i + 1] = v.y, w[4 * i + 2] = v.z, w[4 * i + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) w[i] = src[i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = 0u;
    }
    const int qc = g.kc / 16;                          // 16-byte groups per stage
    uint8_t* dst = AP + ((int64_t)prime * g.m_tiles + ti) * ((int64_t)g.K * 4 * TM) + (int64_t)(q / qc) * (4 * TM * g.kc) +
                   (q % qc) * (TM * 16) + r * 16;
    split_store(dst, TM * g.kc, w);
}
__device__ __forceinline__ void split_b_body(const uint32_t* __restrict__ W, uint8_t* __restrict__ BP, const Region& g,
                                             int64_t block, int prime) {
    const int q_per = g.K / 16;
    const int64_t per_prime = (int64_t)g.n_tiles * TN * q_per;
    const int64_t t = block * blockDim.x + threadIdx.x;
    if (t >= per_prime) return;
    const int c = (int)(t % TN);
    const int q = (int)((t / TN) % q_per);
    const int tj = (int)(t / ((int64_t)TN * q_per));
    const int col = g.c0 + tj * TN + c;
    uint32_t w[16];
    if (col < g.c1) {
        const uint32_t* src = W + ((int64_t)prime * g.n + g.k0 + q * 16) * g.n + col;
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = src[(int64_t)i * g.n];
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = 0u;
    }
    const int qc = g.kc / 16;
    uint8_t* dst = BP + ((int64_t)prime * g.n_tiles + tj) * ((int64_t)g.K * 4 * TN) + (int64_t)(q / qc) * (4 * TN * g.kc) +
                   (q % qc) * (4 * TN * 16) + c * 16;
    split_store(dst, TN * 16, w);
}
// ONE launch splits both operands (the first blocks_a blocks of a prime do the A planes, the rest the B planes): the two
// are independent, and inside a panel each of them is a few blocks per prime that did not fill the GPU on its own.
__global__ void __launch_bounds__(256) k_tc_split(const uint32_t* __restrict__ W, uint8_t* __restrict__ AP,
                                                  uint8_t* __restrict__ BP, Region g, int blocks_a) {
    if ((int)blockIdx.x < blocks_a) split_a_body(W, AP, g, blockIdx.x, blockIdx.y);
    else split_b_body(W, BP, g, (int64_t)blockIdx.x - blocks_a, blockIdx.y);
}
inline void launch_split(const uint32_t* W, uint8_t* AP, uint8_t* BP, const Region& g, int G, cudaStream_t st) {
    const int64_t ta = (int64_t)g.m_tiles * TM * (g.K / 16), tb = (int64_t)g.n_tiles * TN * (g.K / 16);
    const int blocks_a = (int)((ta + 255) / 256), blocks_b = (int)((tb + 255) / 256);
    k_tc_split<<<dim3((unsigned)(blocks_a + blocks_b), G), 256, 0, st>>>(W, AP, BP, g, blocks_a);
}

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol error traps (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 26)) __trap();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, u8 x u8 -> s32, M = 128, K = 32 per instruction, N from the instruction descriptor
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, no swizzle, K-major: element (row, k-byte) lives at
// (row % 8) * 16 + (row / 8) * SBO + (k / 16) * LBO + (k % 16).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

struct GemmArgs {
    uint32_t* W;                // [G][n][n]
    const uint8_t* AP;          // byte planes of L, see split_a_body
    const uint8_t* BP;          // byte planes of U, see split_b_body
    const PrimeRec* primes;     // [G]
    Region g;
};

// DBG (timing experiments of tools/tc_gemm_test only; the library instantiates DBG = 0):
//   bit 1 skip the epilogue arithmetic, bit 2 skip the TMEM loads
template <int DBG, int COAL>
__global__ void __launch_bounds__(THREADS, 1) k_gemm_tc_t(GemmArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const Region& g = a.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int prime = blockIdx.z;
    const bool bst = g.b_stationary != 0;
    const int fixed_tile = blockIdx.x;                                   // row tile (A stationary) or column tile
    const int loop_tiles = bst ? g.m_tiles : g.n_tiles;
    const int t0 = blockIdx.y * g.tiles_per_cta;
    const int ntiles = min(loop_tiles, t0 + g.tiles_per_cta) - t0;
    const int kchunks = g.K / g.kc;                                      // <= 4
    const uint32_t a_chunk = 4u * TM * g.kc, b_chunk = 4u * TN * g.kc;   // bytes per K chunk (four byte planes)
    const uint32_t a_plane = TM * g.kc;                                  // bytes per A byte plane inside a chunk
    const int ksteps = g.kc / 32;                                        // MMA K = 32 bytes
    const uint32_t st_chunk = bst ? b_chunk : a_chunk;                   // stationary operand, per chunk
    const uint32_t rg_chunk = bst ? a_chunk : b_chunk;                   // streamed operand, per chunk
    const uint32_t rg_stride = bst ? (uint32_t)A_CHUNK : (uint32_t)B_CHUNK;   // ring slot size
    const uint32_t STAGES = bst ? STAGES_B : STAGES_A;

    uint8_t* smS = smem;                                                 // stationary planes: kchunks * st_chunk
    uint8_t* smR = smem + (size_t)kchunks * st_chunk;                    // ring: STAGES slots
    uint64_t* bars = reinterpret_cast<uint64_t*>(smR + (size_t)STAGES * rg_stride);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);
    const uint32_t bar_s_full = smem_u32(bars + 0);
    const uint32_t bar_acc_full = smem_u32(bars + 2);                    // [2]
    const uint32_t bar_acc_empty = smem_u32(bars + 4);                   // [2]
    const uint32_t bar_r_full = smem_u32(bars + 8);                      // [STAGES]
    const uint32_t bar_r_empty = smem_u32(bars + 8 + STAGES_A);          // [STAGES]

    if (ntiles <= 0) return;                                             // uniform per CTA
    if (warp == 0) {
        if (lane == 0) {
            mbar_init(bar_s_full, 1);
            for (int b = 0; b < 2; ++b) {
                mbar_init(bar_acc_full + 8 * b, 1);
                mbar_init(bar_acc_empty + 8 * b, 8);
            }
            for (uint32_t s = 0; s < STAGES; ++s) {
                mbar_init(bar_r_full + 8 * s, 1);
                mbar_init(bar_r_empty + 8 * s, 1);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== copy producer =====
        if (lane == 0) {
            const uint8_t* baseA = a.AP + (int64_t)prime * g.m_tiles * ((int64_t)g.K * 4 * TM);
            const uint8_t* baseB = a.BP + (int64_t)prime * g.n_tiles * ((int64_t)g.K * 4 * TN);
            const uint8_t* srcS = bst ? baseB + (int64_t)fixed_tile * ((int64_t)g.K * 4 * TN)
                                      : baseA + (int64_t)fixed_tile * ((int64_t)g.K * 4 * TM);
            mbar_expect_tx(bar_s_full, (uint32_t)kchunks * st_chunk);
            for (int kc = 0; kc < kchunks; ++kc)
                bulk_g2s(smem_u32(smS + (size_t)kc * st_chunk), srcS + (size_t)kc * st_chunk, st_chunk, bar_s_full);
            uint32_t cnt = 0;
            for (int t = 0; t < ntiles; ++t) {
                const uint8_t* srcR = bst ? baseA + (int64_t)(t0 + t) * ((int64_t)g.K * 4 * TM)
                                          : baseB + (int64_t)(t0 + t) * ((int64_t)g.K * 4 * TN);
                for (int kc = 0; kc < kchunks; ++kc, ++cnt) {
                    const uint32_t slot = cnt % STAGES, phase = (cnt / STAGES) & 1u;
                    mbar_wait(bar_r_empty + 8 * slot, phase ^ 1u);
                    mbar_expect_tx(bar_r_full + 8 * slot, rg_chunk);
                    bulk_g2s(smem_u32(smR + (size_t)slot * rg_stride), srcR + (size_t)kc * rg_chunk, rg_chunk,
                             bar_r_full + 8 * slot);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the pipeline (uniform control flow keeps the descriptors in
        // uniform registers), one elected lane issues the tcgen05 instructions =====
        // instruction descriptor: D = s32 (2 << 4), A = B = u8 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t idesc0 = (2u << 4) | ((uint32_t)(TM >> 4) << 24);
        const uint32_t idesc_n4 = idesc0 | ((uint32_t)((4 * TN) >> 3) << 17);     // all four planes: N = 128
        const uint32_t idesc_n3 = idesc0 | ((uint32_t)((3 * TN) >> 3) << 17);     // planes 0..2:     N = 96
        const uint32_t idesc_n1 = idesc0 | ((uint32_t)(TN >> 3) << 17);           // plane 3 alone:   N = 32
        // operand bases: + (byte offset >> 4) selects chunk / plane / K step.  A: LBO (next 16-byte K group) =
        // 128 rows * 16 B; B: the four stacked planes make 4 * TN rows per K group.  SBO (next 8 rows) = 128 B.
        const uint64_t adesc0 = smem_desc(smem_u32(bst ? smR : smS), TM * 16, 128);
        const uint64_t bdesc0 = smem_desc(smem_u32(bst ? smS : smR), 4 * TN * 16, 128);
        const uint32_t a_step = bst ? rg_stride : a_chunk;               // distance between K chunks of A
        const uint32_t b_step = bst ? b_chunk : rg_stride;
        mbar_wait(bar_s_full, 0);
        tc_fence_after();
        uint32_t cnt = 0;
        for (int t = 0; t < ntiles; ++t, cnt += kchunks) {
            const uint32_t buf = (uint32_t)t & 1u, use = (uint32_t)t >> 1;
            mbar_wait(bar_acc_empty + 8 * buf, (use & 1u) ^ 1u);         // the epilogue has drained this buffer
            tc_fence_after();
            const uint32_t d0 = tmem_base + buf * ACC_STRIDE;
            for (int kc = 0; kc < kchunks; ++kc) {
                const uint32_t slot = (cnt + kc) % STAGES, phase = ((cnt + kc) / STAGES) & 1u;
                mbar_wait(bar_r_full + 8 * slot, phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t ad_c = adesc0 + (((bst ? slot : (uint32_t)kc) * a_step) >> 4);
                    const uint64_t bd_c = bdesc0 + (((bst ? (uint32_t)kc : slot) * b_step) >> 4);
                    for (int s = 0; s < ksteps; ++s) {
                        const uint64_t ad_s = ad_c + (((uint32_t)s * (2 * TM * 16)) >> 4);
                        const uint64_t bd_s = bd_c + (((uint32_t)s * (2 * 4 * TN * 16)) >> 4);
                        if ((kc | s) == 0) {
                            // first K step of the tile: every accumulator's first write must not accumulate.
                            // plane a = 0 writes accumulators 0..3 fresh; for a = 1, 2, 3 the planes b = 0..2
                            // accumulate into a..a+2 (already written) and b = 3 opens accumulator a + 3.
                            tc_mma_i8(d0, ad_s, bd_s, idesc_n4, 0u);
#pragma unroll
                            for (int pa = 1; pa < 4; ++pa) {
                                const uint64_t ad = ad_s + (((uint32_t)pa * a_plane) >> 4);
                                tc_mma_i8(d0 + (uint32_t)(pa * TN), ad, bd_s, idesc_n3, 1u);
                                tc_mma_i8(d0 + (uint32_t)((pa + 3) * TN), ad, bd_s + ((3u * TN * 16) >> 4), idesc_n1, 0u);
                            }
                        } else {
#pragma unroll
                            for (int pa = 0; pa < 4; ++pa)
                                tc_mma_i8(d0 + (uint32_t)(pa * TN), ad_s + (((uint32_t)pa * a_plane) >> 4), bd_s, idesc_n4, 1u);
                        }
                    }
                    tc_commit(bar_r_empty + 8 * slot);                   // frees the slot when these MMAs have read it
                    if (kc == kchunks - 1) tc_commit(bar_acc_full + 8 * buf);
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue: 8 warps; warp w reads TMEM lanes 32 * (w % 4) .. + 31, 16 of the 32 columns =====
        // Two thread-to-cell mappings.  TMEM hands lane l the 16 columns of ROW l, and in that mapping a 16-byte access to
        // C touches 32 different rows per warp instruction (32 cache lines for 512 bytes): ncu showed the L1 data pipe
        // busy with exactly these accesses.  So the accumulators are reduced in the row mapping to T = S / 2^32 mod p
        // (which does not need C:  redc(C * 2^32 + S) = C + redc(S) mod p, both canonical), T goes through a per-warp
        // staging block in shared memory, and C is loaded, corrected and stored in a COALESCED mapping: lane l owns the
        // 16-byte chunk l / 8 of the rows 8 i + l % 8, i = 0..3 -- 8 rows x 64 contiguous bytes per instruction.
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const PrimeRec P = a.primes[prime];
        const uint32_t p = P.p, pinv = P.pinv;
        const uint64_t c4 = P.one;                               // 2^32 mod p
        const uint64_t c5 = (c4 << 8) % p, c6 = (c5 << 8) % p;   // 2^40, 2^48 mod p
        uint32_t* Wg = a.W + (int64_t)prime * g.n * g.n;
        const bool vec_ok = (g.n & 3) == 0 && (g.c0 & 3) == 0;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 16);
        if constexpr (COAL) {
            uint8_t* stage = reinterpret_cast<uint8_t*>(bars) + 256 + (size_t)(warp - 2) * (32 * EPI_ROW_BYTES);
            const int crow = lane & 7, cchunk = lane >> 3;
            // rows 8 i + crow (i = 0..3) of this warp's 32 x 16 piece of tile t, words 4 cchunk .. + 3
            auto piece_ptr = [&](int t, int i, bool& live, bool& full, int& colb) -> uint32_t* {
                const int ti = bst ? t0 + t : fixed_tile, tj = bst ? fixed_tile : t0 + t;
                const int row = g.r0 + ti * TM + quad * 32 + 8 * i + crow;
                colb = g.c0 + tj * TN + half * 16 + 4 * cchunk;
                live = t < ntiles && row < g.r1 && colb < g.c1;
                full = live && vec_ok && colb + 4 <= g.c1;
                return Wg + (int64_t)row * g.n + colb;
            };
            // C does not depend on the MMAs, and a warp handles its tiles one after the other: the loads of a tile are
            // issued TWO tiles ahead (three register buffers), so their HBM/L2 latency (about as long as a whole tile
            // takes) is hidden instead of being paid once per tile.
            auto load_c = [&](int t, uint32_t (&cv)[16]) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    bool live, full;
                    int colb;
                    const uint32_t* cp = piece_ptr(t, i, live, full, colb);
                    if (full) {
                        const uint4 v = *reinterpret_cast<const uint4*>(cp);
                        cv[4 * i] = v.x, cv[4 * i + 1] = v.y, cv[4 * i + 2] = v.z, cv[4 * i + 3] = v.w;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) cv[4 * i + e] = (live && colb + e < g.c1) ? cp[e] : 0u;
                    }
                }
            };
            auto process = [&](int t, uint32_t (&cv)[16]) {
                const uint32_t buf = (uint32_t)t & 1u, use = (uint32_t)t >> 1;
                mbar_wait(bar_acc_full + 8 * buf, use & 1u);
                tc_fence_after();
                uint32_t q[7][16];
#pragma unroll
                for (int s = 0; s < 7; ++s) {
                    if (DBG & 4) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) q[s][i] = 0u;
                    } else {
                        tmem_ld16(lane_base + buf * ACC_STRIDE + (uint32_t)(s * TN), q[s]);
                    }
                }
                tmem_wait_ld();
                tc_fence_before();                                   // all accumulator words of this warp are in registers:
                __syncwarp();                                        // hand the buffer back (the MMAs of tile t + 2 reuse it);
                if (lane == 0) mbar_arrive(bar_acc_empty + 8 * buf); // also: every lane is done reading the previous tile's stage
                uint32_t tv[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (DBG & 2) {
                        tv[i] = q[0][i] + q[6][i];
                    } else {
                        // S < 2^60; one conditional subtraction of p * 2^32 (as in mac_lazy) covers primes below 2^28 too,
                        // then REDC gives S / 2^32 mod p in [0, p)
                        uint64_t acc = (uint64_t)q[0][i] + ((uint64_t)q[1][i] << 8) + ((uint64_t)q[2][i] << 16) +
                                       ((uint64_t)q[3][i] << 24) + (uint64_t)q[4][i] * c4 + (uint64_t)q[5][i] * c5 +
                                       (uint64_t)q[6][i] * c6;
                        uint32_t hi = (uint32_t)(acc >> 32);
                        hi = min(hi, hi - p);
                        acc = ((uint64_t)hi << 32) | (uint32_t)acc;
                        tv[i] = mont_redc(acc, p, pinv);
                    }
                }
                // row mapping -> coalesced mapping (rows are EPI_ROW_BYTES apart: both the 16-byte row writes of a quarter
                // warp and the reads of 8 consecutive rows at one chunk hit 8 different bank groups)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    *reinterpret_cast<uint4*>(stage + lane * EPI_ROW_BYTES + 16 * k) =
                        make_uint4(tv[4 * k], tv[4 * k + 1], tv[4 * k + 2], tv[4 * k + 3]);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint4 v = *reinterpret_cast<const uint4*>(stage + (8 * i + crow) * EPI_ROW_BYTES + 16 * cchunk);
                    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (DBG & 2) {
                            cv[4 * i + e] += w4[e];
                        } else {
                            const uint32_t sum = cv[4 * i + e] + w4[e];              // both below p < 2^31
                            cv[4 * i + e] = min(sum, sum - p);
                        }
                    }
                    bool live, full;
                    int colb;
                    uint32_t* cp = piece_ptr(t, i, live, full, colb);
                    if (full) {
                        *reinterpret_cast<uint4*>(cp) = make_uint4(cv[4 * i], cv[4 * i + 1], cv[4 * i + 2], cv[4 * i + 3]);
                    } else if (live) {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (colb + e < g.c1) cp[e] = cv[4 * i + e];
                    }
                }
            };
            uint32_t cva[16], cvb[16], cvc[16];
            load_c(0, cva);
            load_c(1, cvb);
            for (int t = 0; t < ntiles; t += 3) {
                load_c(t + 2, cvc);
                process(t, cva);
                if (t + 1 < ntiles) {
                    load_c(t + 3, cva);
                    process(t + 1, cvb);
                }
                if (t + 2 < ntiles) {
                    load_c(t + 4, cvb);
                    process(t + 2, cvc);
                }
            }
        } else {
            // Tile t of this thread: 16 consecutive words of one row of C.
            auto tile_ptr = [&](int t, bool& live, bool& full) -> uint32_t* {
                const int ti = bst ? t0 + t : fixed_tile, tj = bst ? fixed_tile : t0 + t;
                const int row = g.r0 + ti * TM + quad * 32 + lane;
                const int colb = g.c0 + tj * TN + half * 16;
                live = t < ntiles && row < g.r1 && colb < g.c1;
                full = live && vec_ok && colb + 16 <= g.c1;
                return Wg + (int64_t)row * g.n + colb;
            };
            // C does not depend on the MMAs, and a warp handles its tiles one after the other: the loads of a tile are
            // issued TWO tiles ahead (three register buffers), so their HBM/L2 latency (about as long as a whole tile
            // takes) is hidden instead of being paid once per tile.
            auto load_c = [&](int t, uint32_t (&cv)[16]) {
                bool live, full;
                const uint32_t* cp = tile_ptr(t, live, full);
                if (full) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint4 v = *reinterpret_cast<const uint4*>(cp + 4 * i);
                        cv[4 * i] = v.x, cv[4 * i + 1] = v.y, cv[4 * i + 2] = v.z, cv[4 * i + 3] = v.w;
                    }
                } else {
                    const int colb = g.c0 + (bst ? fixed_tile : t0 + t) * TN + half * 16;
#pragma unroll
                    for (int i = 0; i < 16; ++i) cv[i] = (live && colb + i < g.c1) ? cp[i] : 0u;
                }
            };
            auto process = [&](int t, uint32_t (&cv)[16]) {
                const uint32_t buf = (uint32_t)t & 1u, use = (uint32_t)t >> 1;
                mbar_wait(bar_acc_full + 8 * buf, use & 1u);
                tc_fence_after();
                uint32_t q[7][16];
#pragma unroll
                for (int s = 0; s < 7; ++s) {
                    if (DBG & 4) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) q[s][i] = 0u;
                    } else {
                        tmem_ld16(lane_base + buf * ACC_STRIDE + (uint32_t)(s * TN), q[s]);
                    }
                }
                tmem_wait_ld();
                tc_fence_before();                                   // all accumulator words of this warp are in registers:
                __syncwarp();                                        // hand the buffer back (the MMAs of tile t + 2 reuse it)
                if (lane == 0) mbar_arrive(bar_acc_empty + 8 * buf);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (DBG & 2) {
                        cv[i] += q[0][i] + q[6][i];
                    } else {
                        // C * 2^32 + S  with S < 2^60: bring the sum below p * 2^32 (one conditional subtraction of
                        // p * 2^32, as in mac_lazy), then REDC gives (C + S / 2^32) mod p in [0, p)
                        uint64_t acc = ((uint64_t)cv[i] << 32) + (uint64_t)q[0][i] + ((uint64_t)q[1][i] << 8) +
                                       ((uint64_t)q[2][i] << 16) + ((uint64_t)q[3][i] << 24) + (uint64_t)q[4][i] * c4 +
                                       (uint64_t)q[5][i] * c5 + (uint64_t)q[6][i] * c6;
                        uint32_t hi = (uint32_t)(acc >> 32);
                        hi = min(hi, hi - p);
                        acc = ((uint64_t)hi << 32) | (uint32_t)acc;
                        cv[i] = mont_redc(acc, p, pinv);
                    }
                }
                bool live, full;
                uint32_t* cp = tile_ptr(t, live, full);
                if (full) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        *reinterpret_cast<uint4*>(cp + 4 * i) = make_uint4(cv[4 * i], cv[4 * i + 1], cv[4 * i + 2], cv[4 * i + 3]);
                } else if (live) {
                    const int colb = g.c0 + (bst ? fixed_tile : t0 + t) * TN + half * 16;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (colb + i < g.c1) cp[i] = cv[i];
                }
            };
            uint32_t cva[16], cvb[16], cvc[16];
            load_c(0, cva);
            load_c(1, cvb);
            for (int t = 0; t < ntiles; t += 3) {
                load_c(t + 2, cvc);
                process(t, cva);
                if (t + 1 < ntiles) {
                    load_c(t + 3, cva);
                    process(t + 1, cvb);
                }
                if (t + 2 < ntiles) {
                    load_c(t + 4, cvb);
                    process(t + 2, cvc);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// the library's instantiations: epilogue with C in the coalesced mapping / in the TMEM row mapping
constexpr auto k_gemm_tc = k_gemm_tc_t<0, 1>;
constexpr auto k_gemm_tc_rowmap = k_gemm_tc_t<0, 0>;

#endif  // __CUDACC__

}  // namespace lsx_tc
