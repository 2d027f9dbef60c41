// C-ABI entry points of liblsx (declared in include/lsx.h): context, plans and the batched
// operations.  Each operation validates shapes, stages host buffers when asked to, picks a kernel
// family (fused register-resident kernels for small shapes, the shared-memory tile path otherwise)
// and enqueues everything on the ctx stream.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <thread>

#include "lsx_internal.h"

int lsx_fail(lsx_ctx* ctx, int code, const char* fmt, ...) {
    if (ctx) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        ctx->err = buf;
    }
    return code;
}

// Grow-only device workspace.  Growing synchronises the stream first (nothing in flight may
// still use the old block).
int lsx_ws_reserve(lsx_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->ws_bytes) return LSX_OK;
    LSX_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->d_ws) LSX_CUDA_TRY(ctx, cudaFree(ctx->d_ws));
    ctx->d_ws = nullptr;
    ctx->ws_bytes = 0;
    size_t want = bytes + bytes / 8 + (1u << 20);
    cudaError_t e = cudaMalloc(&ctx->d_ws, want);
    if (e != cudaSuccess) {
        want = bytes;
        e = cudaMalloc(&ctx->d_ws, want);
    }
    if (e != cudaSuccess)
        return lsx_fail(ctx, LSX_ERR_CUDA, "cudaMalloc of %zu workspace bytes failed: %s", want,
                        cudaGetErrorString(e));
    ctx->ws_bytes = want;
    return LSX_OK;
}

static int io_reserve(lsx_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->io_bytes) return LSX_OK;
    LSX_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->d_io) LSX_CUDA_TRY(ctx, cudaFree(ctx->d_io));
    ctx->d_io = nullptr;
    ctx->io_bytes = 0;
    cudaError_t e = cudaMalloc(&ctx->d_io, bytes);
    if (e != cudaSuccess)
        return lsx_fail(ctx, LSX_ERR_CUDA, "cudaMalloc of %zu staging bytes failed: %s", bytes,
                        cudaGetErrorString(e));
    ctx->io_bytes = bytes;
    return LSX_OK;
}

void lsx_timing_begin(lsx_ctx* ctx) {
    if (!ctx->timing) return;
    while ((int)ctx->tev.size() < 2 * (ctx->tev_used + 1)) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) {
            ctx->timing = false;
            return;
        }
        ctx->tev.push_back(e);
    }
    cudaEventRecord(ctx->tev[2 * ctx->tev_used], ctx->stream);
}
void lsx_timing_end(lsx_ctx* ctx) {
    if (!ctx->timing) return;
    cudaEventRecord(ctx->tev[2 * ctx->tev_used + 1], ctx->stream);
    ctx->tev_used++;
}

// ---- prime table upload -------------------------------------------------------------------------
static int upload_tables(lsx_ctx* ctx) {
    std::vector<PrimeRec> recs(LSX_TABLE_PRIMES);
    for (int i = 0; i < LSX_TABLE_PRIMES; ++i) recs[i] = lsx_make_prime_rec(ctx->primes[i]);
    std::vector<uint32_t> garner((size_t)LSX_GARNER_DIM * LSX_GARNER_DIM, 0u);
    for (int i = 0; i < LSX_GARNER_DIM; ++i) {
        for (int j = 0; j < LSX_GARNER_DIM; ++j) {
            if (i == j) continue;
            const uint32_t pi = ctx->primes[i], pj = ctx->primes[j];
            if (pi % pj == 0) continue;
            const uint32_t inv = lsx_inv_mod(pi % pj, pj);                  // p_i^{-1} mod p_j
            garner[(size_t)i * LSX_GARNER_DIM + j] = (uint32_t)(((uint64_t)inv << 32) % pj);   // * R
        }
    }
    LSX_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    LSX_CUDA_TRY(ctx, cudaMemcpy(ctx->d_primes, recs.data(), recs.size() * sizeof(PrimeRec), cudaMemcpyHostToDevice));
    LSX_CUDA_TRY(ctx, cudaMemcpy(ctx->d_garner, garner.data(), garner.size() * 4, cudaMemcpyHostToDevice));
    return LSX_OK;
}

extern "C" {

int lsx_abi_version(void) { return LSX_ABI_VERSION; }

int lsx_create(int device_id, lsx_ctx** out) {
    if (!out) return LSX_ERR_NULL;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return LSX_ERR_NO_DEVICE;
    if (device_id < 0 || device_id >= ndev) return LSX_ERR_NO_DEVICE;
    lsx_ctx* ctx = new (std::nothrow) lsx_ctx();
    if (!ctx) return LSX_ERR_CUDA;
    ctx->device = device_id;
    if (cudaSetDevice(device_id) != cudaSuccess) {
        delete ctx;
        return LSX_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < 2; ++i)
        ok = cudaStreamCreateWithFlags(&ctx->copy_streams[i], cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < 8; ++i)
        ok = cudaEventCreateWithFlags(&ctx->events[i], cudaEventDisableTiming) == cudaSuccess;
    ctx->stream = ctx->own_stream;
    ok = ok && cudaMalloc(&ctx->d_primes, LSX_TABLE_PRIMES * sizeof(PrimeRec)) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->d_garner, (size_t)LSX_GARNER_DIM * LSX_GARNER_DIM * 4) == cudaSuccess;
    if (ok) {
        lsx_fill_prime_table(ctx->primes, LSX_TABLE_PRIMES);
        ok = upload_tables(ctx) == LSX_OK;
    }
    if (!ok) {
        lsx_destroy(ctx);
        return LSX_ERR_CUDA;
    }
    *out = ctx;
    return LSX_OK;
}

void lsx_destroy(lsx_ctx* ctx) {
    if (!ctx) return;
    lsx_multi_release(ctx);
    for (lsx_ctx* peer : ctx->peers) lsx_destroy(peer);
    ctx->peers.clear();
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
    if (ctx->d_ws) cudaFree(ctx->d_ws);
    if (ctx->d_io) cudaFree(ctx->d_io);
    if (ctx->d_primes) cudaFree(ctx->d_primes);
    if (ctx->d_garner) cudaFree(ctx->d_garner);
    for (auto& e : ctx->events)
        if (e) cudaEventDestroy(e);
    for (auto& e : ctx->tev)
        if (e) cudaEventDestroy(e);
    for (auto& e : ctx->pev)
        if (e) cudaEventDestroy(e);
    for (auto& e : ctx->side_events)
        if (e) cudaEventDestroy(e);
    if (ctx->side_fork) cudaEventDestroy(ctx->side_fork);
    for (auto& s : ctx->side_streams)
        if (s) cudaStreamDestroy(s);
    for (auto& s : ctx->copy_streams)
        if (s) cudaStreamDestroy(s);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* lsx_last_error(const lsx_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int lsx_set_stream(lsx_ctx* ctx, void* cuda_stream) {
    if (!ctx) return LSX_ERR_NULL;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return LSX_OK;
}

int lsx_synchronize(lsx_ctx* ctx) {
    if (!ctx) return LSX_ERR_NULL;
    LSX_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    LSX_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LSX_OK;
}

int64_t lsx_launch_count(const lsx_ctx* ctx) { return ctx ? ctx->launches : 0; }

int lsx_last_prime_count(lsx_ctx* ctx, int* out) {
    if (!ctx || !out) return LSX_ERR_NULL;
    *out = 0;
    if (!ctx->last_kword) return LSX_OK;
    LSX_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int32_t kw[2] = {0, 0};
    LSX_CUDA_TRY(ctx, cudaMemcpyAsync(kw, ctx->last_kword, 8, cudaMemcpyDeviceToHost, ctx->stream));
    LSX_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *out = kw[1];
    return LSX_OK;
}

int lsx_timing_enable(lsx_ctx* ctx, int enable) {
    if (!ctx) return LSX_ERR_NULL;
    ctx->timing = enable != 0;
    ctx->tev_used = 0;
    return LSX_OK;
}

int lsx_timing_read(lsx_ctx* ctx, float* ms_out, int cap, int* count) {
    if (!ctx || !count) return LSX_ERR_NULL;
    LSX_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int n = ctx->tev_used < cap ? ctx->tev_used : cap;
    for (int i = 0; i < n; ++i) {
        LSX_CUDA_TRY(ctx, cudaEventSynchronize(ctx->tev[2 * i + 1]));
        float ms = 0.f;
        LSX_CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->tev[2 * i], ctx->tev[2 * i + 1]));
        if (ms_out) ms_out[i] = ms;
    }
    *count = ctx->tev_used;
    ctx->tev_used = 0;
    return LSX_OK;
}

int lsx_debug_set_primes(lsx_ctx* ctx, const uint32_t* primes, int count) {
    if (!ctx) return LSX_ERR_NULL;
    if (count < 0 || count > LSX_TABLE_PRIMES) return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "bad prime count %d", count);
    LSX_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    lsx_fill_prime_table(ctx->primes, LSX_TABLE_PRIMES);
    if (count > 0) {
        if (!primes) return LSX_ERR_NULL;
        // the replaced primes must be odd primes < 2^31, distinct from each other; table primes
        // that collide with them are dropped from the tail
        std::vector<uint32_t> t;
        for (int i = 0; i < count; ++i) {
            uint32_t p = primes[i];
            if (p < 3 || p >= 0x80000000u || !lsx_is_prime_u32(p) || std::find(t.begin(), t.end(), p) != t.end())
                return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "prime %u is not usable", p);
            t.push_back(p);
        }
        std::vector<uint32_t> full;
        lsx_fill_prime_table(full, LSX_TABLE_PRIMES + count);
        for (uint32_t p : full) {
            if ((int)t.size() >= LSX_TABLE_PRIMES) break;
            if (std::find(t.begin(), t.begin() + count, p) == t.begin() + count) t.push_back(p);
        }
        ctx->primes = t;
    }
    return upload_tables(ctx);
}

int lsx_get_primes(const lsx_ctx* ctx, uint32_t* out, int count) {
    if (!ctx || !out) return LSX_ERR_NULL;
    if (count < 0 || count > LSX_TABLE_PRIMES) return LSX_ERR_BAD_SHAPE;
    memcpy(out, ctx->primes.data(), (size_t)count * 4);
    return LSX_OK;
}

// ---- plans ----------------------------------------------------------------------------------------
static int fill_plan(lsx_plan* out, int op, int m, int n, int bar, int max_rank, int gen_cap, int64_t a, int64_t b,
                     bool has_right, bool right_identity) {
    if (!out) return LSX_ERR_NULL;
    memset(out, 0, sizeof *out);
    if (m < 1 || n < 1 || bar < 1 || bar > n || a < 0 || b < 0 || a > 0x7fffffffLL || b > 0x7fffffffLL)
        return LSX_ERR_BAD_SHAPE;
    if (m > 254) return LSX_ERR_UNSUPPORTED;
    const int slots = m < bar ? m : bar;
    if (max_rank <= 0 || max_rank > slots) max_rank = slots;
    out->op = op;
    out->m = m;
    out->n = n;
    out->bar_col = bar;
    out->max_rank = max_rank;
    out->pivot_slots = slots;
    out->gen_cap = gen_cap;
    out->a_abs_max = a;
    out->b_abs_max = b;
    out->log2_bound = lsx_log2_minor_bound(m, bar, has_right, a, b, right_identity, max_rank);
    lsx_bits_to_plan(out->log2_bound, &out->n_primes, &out->limbs);
    if (out->n_primes > LSX_MAX_BATCH_PRIMES) return LSX_ERR_BOUND;
    return LSX_OK;
}

int lsx_plan_rref(int m, int n, int bar_col, int64_t a_abs_max, int64_t b_abs_max, int max_rank, lsx_plan* out) {
    return fill_plan(out, LSX_OP_RREF, m, n, bar_col, max_rank, 0, a_abs_max, b_abs_max, bar_col < n, false);
}
int lsx_plan_inverse(int n, int64_t a_abs_max, lsx_plan* out) {
    if (n < 1) return LSX_ERR_BAD_SHAPE;
    return fill_plan(out, LSX_OP_INVERSE, n, 2 * n, n, n, 0, a_abs_max, 1, true, true);
}
int lsx_plan_det(int n, int64_t a_abs_max, lsx_plan* out) {
    if (n < 1) return LSX_ERR_BAD_SHAPE;
    return fill_plan(out, LSX_OP_DET, n, n, n, n, 0, a_abs_max, 0, false, false);
}
int lsx_plan_rank(int m, int n, int64_t a_abs_max, lsx_plan* out) {
    return fill_plan(out, LSX_OP_RANK, m, n, n, 0, 0, a_abs_max, 0, false, false);
}
int lsx_plan_solve(int m, int n, int64_t a_abs_max, int64_t b_abs_max, int max_rank, int gen_cap, lsx_plan* out) {
    if (n < 1 || gen_cap < 0 || gen_cap > n) return LSX_ERR_BAD_SHAPE;
    return fill_plan(out, LSX_OP_SOLVE, m, n + 1, n, max_rank, gen_cap, a_abs_max, b_abs_max, true, false);
}

}  // extern "C"

// ---- running a job: staging, chunking, kernel-family choice --------------------------------------
namespace {

struct Buf {          // one caller buffer: where it lives on the host side and its size per matrix
    const void* src;  // input: caller pointer (host or device)
    void* dst;        // output: caller pointer
    size_t per;       // bytes per matrix
    void** slot;      // the ElimJob field that receives the device pointer
    bool input;
};

size_t up256(size_t x) { return (x + 255) / 256 * 256; }

// Sub-job for matrices [b0, b0 + cnt) of `job` (all pointers are device pointers).
ElimJob slice_job(const ElimJob& job, int64_t b0, int64_t cnt) {
    const int m = job.m, L = job.L;
    const int slots = m < job.bar ? m : job.bar;
    const int nvars = job.n - 1;
    ElimJob c = job;
    c.batch = cnt;
    c.A = job.in_i8 ? (const int32_t*)((const int8_t*)job.A + b0 * m * job.n_in) : job.A + b0 * m * job.n_in;
    if (job.bvec) c.bvec = job.bvec + b0 * m;
    if (job.num) {
        const int64_t per = job.op == LSX_OP_INVERSE ? (int64_t)m * m : (int64_t)m * job.n;
        c.num = job.num + b0 * per * L;
    }
    if (job.den) c.den = job.den + b0 * L;
    if (job.particular) c.particular = job.particular + b0 * nvars * L;
    if (job.generators) c.generators = job.generators + b0 * nvars * job.gen_cap * L;
    if (job.pivot_col) c.pivot_col = job.pivot_col + b0 * slots;
    if (job.rank) c.rank = job.rank + b0;
    c.status = job.status + b0;
    return c;
}

// clear_status: the per-matrix status words still have to be zeroed on the stream.  The fused small-matrix kernel
// writes every status word itself, so the memset (a second launch per call: 5 % of a 2^20 x 8x8 step) is skipped for it.
int run_chunks(lsx_ctx* ctx, const ElimJob& job, bool clear_status) {
    // The tile path keeps K residue planes per matrix in scratch; bound the scratch by chunking.
    static const size_t budget = []() {
        const char* e = getenv("LSX_WS_MB");
        size_t mb = e ? (size_t)strtoull(e, nullptr, 10) : 8192;   // of 180 GB; 2 GiB cost 2-3 % on 2^16 64x64 tiles (6.5 x the launches)
        return (mb < 16 ? 16 : mb) << 20;
    }();
    int handled = 0;
    int rc = lsx_run_small(ctx, job, &handled);
    if (rc != LSX_OK) return rc;
    if (handled) return LSX_OK;
    if (clear_status) LSX_CUDA_TRY(ctx, cudaMemsetAsync(job.status, 0, (size_t)job.batch * 4, ctx->stream));
    if (job.in_i8) return lsx_fail(ctx, LSX_ERR_UNSUPPORTED, "int8 input is served by the fused small-matrix inverse only");
    rc = lsx_run_subwarp(ctx, job, &handled);
    if (rc != LSX_OK) return rc;
    if (handled) return LSX_OK;

    const size_t per1 = lsx_generic_ws_bytes(job, 1), per2 = lsx_generic_ws_bytes(job, 2);
    const size_t slope = per2 > per1 ? per2 - per1 : 1;
    int64_t chunk = (int64_t)((budget > per1 ? budget - per1 : 0) / slope) + 1;
    if (chunk > job.batch) chunk = job.batch;
    if (chunk < 1) chunk = 1;
    rc = lsx_ws_reserve(ctx, lsx_generic_ws_bytes(job, chunk) + 4096);
    if (rc != LSX_OK) return rc;
    for (int64_t b0 = 0; b0 < job.batch; b0 += chunk) {
        rc = lsx_run_generic(ctx, slice_job(job, b0, std::min(chunk, job.batch - b0)), nullptr, nullptr, 0, 0);
        if (rc != LSX_OK) return rc;
    }
    return LSX_OK;
}

cudaEvent_t pipe_event(lsx_ctx* ctx, size_t i) {
    while (ctx->pev.size() <= i) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        ctx->pev.push_back(e);
    }
    return ctx->pev[i];
}

// Stage (mem == HOST) or pass through (mem == DEVICE) the buffers, clear status, run, copy back.
// Host calls are pipelined in slices: copy-in, kernels and copy-out of different slices overlap on
// three streams (the two copy directions use separate DMA engines), so a call with pinned host
// buffers runs at the speed of the slower PCIe direction rather than the sum of all three phases.
int run_job(lsx_ctx* ctx, ElimJob& job, int mem, Buf* bufs, int nbufs, int32_t* status_user) {
    LSX_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (mem != LSX_MEM_HOST && mem != LSX_MEM_DEVICE) return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "bad mem flag %d", mem);
    if (job.batch < 0) return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "negative batch");
    if (job.batch == 0) return LSX_OK;
    if (job.batch > 0x7fffffffLL) return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "batch above 2^31-1");
    if (!ctx->peers.empty()) {
        // a context over several GPUs (lsx_create_multi): host buffers are sharded BY MATRIX, one host thread per GPU
        // running this very function on its contiguous slice; matrices are independent, nothing is exchanged
        if (mem != LSX_MEM_HOST)
            return lsx_fail(ctx, LSX_ERR_UNSUPPORTED, "a multi-GPU context shards host buffers; device buffers belong to "
                                                      "one GPU: use one context per GPU for device-resident data");
        std::vector<lsx_ctx*> dev{ctx};
        dev.insert(dev.end(), ctx->peers.begin(), ctx->peers.end());
        const int world = (int)std::min<int64_t>((int64_t)dev.size(), job.batch);
        std::vector<int> rcs(world, LSX_OK);
        std::vector<std::thread> th;
        std::vector<std::vector<lsx_ctx*>> saved(world);
        for (int d = 0; d < world; ++d) {
            const int64_t base = job.batch / world, extra = job.batch % world;
            const int64_t b0 = d * base + std::min<int64_t>(d, extra), cnt = base + (d < extra ? 1 : 0);
            th.emplace_back([&, d, b0, cnt]() {
                lsx_ctx* c = dev[d];
                ElimJob jd = job;
                jd.batch = cnt;
                std::vector<Buf> bd(bufs, bufs + nbufs);
                for (Buf& b : bd) {
                    if (b.src) b.src = (const char*)b.src + b.per * (size_t)b0;
                    if (b.dst) b.dst = (char*)b.dst + b.per * (size_t)b0;
                    if (b.slot) b.slot = (void**)((char*)&jd + ((char*)b.slot - (char*)&job));   // same field of the copy
                }
                saved[d].swap(c->peers);                           // the slice runs as a plain single-device call
                rcs[d] = run_job(c, jd, mem, bd.data(), nbufs, status_user + b0);
                saved[d].swap(c->peers);
            });
        }
        for (auto& t : th) t.join();
        for (int d = 0; d < world; ++d)
            if (rcs[d] != LSX_OK) {
                if (d) ctx->err = dev[d]->err;
                return rcs[d];
            }
        return LSX_OK;
    }
    const size_t st_bytes = (size_t)job.batch * 4;
    if (mem == LSX_MEM_DEVICE) {
        for (int i = 0; i < nbufs; ++i)
            if (bufs[i].slot) *bufs[i].slot = bufs[i].input ? const_cast<void*>(bufs[i].src) : bufs[i].dst;
        job.status = status_user;
        return run_chunks(ctx, job, true);
    }
    size_t total = up256(st_bytes), per_matrix = 4;
    for (int i = 0; i < nbufs; ++i)
        if (bufs[i].slot && (bufs[i].input ? bufs[i].src != nullptr : bufs[i].dst != nullptr)) {
            total += up256(bufs[i].per * (size_t)job.batch);
            per_matrix += bufs[i].per;
        }
    int rc = io_reserve(ctx, total);
    if (rc != LSX_OK) return rc;
    char* base = (char*)ctx->d_io;
    size_t off = 0;
    job.status = (int32_t*)base;
    off += up256(st_bytes);
    std::vector<void*> dev(nbufs, nullptr);
    for (int i = 0; i < nbufs; ++i) {
        const bool present = bufs[i].slot && (bufs[i].input ? bufs[i].src != nullptr : bufs[i].dst != nullptr);
        if (!present) continue;
        dev[i] = base + off;
        off += up256(bufs[i].per * (size_t)job.batch);
        *bufs[i].slot = dev[i];
    }
    // slices of about 32 MiB of traffic, at most 64 of them
    int64_t slice = (int64_t)((32u << 20) / per_matrix);
    if (slice < 1) slice = 1;
    if ((job.batch + slice - 1) / slice > 64) slice = (job.batch + 63) / 64;
    cudaStream_t s_in = ctx->copy_streams[0], s_out = ctx->copy_streams[1], s_run = ctx->stream;
    // the staging block may still be read by an earlier call's copies on this ctx: they were synchronised
    // before that call returned, so only ordering against s_run matters here
    size_t ev = 0;
    for (int64_t b0 = 0; b0 < job.batch; b0 += slice) {
        const int64_t cnt = std::min(slice, job.batch - b0);
        for (int i = 0; i < nbufs; ++i)
            if (dev[i] && bufs[i].input)
                LSX_CUDA_TRY(ctx, cudaMemcpyAsync((char*)dev[i] + bufs[i].per * (size_t)b0,
                                                  (const char*)bufs[i].src + bufs[i].per * (size_t)b0,
                                                  bufs[i].per * (size_t)cnt, cudaMemcpyHostToDevice, s_in));
        cudaEvent_t e_in = pipe_event(ctx, ev++), e_done = pipe_event(ctx, ev++);
        if (!e_in || !e_done) return lsx_fail(ctx, LSX_ERR_CUDA, "cudaEventCreate failed");
        LSX_CUDA_TRY(ctx, cudaEventRecord(e_in, s_in));
        LSX_CUDA_TRY(ctx, cudaStreamWaitEvent(s_run, e_in, 0));
        rc = run_chunks(ctx, slice_job(job, b0, cnt), true);
        if (rc != LSX_OK) {
            cudaStreamSynchronize(s_in);
            cudaStreamSynchronize(s_run);
            cudaStreamSynchronize(s_out);
            return rc;
        }
        LSX_CUDA_TRY(ctx, cudaEventRecord(e_done, s_run));
        LSX_CUDA_TRY(ctx, cudaStreamWaitEvent(s_out, e_done, 0));
        for (int i = 0; i < nbufs; ++i)
            if (dev[i] && !bufs[i].input)
                LSX_CUDA_TRY(ctx, cudaMemcpyAsync((char*)bufs[i].dst + bufs[i].per * (size_t)b0,
                                                  (const char*)dev[i] + bufs[i].per * (size_t)b0,
                                                  bufs[i].per * (size_t)cnt, cudaMemcpyDeviceToHost, s_out));
        LSX_CUDA_TRY(ctx, cudaMemcpyAsync(status_user + b0, job.status + b0, (size_t)cnt * 4, cudaMemcpyDeviceToHost,
                                          s_out));
    }
    LSX_CUDA_TRY(ctx, cudaStreamSynchronize(s_in));
    LSX_CUDA_TRY(ctx, cudaStreamSynchronize(s_run));
    LSX_CUDA_TRY(ctx, cudaStreamSynchronize(s_out));
    return LSX_OK;
}

int check_plan(lsx_ctx* ctx, const lsx_plan* plan, int op) {
    if (!ctx) return LSX_ERR_NULL;
    if (!plan) return lsx_fail(ctx, LSX_ERR_NULL, "plan is NULL");
    if (plan->op != op) return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "plan is for operation %d, call needs %d", plan->op, op);
    if (plan->m < 1 || plan->n < 1 || plan->bar_col < 1 || plan->bar_col > plan->n || plan->m > 254)
        return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "bad plan shape %dx%d bar %d", plan->m, plan->n, plan->bar_col);
    if (plan->n_primes < 1 || plan->n_primes > LSX_MAX_BATCH_PRIMES || plan->limbs < 1 ||
        plan->limbs > LSX_MAX_BATCH_PRIMES)
        return lsx_fail(ctx, LSX_ERR_BOUND, "plan needs %d primes / %d limbs", plan->n_primes, plan->limbs);
    return LSX_OK;
}

ElimJob job_from_plan(const lsx_plan* plan, int64_t batch) {
    ElimJob j;
    j.batch = batch;
    j.m = plan->m;
    j.n = plan->n;
    j.n_in = plan->n;
    j.bar = plan->bar_col;
    j.a_abs_max = plan->a_abs_max;
    j.b_abs_max = plan->b_abs_max;
    j.max_rank = plan->max_rank;
    j.op = plan->op;
    j.K = plan->n_primes;
    j.L = plan->limbs;
    j.gen_cap = plan->gen_cap;
    return j;
}

}  // namespace

extern "C" {

int lsx_rref_batch(lsx_ctx* ctx, const lsx_plan* plan, const int32_t* A, int64_t batch, int mem, uint32_t* num,
                   uint32_t* den, int32_t* pivot_col, int32_t* rank, int32_t* status) {
    int rc = check_plan(ctx, plan, LSX_OP_RREF);
    if (rc != LSX_OK) return rc;
    if (batch > 0 && (!A || !num || !den || !status)) return lsx_fail(ctx, LSX_ERR_NULL, "rref: NULL buffer");
    ElimJob j = job_from_plan(plan, batch);
    const size_t L4 = (size_t)j.L * 4;
    Buf bufs[] = {
        {A, nullptr, (size_t)j.m * j.n * 4, (void**)&j.A, true},
        {nullptr, num, (size_t)j.m * j.n * L4, (void**)&j.num, false},
        {nullptr, den, L4, (void**)&j.den, false},
        {nullptr, pivot_col, (size_t)plan->pivot_slots * 4, (void**)&j.pivot_col, false},
        {nullptr, rank, 4, (void**)&j.rank, false},
    };
    return run_job(ctx, j, mem, bufs, 5, status);
}

int lsx_inverse_batch(lsx_ctx* ctx, const lsx_plan* plan, const int32_t* A, int64_t batch, int mem, uint32_t* adj,
                      uint32_t* det, int32_t* status) {
    int rc = check_plan(ctx, plan, LSX_OP_INVERSE);
    if (rc != LSX_OK) return rc;
    if (plan->n != 2 * plan->m || plan->bar_col != plan->m)
        return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "inverse plan must be n x 2n with bar n");
    if (batch > 0 && (!A || !adj || !det || !status)) return lsx_fail(ctx, LSX_ERR_NULL, "inverse: NULL buffer");
    ElimJob j = job_from_plan(plan, batch);
    j.n_in = j.m;
    j.right_identity = 1;
    const size_t L4 = (size_t)j.L * 4;
    Buf bufs[] = {
        {A, nullptr, (size_t)j.m * j.m * 4, (void**)&j.A, true},
        {nullptr, adj, (size_t)j.m * j.m * L4, (void**)&j.num, false},
        {nullptr, det, L4, (void**)&j.den, false},
    };
    return run_job(ctx, j, mem, bufs, 3, status);
}

int lsx_inverse_batch_i8(lsx_ctx* ctx, const lsx_plan* plan, const int8_t* A, int64_t batch, int mem, uint32_t* adj,
                         uint32_t* det, int32_t* status) {
    int rc = check_plan(ctx, plan, LSX_OP_INVERSE);
    if (rc != LSX_OK) return rc;
    if (plan->n != 2 * plan->m || plan->bar_col != plan->m)
        return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "inverse plan must be n x 2n with bar n");
    if (batch > 0 && (!A || !adj || !det || !status)) return lsx_fail(ctx, LSX_ERR_NULL, "inverse: NULL buffer");
    ElimJob j = job_from_plan(plan, batch);
    j.n_in = j.m;
    j.right_identity = 1;
    j.in_i8 = 1;
    const size_t L4 = (size_t)j.L * 4;
    Buf bufs[] = {
        {A, nullptr, (size_t)j.m * j.m, (void**)&j.A, true},
        {nullptr, adj, (size_t)j.m * j.m * L4, (void**)&j.num, false},
        {nullptr, det, L4, (void**)&j.den, false},
    };
    return run_job(ctx, j, mem, bufs, 3, status);
}

int lsx_det_batch(lsx_ctx* ctx, const lsx_plan* plan, const int32_t* A, int64_t batch, int mem, uint32_t* det,
                  int32_t* rank, int32_t* status) {
    int rc = check_plan(ctx, plan, LSX_OP_DET);
    if (rc != LSX_OK) return rc;
    if (plan->n != plan->m) return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "determinant needs a square matrix");
    if (batch > 0 && (!A || !det || !status)) return lsx_fail(ctx, LSX_ERR_NULL, "det: NULL buffer");
    ElimJob j = job_from_plan(plan, batch);
    const size_t L4 = (size_t)j.L * 4;
    Buf bufs[] = {
        {A, nullptr, (size_t)j.m * j.n * 4, (void**)&j.A, true},
        {nullptr, det, L4, (void**)&j.den, false},
        {nullptr, rank, 4, (void**)&j.rank, false},
    };
    return run_job(ctx, j, mem, bufs, 3, status);
}

int lsx_rank_batch(lsx_ctx* ctx, const lsx_plan* plan, const int32_t* A, int64_t batch, int mem, int32_t* rank,
                   int32_t* status) {
    int rc = check_plan(ctx, plan, LSX_OP_RANK);
    if (rc != LSX_OK) return rc;
    if (batch > 0 && (!A || !rank || !status)) return lsx_fail(ctx, LSX_ERR_NULL, "rank: NULL buffer");
    ElimJob j = job_from_plan(plan, batch);
    Buf bufs[] = {
        {A, nullptr, (size_t)j.m * j.n * 4, (void**)&j.A, true},
        {nullptr, rank, 4, (void**)&j.rank, false},
    };
    return run_job(ctx, j, mem, bufs, 2, status);
}

int lsx_solve_batch(lsx_ctx* ctx, const lsx_plan* plan, const int32_t* A, const int32_t* b, int64_t batch, int mem,
                    uint32_t* den, uint32_t* particular, uint32_t* generators, int32_t* pivot_col, int32_t* rank,
                    int32_t* status) {
    int rc = check_plan(ctx, plan, LSX_OP_SOLVE);
    if (rc != LSX_OK) return rc;
    if (plan->bar_col != plan->n - 1) return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "solve plan must have bar = n - 1");
    if (batch > 0 && (!A || !b || !den || !particular || !status || (plan->gen_cap > 0 && !generators)))
        return lsx_fail(ctx, LSX_ERR_NULL, "solve: NULL buffer");
    ElimJob j = job_from_plan(plan, batch);
    j.n_in = j.n - 1;
    const int nvars = j.n - 1;
    const size_t L4 = (size_t)j.L * 4;
    Buf bufs[] = {
        {A, nullptr, (size_t)j.m * nvars * 4, (void**)&j.A, true},
        {b, nullptr, (size_t)j.m * 4, (void**)&j.bvec, true},
        {nullptr, den, L4, (void**)&j.den, false},
        {nullptr, particular, (size_t)nvars * L4, (void**)&j.particular, false},
        {nullptr, generators, (size_t)nvars * j.gen_cap * L4, (void**)&j.generators, false},
        {nullptr, pivot_col, (size_t)plan->pivot_slots * 4, (void**)&j.pivot_col, false},
        {nullptr, rank, 4, (void**)&j.rank, false},
    };
    return run_job(ctx, j, mem, bufs, 7, status);
}

}  // extern "C"
