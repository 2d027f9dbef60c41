// Several GPUs behind ONE context of the C-ABI: a caller that links liblsx.so directly --
// the Rust crate is the stated second binding -- gets the box's GPUs without torch.distributed.
//
//   * lsx_create_multi builds one single-device ctx per GPU; the first is the handle, the others hang off it.
//   * Batched calls with host buffers shard BY MATRIX: contiguous slices, one host thread per GPU running the
//     ordinary single-device pipeline on its slice (lsx_api.cu: run_job).  Matrices are independent: no collective.
//   * lsx_det_large shards BY PRIME: every GPU gets its own copy of A and a contiguous range of table primes, the
//     residue vectors are ALL-GATHERED over NVLink by NCCL (one communicator per GPU, created on first use;
//     libnccl is loaded at run time so that liblsx.so has no link-time dependency on it; if it cannot be loaded
//     the gather falls back to peer copies) and the Garner CRT runs on the first GPU.
#include <dlfcn.h>

#include <algorithm>
#include <thread>

#include "lsx_internal.h"

namespace {

// ---- the few NCCL entry points used, resolved with dlsym ------------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
constexpr int kNcclUint32 = 3;            // ncclUint32 in nccl.h (ncclInt8 0, ncclUint8 1, ncclInt32 2, ncclUint32 3)

struct NcclApi {
    void* so = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

struct MultiState {
    NcclApi api;
    std::vector<ncclComm_t> comms;        // one per device, in device-list order
    bool tried = false;
};

void load_nccl(NcclApi& a) {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        a.so = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (a.so) break;
    }
    if (!a.so) return;
    a.CommInitAll = (decltype(a.CommInitAll))dlsym(a.so, "ncclCommInitAll");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.so, "ncclCommDestroy");
    a.GroupStart = (decltype(a.GroupStart))dlsym(a.so, "ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.so, "ncclGroupEnd");
    a.AllGather = (decltype(a.AllGather))dlsym(a.so, "ncclAllGather");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.so, "ncclGetErrorString");
    a.ok = a.CommInitAll && a.CommDestroy && a.GroupStart && a.GroupEnd && a.AllGather;
}

std::vector<lsx_ctx*> devices_of(lsx_ctx* ctx) {
    std::vector<lsx_ctx*> d{ctx};
    d.insert(d.end(), ctx->peers.begin(), ctx->peers.end());
    return d;
}

void shard(int total, int r, int world, int* b, int* e) {
    const int base = total / world, extra = total % world;
    *b = r * base + std::min(r, extra);
    *e = *b + base + (r < extra ? 1 : 0);
}

// residues of every device, padded to `width` words per device, -> contiguous [total] on the first device
__global__ void k_compact(const uint32_t* __restrict__ gathered, int width, int world, int total, uint32_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int base = total / world, extra = total % world;
    // inverse of shard(): rank r owns [r * base + min(r, extra), ...)
    int r = 0, b = 0;
    for (; r < world; ++r) {
        const int sz = base + (r < extra ? 1 : 0);
        if (i < b + sz) break;
        b += sz;
    }
    out[i] = gathered[(size_t)r * width + (i - b)];
}

}  // namespace

void lsx_multi_release(lsx_ctx* ctx) {
    MultiState* st = (MultiState*)ctx->nccl;
    if (!st) return;
    for (ncclComm_t c : st->comms)
        if (c && st->api.CommDestroy) st->api.CommDestroy(c);
    delete st;
    ctx->nccl = nullptr;
}

extern "C" {

int lsx_create_multi(const int* device_ids, int n_dev, lsx_ctx** out) {
    if (!out || !device_ids) return LSX_ERR_NULL;
    *out = nullptr;
    if (n_dev < 1 || n_dev > 64) return LSX_ERR_BAD_SHAPE;
    for (int i = 0; i < n_dev; ++i)
        for (int j = 0; j < i; ++j)
            if (device_ids[i] == device_ids[j]) return LSX_ERR_BAD_SHAPE;
    lsx_ctx* head = nullptr;
    int rc = lsx_create(device_ids[0], &head);
    if (rc != LSX_OK) return rc;
    for (int i = 1; i < n_dev; ++i) {
        lsx_ctx* c = nullptr;
        rc = lsx_create(device_ids[i], &c);
        if (rc != LSX_OK) {
            lsx_destroy(head);
            return rc;
        }
        head->peers.push_back(c);
    }
    *out = head;
    return LSX_OK;
}

int lsx_device_count(const lsx_ctx* ctx) { return ctx ? 1 + (int)ctx->peers.size() : 0; }

int lsx_multi_uses_nccl(const lsx_ctx* ctx) {
    const MultiState* st = ctx ? (const MultiState*)ctx->nccl : nullptr;
    return st && st->api.ok && !st->comms.empty() ? 1 : 0;
}

int lsx_det_large(lsx_ctx* ctx, const int32_t* A, int n, int limbs_cap, uint32_t* det_words, int* limbs_out,
                  int* n_primes_out) {
    if (!ctx) return LSX_ERR_NULL;
    if (!A || !det_words) return lsx_fail(ctx, LSX_ERR_NULL, "det_large: NULL buffer");
    if (n < 1 || limbs_cap < 1) return lsx_fail(ctx, LSX_ERR_BAD_SHAPE, "det_large: bad n or limbs_cap");
    int K = 0;
    double bits = 0.0;
    int rc = lsx_det_large_prime_count_for(ctx, A, n, LSX_MEM_HOST, &K, &bits);
    if (rc != LSX_OK) return rc;
    const int limbs = (int)(bits + 2) / 32 + 1;
    if (limbs_out) *limbs_out = limbs;
    if (n_primes_out) *n_primes_out = K;
    if (limbs > limbs_cap) return lsx_fail(ctx, LSX_ERR_BOUND, "det_large: the determinant needs %d limbs, room for %d", limbs, limbs_cap);
    std::vector<lsx_ctx*> dev = devices_of(ctx);
    const int world = (int)std::min<size_t>(dev.size(), (size_t)K);
    const int width = (K + world - 1) / world;
    const size_t abytes = (size_t)n * n * 4;
    struct PerDev {
        int32_t* dA = nullptr;
        uint32_t* local = nullptr;     // [width]
        uint32_t* all = nullptr;       // [world][width]
        int rc = LSX_OK;
    };
    std::vector<PerDev> pd(world);
    auto cleanup = [&]() {
        for (int d = 0; d < world; ++d) {
            cudaSetDevice(dev[d]->device);
            if (pd[d].dA) cudaFree(pd[d].dA);
            if (pd[d].local) cudaFree(pd[d].local);
            if (pd[d].all) cudaFree(pd[d].all);
        }
    };
    // ---- residues: one host thread per GPU, its own copy of A, its own prime range ----
    {
        std::vector<std::thread> th;
        for (int d = 0; d < world; ++d)
            th.emplace_back([&, d]() {
                lsx_ctx* c = dev[d];
                PerDev& P = pd[d];
                int b, e;
                shard(K, d, world, &b, &e);
                if (cudaSetDevice(c->device) != cudaSuccess || cudaMalloc(&P.dA, abytes) != cudaSuccess ||
                    cudaMalloc(&P.local, (size_t)width * 4) != cudaSuccess ||
                    cudaMalloc(&P.all, (size_t)world * width * 4) != cudaSuccess) {
                    P.rc = lsx_fail(c, LSX_ERR_CUDA, "det_large: device allocation failed");
                    return;
                }
                cudaMemsetAsync(P.local, 0, (size_t)width * 4, c->stream);
                if (cudaMemcpyAsync(P.dA, A, abytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) {
                    P.rc = lsx_fail(c, LSX_ERR_CUDA, "det_large: H2D copy failed");
                    return;
                }
                P.rc = lsx_det_large_residues(c, P.dA, n, b, e - b, LSX_MEM_DEVICE, P.local, nullptr);
                if (P.rc == LSX_OK && cudaStreamSynchronize(c->stream) != cudaSuccess)
                    P.rc = lsx_fail(c, LSX_ERR_CUDA, "det_large: residue kernels failed");
            });
        for (auto& t : th) t.join();
    }
    for (int d = 0; d < world; ++d)
        if (pd[d].rc != LSX_OK) {
            if (d) ctx->err = dev[d]->err;
            cleanup();
            return pd[d].rc;
        }
    // ---- all-gather of the residues (NCCL over NVLink; peer copies if libnccl is unavailable) ----
    const uint32_t* gathered = pd[0].local;                       // world == 1: nothing to gather
    if (world > 1) {
        MultiState* st = (MultiState*)ctx->nccl;
        if (!st) ctx->nccl = st = new MultiState();
        if (!st->tried) {
            st->tried = true;
            load_nccl(st->api);
            if (st->api.ok) {
                std::vector<int> ids;
                for (lsx_ctx* c : devices_of(ctx)) ids.push_back(c->device);
                st->comms.assign(ids.size(), nullptr);
                if (st->api.CommInitAll(st->comms.data(), (int)ids.size(), ids.data()) != 0) {
                    st->comms.clear();
                    st->api.ok = false;
                }
            }
        }
        bool done = false;
        if (st->api.ok && world == (int)st->comms.size()) {
            ncclResult_t r = st->api.GroupStart();
            for (int d = 0; d < world && r == 0; ++d) {
                cudaSetDevice(dev[d]->device);
                r = st->api.AllGather(pd[d].local, pd[d].all, (size_t)width, kNcclUint32, st->comms[d], dev[d]->stream);
            }
            const ncclResult_t r2 = st->api.GroupEnd();
            done = r == 0 && r2 == 0;
            for (int d = 0; d < world; ++d) {
                cudaSetDevice(dev[d]->device);
                if (cudaStreamSynchronize(dev[d]->stream) != cudaSuccess) done = false;
            }
            dev[0]->launches += 1;                                 // the all-gather kernel on this device
        }
        if (!done) {                                              // fallback: gather onto the first device by peer copies
            cudaSetDevice(dev[0]->device);
            for (int d = 0; d < world; ++d)
                if (cudaMemcpyPeerAsync(pd[0].all + (size_t)d * width, dev[0]->device, pd[d].local, dev[d]->device,
                                        (size_t)width * 4, dev[0]->stream) != cudaSuccess) {
                    cleanup();
                    return lsx_fail(ctx, LSX_ERR_CUDA, "det_large: gathering the residues failed");
                }
        }
        gathered = pd[0].all;
    }
    // ---- CRT on the first device ----
    cudaSetDevice(dev[0]->device);
    uint32_t* flat = nullptr;
    if (cudaMalloc(&flat, (size_t)K * 4 + (size_t)limbs * 4) != cudaSuccess) {
        cleanup();
        return lsx_fail(ctx, LSX_ERR_CUDA, "det_large: device allocation failed");
    }
    k_compact<<<(K + 255) / 256, 256, 0, ctx->stream>>>(gathered, width, world, K, flat);
    ctx->launches++;
    rc = lsx_crt_signed(ctx, flat, K, limbs, LSX_MEM_DEVICE, flat + K);
    if (rc == LSX_OK) {
        cudaError_t e = cudaMemcpyAsync(det_words, flat + K, (size_t)limbs * 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = lsx_fail(ctx, LSX_ERR_CUDA, "det_large: D2H copy failed: %s", cudaGetErrorString(e));
    }
    cudaFree(flat);
    cleanup();
    return rc;
}

}  // extern "C"
