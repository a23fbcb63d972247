// The ONE exchange step of the path (north_star: "NCCL over NVLink used only for the overlap-region weight-sum
// allreduce"; SURVEY 8e view-sharded mode): every rank accumulates its views' partial [sum w*I, sum w] for a block
// region (bs_fuse_accumulate), bs_fuse_allreduce sums both buffers across the ranks in place -- one grouped NCCL
// all-reduce on the context's stream, so it is ordered after the accumulate kernels and before bs_fuse_finish with no
// host synchronisation -- and bs_fuse_finish divides and converts.
//
// NCCL is bound at run time (dlopen "libnccl.so.2": the copy torch already loaded in a torchrun process, else the
// system one), so libbsgpu.so has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>

#include <cstring>

#include "bs_internal.cuh"

namespace {

typedef struct { char internal[128]; } nccl_unique_id;     // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* nccl_comm_t;
constexpr int kNcclFloat32 = 7, kNcclSum = 0;              // ncclDataType_t / ncclRedOp_t values of nccl.h

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(nccl_unique_id*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (a.lib) break;
    }
    if (!a.lib) return a;
    a.GetUniqueId = (int (*)(nccl_unique_id*))dlsym(a.lib, "ncclGetUniqueId");
    a.CommInitRank = (int (*)(nccl_comm_t*, int, nccl_unique_id, int))dlsym(a.lib, "ncclCommInitRank");
    a.CommDestroy = (int (*)(nccl_comm_t))dlsym(a.lib, "ncclCommDestroy");
    a.AllReduce = (int (*)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(a.lib, "ncclAllReduce");
    a.GroupStart = (int (*)())dlsym(a.lib, "ncclGroupStart");
    a.GroupEnd = (int (*)())dlsym(a.lib, "ncclGroupEnd");
    a.GetErrorString = (const char* (*)(int))dlsym(a.lib, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GroupStart && a.GroupEnd && a.GetErrorString;
    return a;
}

#define BS_NCCL(ctx, call)                                                                                      \
    do {                                                                                                        \
        int r__ = (call);                                                                                       \
        if (r__ != 0) return bs_set_error((ctx), BS_ERR_CUDA, "%s failed: %s", #call, nccl().GetErrorString(r__)); \
    } while (0)

}  // namespace

void bs_comm_free(bs_ctx* ctx) {
    if (ctx->nccl_comm && nccl().ok) nccl().CommDestroy((nccl_comm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
}

extern "C" {

int bs_comm_unique_id(unsigned char id[128]) {
    if (!id) return BS_ERR_ARG;
    if (!nccl().ok) return bs_set_error(nullptr, BS_ERR_UNSUPPORTED, "bs_comm_unique_id: libnccl.so.2 not found");
    nccl_unique_id u;
    const int r = nccl().GetUniqueId(&u);
    if (r != 0) return bs_set_error(nullptr, BS_ERR_CUDA, "ncclGetUniqueId: %s", nccl().GetErrorString(r));
    memcpy(id, u.internal, 128);
    return BS_OK;
}

int bs_comm_init(bs_ctx* ctx, int n_ranks, int rank, const unsigned char id[128]) {
    if (!ctx) return BS_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return bs_set_error(ctx, BS_ERR_ARG, "bs_comm_init: bad argument");
    if (!nccl().ok) return bs_set_error(ctx, BS_ERR_UNSUPPORTED, "bs_comm_init: libnccl.so.2 not found");
    BS_CUDA(ctx, cudaSetDevice(ctx->device));
    bs_comm_free(ctx);
    nccl_unique_id u;
    memcpy(u.internal, id, 128);
    nccl_comm_t c = nullptr;
    BS_NCCL(ctx, nccl().CommInitRank(&c, n_ranks, u, rank));
    ctx->nccl_comm = c;
    ctx->nccl_ranks = n_ranks;
    return BS_OK;
}

int bs_comm_destroy(bs_ctx* ctx) {
    if (!ctx) return BS_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    BS_CUDA(ctx, cudaSetDevice(ctx->device));
    BS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    bs_comm_free(ctx);
    return BS_OK;
}

int bs_fuse_allreduce(bs_ctx* ctx, float* sum_wi_dev, float* sum_w_dev, long long n) {
    if (!ctx) return BS_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!sum_wi_dev || !sum_w_dev || n <= 0) return bs_set_error(ctx, BS_ERR_ARG, "bs_fuse_allreduce: bad argument");
    if (!ctx->nccl_comm) return bs_set_error(ctx, BS_ERR_ARG, "bs_fuse_allreduce: call bs_comm_init first");
    BS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->nccl_ranks == 1) return BS_OK;
    // the two buffers travel as one group: a single fused launch on the NVLink / NVSwitch fabric
    BS_NCCL(ctx, nccl().GroupStart());
    BS_NCCL(ctx, nccl().AllReduce(sum_wi_dev, sum_wi_dev, (size_t)n, kNcclFloat32, kNcclSum, (nccl_comm_t)ctx->nccl_comm, ctx->stream));
    BS_NCCL(ctx, nccl().AllReduce(sum_w_dev, sum_w_dev, (size_t)n, kNcclFloat32, kNcclSum, (nccl_comm_t)ctx->nccl_comm, ctx->stream));
    BS_NCCL(ctx, nccl().GroupEnd());
    ctx->launches++;
    return BS_OK;
}

}  // extern "C"
