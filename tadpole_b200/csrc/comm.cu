// comm.cu -- multi-GPU plumbing of one context: one process per GPU, NCCL over NVLink / NVSwitch.
//
// The path shards at three levels.  Independent calls and chromosome arms need no collective at all.
// Inside one call on a large matrix the row blocks of the correlation matrix, of M = Xc Xc^T and of every operator
// application of the subspace iteration are computed by their owner rank and all-gathered (in place, stream ordered
// with the kernels), the candidates of the sweep are dealt out rank-interleaved, and the per-candidate level counts
// and Calinski-Harabasz rows are combined with one small all-reduce.  Everything else (b x b problems, rotations,
// Gram matrices) is replicated: every rank runs the same deterministic kernels on the same bits, so all ranks take
// the same host-side decisions without exchanging them.
//
// NCCL is bound at run time (dlopen): the single-GPU path never touches it, and a process that already carries an
// NCCL (torch's) shares that copy instead of loading a second one.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <mutex>

struct TpNccl {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
};
static TpNccl g_nccl;

static int nccl_bind() {
    static std::mutex mu;                 // several rank threads of one process may arrive together
    std::lock_guard<std::mutex> lock(mu);
    if (g_nccl.lib) return TP_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) { h = dlopen(nm, RTLD_NOW | RTLD_NOLOAD); if (h) break; }     // a copy already in the process
    // RTLD_LOCAL: a copy loaded here must not satisfy the NCCL symbols of a library loaded later (torch imported after the
    // first multi-device context would otherwise bind to the system's older libnccl and fail on symbols it lacks)
    for (const char *nm : names) { if (h) break; h = dlopen(nm, RTLD_NOW | RTLD_LOCAL); }
    if (!h) { tp_set_error("NCCL not found (%s); multi-GPU calls need libnccl.so.2", dlerror()); return TP_ERR_CUDA; }
#define BIND(f)                                                                  \
    do {                                                                         \
        *(void **)(&g_nccl.f) = dlsym(h, "nccl" #f);                             \
        if (!g_nccl.f) { tp_set_error("NCCL symbol nccl" #f " missing"); return TP_ERR_CUDA; } \
    } while (0)
    BIND(GetUniqueId); BIND(CommInitRank); BIND(CommDestroy); BIND(GetErrorString);
    BIND(AllGather); BIND(AllReduce); BIND(Broadcast); BIND(CommInitAll); BIND(GroupStart); BIND(GroupEnd);
#undef BIND
    g_nccl.lib = h;
    return TP_OK;
}

#define TP_NCCL(call)                                                                                     \
    do {                                                                                                  \
        ncclResult_t r_ = (call);                                                                         \
        if (r_ != ncclSuccess) {                                                                          \
            tp_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_));       \
            return TP_ERR_CUDA;                                                                           \
        }                                                                                                 \
    } while (0)

extern "C" int tp_comm_unique_id(void *id128) {
    TP_ARG(id128, "tp_comm_unique_id: null argument");
    TP_TRY(nccl_bind());
    ncclUniqueId id;
    TP_NCCL(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, sizeof(id));
    return TP_OK;
}

extern "C" int tp_ctx_comm_init(tp_ctx *ctx, const void *id128, int rank, int nranks, int slot) {
    TP_ARG(ctx && id128 && nranks >= 1 && rank >= 0 && rank < nranks, "tp_ctx_comm_init: bad arguments");
    TP_ARG(slot >= 0 && slot < TP_COMM_SLOTS, "tp_ctx_comm_init: slot out of range");
    TP_ARG(!ctx->comm[slot].handle, "tp_ctx_comm_init: slot already holds a communicator");
    TP_CUDA(cudaSetDevice(ctx->device));
    TP_TRY(nccl_bind());
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    TP_NCCL(g_nccl.CommInitRank(&c, nranks, id, rank));
    ctx->comm[slot].handle = (void *)c;
    ctx->comm[slot].rank = rank;
    ctx->comm[slot].nranks = nranks;
    return TP_OK;
}

extern "C" int tp_ctx_comm_select(tp_ctx *ctx, int slot) {
    TP_ARG(ctx && slot >= -1 && slot < TP_COMM_SLOTS, "tp_ctx_comm_select: slot out of range");
    TP_ARG(slot < 0 || ctx->comm[slot].handle, "tp_ctx_comm_select: empty slot");
    ctx->comm_cur = slot;
    return TP_OK;
}

extern "C" int tp_ctx_comm_info(tp_ctx *ctx, int *rank_out, int *nranks_out) {
    TP_ARG(ctx, "tp_ctx_comm_info: null context");
    if (rank_out) *rank_out = tp_rank(ctx);
    if (nranks_out) *nranks_out = tp_nranks(ctx);
    return TP_OK;
}

int tp_comm_destroy_all(tp_ctx *ctx) {
    for (int s = 0; s < TP_COMM_SLOTS; s++)
        if (ctx->comm[s].handle) { g_nccl.CommDestroy((ncclComm_t)ctx->comm[s].handle); ctx->comm[s].handle = nullptr; }
    ctx->comm_cur = -1;
    return TP_OK;
}

// buf holds nranks chunks of `chunk` doubles; chunk `rank` is valid on entry, all of them on return
int tp_comm_allgather(tp_ctx *ctx, double *buf, size_t chunk) {
    if (tp_nranks(ctx) == 1) return TP_OK;
    const TpCommSlot &c = ctx->comm[ctx->comm_cur];
    tp_prof_begin(ctx, PC_COMM);
    TP_NCCL(g_nccl.AllGather(buf + (size_t)c.rank * chunk, buf, chunk, ncclDouble, (ncclComm_t)c.handle, ctx->stream));
    tp_prof_end(ctx);
    return TP_OK;
}
int tp_comm_allreduce_sum(tp_ctx *ctx, void *buf, size_t count, int is_double) {
    if (tp_nranks(ctx) == 1) return TP_OK;
    const TpCommSlot &c = ctx->comm[ctx->comm_cur];
    tp_prof_begin(ctx, PC_COMM);
    TP_NCCL(g_nccl.AllReduce(buf, buf, count, is_double ? ncclDouble : ncclInt32, ncclSum, (ncclComm_t)c.handle, ctx->stream));
    tp_prof_end(ctx);
    return TP_OK;
}
int tp_comm_bcast(tp_ctx *ctx, double *buf, size_t count, int root) {
    if (tp_nranks(ctx) == 1) return TP_OK;
    const TpCommSlot &c = ctx->comm[ctx->comm_cur];
    tp_prof_begin(ctx, PC_COMM);
    TP_NCCL(g_nccl.Broadcast(buf, buf, count, ncclDouble, root, (ncclComm_t)c.handle, ctx->stream));
    tp_prof_end(ctx);
    return TP_OK;
}

// raw bytes, in place (the root's buffer is the source, everyone else's the destination)
int tp_comm_bcast_bytes(tp_ctx *ctx, void *buf, size_t bytes, int root) {
    if (tp_nranks(ctx) == 1) return TP_OK;
    const TpCommSlot &c = ctx->comm[ctx->comm_cur];
    if (!ctx->comm_grouped) tp_prof_begin(ctx, PC_COMM);
    TP_NCCL(g_nccl.Broadcast(buf, buf, bytes, ncclChar, root, (ncclComm_t)c.handle, ctx->stream));
    if (!ctx->comm_grouped) tp_prof_end(ctx);
    return TP_OK;
}
// the collectives between begin and end are issued as one NCCL group (timed as one)
int tp_comm_group_begin(tp_ctx *ctx) {
    if (tp_nranks(ctx) == 1) return TP_OK;
    tp_prof_begin(ctx, PC_COMM);
    ctx->comm_grouped = true;
    TP_NCCL(g_nccl.GroupStart());
    return TP_OK;
}
int tp_comm_group_end(tp_ctx *ctx) {
    if (tp_nranks(ctx) == 1) return TP_OK;
    ctx->comm_grouped = false;
    TP_NCCL(g_nccl.GroupEnd());
    tp_prof_end(ctx);
    return TP_OK;
}

// One process driving several GPUs (multi-device contexts, group.cu): all communicators of the group at once, member i = rank i
int tp_comm_init_all(tp_ctx **members, int nmembers, int slot) {
    TP_ARG(members && nmembers >= 2 && slot >= 0 && slot < TP_COMM_SLOTS, "tp_comm_init_all: bad arguments");
    TP_TRY(nccl_bind());
    std::vector<int> devs(nmembers);
    std::vector<ncclComm_t> comms(nmembers, nullptr);
    for (int i = 0; i < nmembers; i++) {
        TP_ARG(!members[i]->comm[slot].handle, "tp_comm_init_all: slot already holds a communicator");
        devs[i] = members[i]->device;
    }
    TP_NCCL(g_nccl.CommInitAll(comms.data(), nmembers, devs.data()));
    for (int i = 0; i < nmembers; i++) {
        members[i]->comm[slot].handle = (void *)comms[i];
        members[i]->comm[slot].rank = i;
        members[i]->comm[slot].nranks = nmembers;
    }
    return TP_OK;
}
