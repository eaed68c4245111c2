// Multi-GPU plumbing: one process per GPU, NCCL over NVLink.  The prove path shards by LDE coset ranges;
// the only data-path collectives are all-gathers at the commitment points (subtree roots, constraint accumulators,
// composition evaluations, opened values) plus one integer max-reduce (constraint-violation flag).
// NCCL is loaded at run time (dlopen of the libnccl.so.2 that torch already mapped, or $DG_NCCL_LIB), so the library has no
// link-time dependency on it and single-GPU use never touches it.
#include <dlfcn.h>
#include "common.cuh"

namespace dg {

namespace {
struct ncclUniqueId { char internal[128]; };
typedef void *ncclComm_t;
typedef int (*fn_GetUniqueId)(ncclUniqueId *);
typedef int (*fn_CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
typedef int (*fn_AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
typedef int (*fn_AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*fn_CommDestroy)(ncclComm_t);
typedef int (*fn_SendRecv)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*fn_Group)(void);
typedef const char *(*fn_GetErrorString)(int);

struct Nccl {
    void *handle = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr;
    fn_CommInitRank CommInitRank = nullptr;
    fn_AllGather AllGather = nullptr;
    fn_AllReduce AllReduce = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
    fn_SendRecv Send = nullptr, Recv = nullptr;
    fn_Group GroupStart = nullptr, GroupEnd = nullptr;
    fn_GetErrorString GetErrorString = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
} g_nccl;

void load_nccl() {
    if (g_nccl.handle) return;
    const char *env = getenv("DG_NCCL_LIB");
    const char *candidates[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char *name : candidates) {
        if (!name) continue;
        g_nccl.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) throw Error(-2, std::string("cannot load NCCL (set DG_NCCL_LIB): ") + dlerror());
    auto sym = [&](const char *n) { void *p = dlsym(g_nccl.handle, n); if (!p) throw Error(-2, std::string("NCCL symbol missing: ") + n); return p; };
    g_nccl.GetUniqueId = (fn_GetUniqueId)sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (fn_CommInitRank)sym("ncclCommInitRank");
    g_nccl.AllGather = (fn_AllGather)sym("ncclAllGather");
    g_nccl.AllReduce = (fn_AllReduce)sym("ncclAllReduce");
    g_nccl.CommDestroy = (fn_CommDestroy)sym("ncclCommDestroy");
    g_nccl.Send = (fn_SendRecv)sym("ncclSend");
    g_nccl.Recv = (fn_SendRecv)sym("ncclRecv");
    g_nccl.GroupStart = (fn_Group)sym("ncclGroupStart");
    g_nccl.GroupEnd = (fn_Group)sym("ncclGroupEnd");
    g_nccl.GetErrorString = (fn_GetErrorString)sym("ncclGetErrorString");
    g_nccl.CommInitAll = (int (*)(ncclComm_t *, int, const int *))sym("ncclCommInitAll");
}
void nccl_check(int rc, const char *what) {
    if (rc != 0) throw Error(-2, std::string("NCCL ") + what + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
}
}  // namespace

void comm_unique_id(uint8_t out[128]) {
    load_nccl();
    ncclUniqueId id;
    nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(out, id.internal, 128);
}

void comm_init(Context &c, int rank, int world, const uint8_t id_bytes[128]) {
    DG_REQUIRE(world == 1 || world == 2 || world == 4 || world == 8, "world size must be 1, 2, 4 or 8");
    DG_REQUIRE(rank >= 0 && rank < world, "invalid rank");
    if (world > 1) {
        load_nccl();
        ncclUniqueId id;
        memcpy(id.internal, id_bytes, 128);
        if (c.nccl_comm) { g_nccl.CommDestroy(c.nccl_comm); c.nccl_comm = nullptr; }      // a second dg_comm_init replaces the communicator
        nccl_check(g_nccl.CommInitRank(&c.nccl_comm, world, id, rank), "ncclCommInitRank");
    }
    c.rank = rank;
    c.world = world;
}

// single process, one context per device: one communicator per context, created together
void comm_init_all(std::vector<Context *> &ctxs) {
    load_nccl();
    const int n = (int)ctxs.size();
    std::vector<ncclComm_t> comms(n);
    std::vector<int> devs(n);
    for (int i = 0; i < n; i++) devs[i] = ctxs[i]->device;
    nccl_check(g_nccl.CommInitAll(comms.data(), n, devs.data()), "ncclCommInitAll");
    for (int i = 0; i < n; i++) { ctxs[i]->nccl_comm = comms[i]; ctxs[i]->rank = i; ctxs[i]->world = n; }
}

void comm_finalize(Context &c) {
    if (c.nccl_comm) { g_nccl.CommDestroy(c.nccl_comm); c.nccl_comm = nullptr; }
    c.rank = 0;
    c.world = 1;
}

// recv = concatenation over ranks of each rank's `bytes` bytes (rank-major); with one rank it is a device copy
void comm_all_gather(Context &c, const void *send, void *recv, size_t bytes, cudaStream_t stream) {
    if (!stream) stream = c.stream;
    if (c.world == 1) {
        if (send != recv) DG_CUDA(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, stream));
        return;
    }
    nccl_check(g_nccl.AllGather(send, recv, bytes, /*ncclUint8*/ 1, (ncclComm_t)c.nccl_comm, stream), "ncclAllGather");
}

// recv[g] (bytes each) = the chunk rank g sent to this rank; send[h] = the chunk for rank h.  One grouped send/recv per peer.
void comm_all_to_all(Context &c, const void *send, void *recv, size_t bytes) {
    if (c.world == 1) {
        if (send != recv) DG_CUDA(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, c.stream));
        return;
    }
    nccl_check(g_nccl.GroupStart(), "ncclGroupStart");
    for (int p = 0; p < c.world; p++) {
        nccl_check(g_nccl.Send((const uint8_t *)send + (size_t)p * bytes, bytes, /*ncclUint8*/ 1, p, (ncclComm_t)c.nccl_comm, c.stream), "ncclSend");
        nccl_check(g_nccl.Recv((uint8_t *)recv + (size_t)p * bytes, bytes, /*ncclUint8*/ 1, p, (ncclComm_t)c.nccl_comm, c.stream), "ncclRecv");
    }
    nccl_check(g_nccl.GroupEnd(), "ncclGroupEnd");
}

// recv = element-wise sum over the ranks of `count` uint32 values
void comm_all_reduce_sum_u32(Context &c, const unsigned *send, unsigned *recv, size_t count) {
    if (c.world == 1) {
        if (send != recv) DG_CUDA(cudaMemcpyAsync(recv, send, count * 4, cudaMemcpyDeviceToDevice, c.stream));
        return;
    }
    nccl_check(g_nccl.AllReduce(send, recv, count, /*ncclUint32*/ 3, /*ncclSum*/ 0, (ncclComm_t)c.nccl_comm, c.stream), "ncclAllReduce");
}

// in-place max over ranks of `count` uint32 values
void comm_all_reduce_max_u32(Context &c, unsigned *buf, size_t count) {
    if (c.world == 1) return;
    nccl_check(g_nccl.AllReduce(buf, buf, count, /*ncclUint32*/ 3, /*ncclMax*/ 2, (ncclComm_t)c.nccl_comm, c.stream), "ncclAllReduce");
}

}  // namespace dg
