// Multi-GPU plumbing: NCCL over NVLink 5 / NVSwitch, one process per GPU.
// The reference has no distributed path; this implements SURVEY.md 8(e):
// halo exchange of interface node values (ncclSend/ncclRecv in one group, <= 2 neighbours for slab
// partitions) and packed dot-product all-reduces (<= 5 doubles).
#include "tefem_common.cuh"
#include "tefem_comm.h"
#include "tefem_peer.cuh"

#include <dlfcn.h>
#include <nccl.h>

namespace tefem {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int load_nccl(const char* path) {
    if (g_nccl.handle) return TEFEM_OK;
    void* h = nullptr;
    if (path && path[0]) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // the copy torch already loaded
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error("cannot load NCCL (%s): %s", path ? path : "libnccl.so.2", dlerror());
        return TEFEM_ERR_NCCL;
    }
#define TEFEM_SYM(field, name)                                            \
    *(void**)(&g_nccl.field) = dlsym(h, name);                            \
    if (!g_nccl.field) { set_error("NCCL symbol %s missing", name); return TEFEM_ERR_NCCL; }
    TEFEM_SYM(GetUniqueId, "ncclGetUniqueId")
    TEFEM_SYM(CommInitRank, "ncclCommInitRank")
    TEFEM_SYM(CommDestroy, "ncclCommDestroy")
    TEFEM_SYM(AllReduce, "ncclAllReduce")
    TEFEM_SYM(Send, "ncclSend")
    TEFEM_SYM(Recv, "ncclRecv")
    TEFEM_SYM(GroupStart, "ncclGroupStart")
    TEFEM_SYM(GroupEnd, "ncclGroupEnd")
    TEFEM_SYM(GetErrorString, "ncclGetErrorString")
#undef TEFEM_SYM
    g_nccl.handle = h;
    return TEFEM_OK;
}

#define TEFEM_NCCL_CHECK(expr)                                                                   \
    do {                                                                                         \
        ncclResult_t _r = (expr);                                                                \
        if (_r != ncclSuccess) {                                                                 \
            tefem::set_error("%s failed: %s", #expr, tefem::g_nccl.GetErrorString(_r));          \
            return TEFEM_ERR_NCCL;                                                               \
        }                                                                                        \
    } while (0)

}  // namespace tefem

using namespace tefem;

struct tefem_comm {
    ncclComm_t comm = nullptr;
    int n_ranks = 1, rank = 0;
    // peer-memory region (tefem_peer.cuh): local allocation, the mapped regions of all ranks, epochs per channel
    char* shm_local = nullptr;
    char* shm_base[PEER_MAX_RANKS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int64_t shm_cap = 0;
    uint32_t epoch[PEER_N_CHANNELS] = {0, 0, 0, 0};
    int* d_err = nullptr;
    long long timeout = PEER_TIMEOUT_DEFAULT;
    // peer-mapped vector slots (SpMV inputs whose halo tails the neighbours write directly, tefem_krylov.cu)
    char* vec_local = nullptr;
    char* vec_base[PEER_MAX_RANKS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int n_slots = 0;
    int64_t slot_doubles = 0;
};

extern "C" int tefem_comm_unique_id(const char* h_nccl_lib_path, char* h_id) {
    TEFEM_REQUIRE(h_id, "null id buffer");
    int rc = load_nccl(h_nccl_lib_path);
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    TEFEM_NCCL_CHECK(g_nccl.GetUniqueId(&id));
    memcpy(h_id, &id, sizeof(id));
    return TEFEM_OK;
}

extern "C" int tefem_comm_create(tefem_comm** out, const char* h_nccl_lib_path, const char* h_id, int n_ranks, int rank) {
    TEFEM_REQUIRE(out && h_id && n_ranks > 0 && rank >= 0 && rank < n_ranks, "arguments");
    int rc = load_nccl(h_nccl_lib_path);
    if (rc) return rc;
    ncclUniqueId id;
    memcpy(&id, h_id, sizeof(id));
    tefem_comm* c = new tefem_comm();
    c->n_ranks = n_ranks; c->rank = rank;
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, n_ranks, id, rank);
    if (r != ncclSuccess) {
        set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
        delete c;
        return TEFEM_ERR_NCCL;
    }
    *out = c;
    return TEFEM_OK;
}

extern "C" int tefem_comm_destroy(tefem_comm* c) {
    if (!c) return TEFEM_OK;
    for (int r = 0; r < c->n_ranks && r < PEER_MAX_RANKS; ++r) {
        if (r != c->rank && c->shm_base[r]) cudaIpcCloseMemHandle(c->shm_base[r]);
        if (r != c->rank && c->vec_base[r]) cudaIpcCloseMemHandle(c->vec_base[r]);
    }
    if (c->shm_local) cudaFree(c->shm_local);
    if (c->vec_local) cudaFree(c->vec_local);
    if (c->d_err) cudaFree(c->d_err);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    delete c;
    return TEFEM_OK;
}

extern "C" int tefem_comm_allreduce_sum(tefem_comm* c, double* buf, int n, void* stream) {
    TEFEM_REQUIRE(c && buf && n > 0, "arguments");
    TEFEM_NCCL_CHECK(g_nccl.AllReduce(buf, buf, (size_t)n, ncclDouble, ncclSum, c->comm, as_stream(stream)));
    return TEFEM_OK;
}

int tefem_comm_neighbor_exchange(tefem_comm* c, int n_nbr, const int32_t* nbr, const int32_t* send_ptr,
                                 const double* send, const int32_t* recv_ptr, double* recv, cudaStream_t st) {
    TEFEM_REQUIRE(c, "null communicator");
    TEFEM_NCCL_CHECK(g_nccl.GroupStart());
    for (int k = 0; k < n_nbr; ++k) {
        const size_t ns = 4 * (size_t)(send_ptr[k + 1] - send_ptr[k]), nr = 4 * (size_t)(recv_ptr[k + 1] - recv_ptr[k]);
        if (ns) TEFEM_NCCL_CHECK(g_nccl.Send(send + 4 * (size_t)send_ptr[k], ns, ncclDouble, nbr[k], c->comm, st));
        if (nr) TEFEM_NCCL_CHECK(g_nccl.Recv(recv + 4 * (size_t)recv_ptr[k], nr, ncclDouble, nbr[k], c->comm, st));
    }
    TEFEM_NCCL_CHECK(g_nccl.GroupEnd());
    return TEFEM_OK;
}

// ---- peer-memory region --------------------------------------------------------------------------

extern "C" int tefem_comm_shm_alloc(tefem_comm* c, int64_t big_cap_doubles, char* h_handle) {
    TEFEM_REQUIRE(c && h_handle && big_cap_doubles >= 0, "arguments");
    TEFEM_REQUIRE(c->n_ranks <= PEER_MAX_RANKS, "peer-memory reductions support up to 8 ranks (one node)");
    TEFEM_REQUIRE(!c->shm_local, "peer-memory region already allocated");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    const int64_t cap = (big_cap_doubles + 15) / 16 * 16;
    const size_t bytes = (size_t)PEER_BIG_OFFSET + 2 * (size_t)cap * sizeof(double);
    TEFEM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&c->shm_local), bytes));
    TEFEM_CUDA_CHECK(cudaMemset(c->shm_local, 0, bytes));
    TEFEM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&c->d_err), sizeof(int)));
    TEFEM_CUDA_CHECK(cudaMemset(c->d_err, 0, sizeof(int)));
    TEFEM_CUDA_CHECK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    TEFEM_CUDA_CHECK(cudaIpcGetMemHandle(&h, c->shm_local));
    memcpy(h_handle, &h, sizeof(h));
    c->shm_cap = cap;
    return TEFEM_OK;
}

extern "C" int tefem_comm_shm_open(tefem_comm* c, const char* h_handles) {
    TEFEM_REQUIRE(c && h_handles && c->shm_local, "allocate the local region first");
    for (int r = 0; r < c->n_ranks; ++r) {
        if (r == c->rank) { c->shm_base[r] = c->shm_local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, h_handles + 64 * (size_t)r, sizeof(h));
        void* p = nullptr;
        TEFEM_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->shm_base[r] = reinterpret_cast<char*>(p);
    }
    return TEFEM_OK;
}

// ---- peer-mapped vector slots ----------------------------------------------------------------------
// Every rank allocates n_slots vectors of slot_doubles doubles (the same sizes on all ranks, so that slot k of rank q
// sits at vec_base[q] + k * slot_doubles) and maps the allocations of the others with CUDA IPC.  The Krylov driver
// keeps its SpMV input vectors there: a neighbour writes the halo tail of my vector with plain remote stores.
extern "C" int tefem_comm_vec_alloc(tefem_comm* c, int n_slots, int64_t slot_doubles, char* h_handle) {
    TEFEM_REQUIRE(c && h_handle && n_slots > 0 && slot_doubles > 0, "arguments");
    TEFEM_REQUIRE(c->n_ranks <= PEER_MAX_RANKS, "peer-memory vectors support up to 8 ranks (one node)");
    TEFEM_REQUIRE(!c->vec_local, "peer-memory vectors already allocated");
    const int64_t slot = (slot_doubles + 15) / 16 * 16;
    const size_t bytes = (size_t)n_slots * (size_t)slot * sizeof(double);
    TEFEM_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&c->vec_local), bytes));
    TEFEM_CUDA_CHECK(cudaMemset(c->vec_local, 0, bytes));
    TEFEM_CUDA_CHECK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    TEFEM_CUDA_CHECK(cudaIpcGetMemHandle(&h, c->vec_local));
    memcpy(h_handle, &h, sizeof(h));
    c->n_slots = n_slots;
    c->slot_doubles = slot;
    return TEFEM_OK;
}

extern "C" int tefem_comm_vec_open(tefem_comm* c, const char* h_handles) {
    TEFEM_REQUIRE(c && h_handles && c->vec_local, "allocate the local vectors first");
    for (int r = 0; r < c->n_ranks; ++r) {
        if (r == c->rank) { c->vec_base[r] = c->vec_local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, h_handles + 64 * (size_t)r, sizeof(h));
        void* p = nullptr;
        TEFEM_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->vec_base[r] = reinterpret_cast<char*>(p);
    }
    return TEFEM_OK;
}

bool tefem_comm_has_vectors(const tefem_comm* c) {
    return c && c->vec_local && c->vec_base[c->n_ranks - 1] != nullptr && c->vec_base[0] != nullptr;
}
int tefem_comm_n_slots(const tefem_comm* c) { return c ? c->n_slots : 0; }
int64_t tefem_comm_slot_doubles(const tefem_comm* c) { return c ? c->slot_doubles : 0; }
double* tefem_comm_slot(const tefem_comm* c, int rank, int slot) {
    return reinterpret_cast<double*>(c->vec_base[rank]) + (int64_t)slot * c->slot_doubles;
}
const int* tefem_comm_err_ptr(const tefem_comm* c) { return c ? c->d_err : nullptr; }
int tefem_comm_report_timeout() {
    set_error("peer-memory exchange timed out waiting for another rank (communicator marked failed; "
              "tefem_comm_reset_error after all ranks have been resynchronised)");
    return TEFEM_ERR_NCCL;
}
int tefem_comm_rank(const tefem_comm* c) { return c ? c->rank : 0; }
int tefem_comm_world(const tefem_comm* c) { return c ? c->n_ranks : 1; }

extern "C" int tefem_comm_set_timeout(tefem_comm* c, double seconds) {
    TEFEM_REQUIRE(c && seconds > 0.0, "arguments");
    c->timeout = (long long)(seconds * 1.9e9);          // cycles of the SM clock at its nominal maximum
    return TEFEM_OK;
}

bool tefem_comm_has_peer(const tefem_comm* c) { return c && c->shm_base[c->n_ranks - 1] != nullptr && c->shm_base[0] != nullptr; }
int64_t tefem_comm_peer_cap(const tefem_comm* c) { return c ? c->shm_cap : 0; }

PeerCtx tefem_comm_peer_ctx(tefem_comm* c, int channel) {
    PeerCtx ctx;
    ctx.world = c->n_ranks; ctx.rank = c->rank; ctx.channel = channel; ctx.cap = c->shm_cap;
    ctx.epoch = ++c->epoch[channel];
    for (int r = 0; r < c->n_ranks; ++r) ctx.base[r] = c->shm_base[r];
    ctx.err = c->d_err;
    ctx.timeout = c->timeout;
    return ctx;
}

// Call after a stream synchronisation: reports a time-out recorded by a peer wait.  The flag stays set (every later
// solve on this communicator fails fast instead of computing with partial sums) until tefem_comm_reset_error.
extern "C" int tefem_comm_check(tefem_comm* c) {
    if (!c || !c->d_err) return TEFEM_OK;
    int e = 0;
    TEFEM_CUDA_CHECK(cudaMemcpy(&e, c->d_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (e) {
        set_error("peer-memory exchange timed out waiting for another rank (communicator marked failed; "
                  "tefem_comm_reset_error after all ranks have been resynchronised)");
        return TEFEM_ERR_NCCL;
    }
    return TEFEM_OK;
}

extern "C" int tefem_comm_reset_error(tefem_comm* c) {
    TEFEM_REQUIRE(c, "null communicator");
    if (c->d_err) TEFEM_CUDA_CHECK(cudaMemset(c->d_err, 0, sizeof(int)));
    return TEFEM_OK;
}
