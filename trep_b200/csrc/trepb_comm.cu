// Multi-GPU part of the C ABI (include/trepb.h, "multi-GPU"): one process per GPU.
//
// The batched MidpointVI path shards with no exchange during compute; the one
// exchange it has is collecting the A / B slabs of DSystem.linearize_trajectory
// (trep/discopt/dsystem.py:406-423) on the GPU that runs the sequential Riccati sweep
// (trep/discopt/dlqr.py:9-81).  Two ways, both device-to-device over NVLink / NVSwitch:
//
//   trepb_comm_allgather_dev / trepb_comm_gather_dev   NCCL (ncclAllGather; grouped ncclSend/ncclRecv)
//       on slabs already written to local HBM.  NCCL is loaded at run time (dlopen of libnccl.so.2:
//       the copy already in the process if there is one - e.g. the one a torch import brought - else
//       the system library), so libtrepb.so has no link-time dependency on it.
//   trepb_ipc_export / trepb_ipc_open                   the root's slab mapped into every rank (CUDA IPC,
//       peer access over NVLink): a rank passes `mapped_root_slab + its offset` as the A / B output
//       pointers of trepb_linearize_batch_dev and the linearize kernel's own stores land in the root's
//       HBM - the gather is fused into the kernel, there is no second pass over the data.
#include <cuda_runtime.h>
#include "trepb_nvtx.h"
#include <dlfcn.h>
#include <nccl.h>   // types and enums only; every function is looked up with dlsym
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <string>

#include "../../include/trepb.h"
#include "trepb_err.h"

using namespace trepb;

namespace {
int cfail(int code, const std::string& m) { last_error() = m; return code; }

struct Nccl {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string why;
};

Nccl& nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[3] = {getenv("TREPB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            // RTLD_NOLOAD first: reuse a copy that is already mapped (one NCCL per process)
            n.h = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
            if (!n.h) n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (n.h) break;
        }
        if (!n.h) { n.why = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found"); return; }
#define SYM(field, name) *(void**)(&n.field) = dlsym(n.h, name); if (!n.field) { n.why = std::string("libnccl lacks ") + name; n.h = nullptr; return; }
        SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
        SYM(AllGather, "ncclAllGather") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart")
        SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString") SYM(GetVersion, "ncclGetVersion")
#undef SYM
    });
    return n;
}

int nccl_fail(ncclResult_t r, const char* what) {
    Nccl& n = nccl();
    return cfail(TREPB_ERR_CUDA, std::string(what) + ": " + (n.GetErrorString ? n.GetErrorString(r) : "NCCL error"));
}
#define NC(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return nccl_fail(r_, #call); } while (0)
#define CUC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cfail(TREPB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
}  // namespace

struct trepb_comm {
    int device = 0, rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
};

extern "C" {

int trepb_comm_available(int* version) {
    TREPB_NVTX("trepb_comm_available");
    Nccl& n = nccl();
    if (!n.h) return cfail(TREPB_ERR_UNSUPPORTED, n.why);
    if (version) { int v = 0; n.GetVersion(&v); *version = v; }
    return TREPB_OK;
}

int trepb_comm_unique_id(char* id) {
    TREPB_NVTX("trepb_comm_unique_id");
    static_assert(sizeof(ncclUniqueId) == TREPB_COMM_ID_BYTES, "TREPB_COMM_ID_BYTES must be sizeof(ncclUniqueId)");
    if (!id) return cfail(TREPB_ERR_INVALID, "null argument");
    Nccl& n = nccl();
    if (!n.h) return cfail(TREPB_ERR_UNSUPPORTED, n.why);
    ncclUniqueId u;
    NC(n.GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return TREPB_OK;
}

int trepb_comm_create(int device, int rank, int nranks, const char* id, trepb_comm** out) {
    TREPB_NVTX("trepb_comm_create");
    if (!out || !id) return cfail(TREPB_ERR_INVALID, "null argument");
    *out = nullptr;
    if (nranks < 1 || rank < 0 || rank >= nranks) return cfail(TREPB_ERR_INVALID, "rank must be in [0, nranks)");
    Nccl& n = nccl();
    if (!n.h) return cfail(TREPB_ERR_UNSUPPORTED, n.why);
    CUC(cudaSetDevice(device));
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    trepb_comm* c = new trepb_comm();
    c->device = device; c->rank = rank; c->nranks = nranks;
    ncclResult_t r = n.CommInitRank(&c->comm, nranks, u, rank);
    if (r != ncclSuccess) { delete c; return nccl_fail(r, "ncclCommInitRank"); }
    *out = c;
    return TREPB_OK;
}

void trepb_comm_destroy(trepb_comm* c) {
    if (!c) return;
    Nccl& n = nccl();
    if (n.h && c->comm) { cudaSetDevice(c->device); n.CommDestroy(c->comm); }
    delete c;
}

int trepb_comm_rank(const trepb_comm* c, int* rank, int* nranks) {
    TREPB_NVTX("trepb_comm_rank");
    if (!c) return cfail(TREPB_ERR_INVALID, "null communicator");
    if (rank) *rank = c->rank;
    if (nranks) *nranks = c->nranks;
    return TREPB_OK;
}

int trepb_comm_allgather_dev(trepb_comm* c, const void* send, void* recv, int64_t bytes_per_rank, void* stream) {
    TREPB_NVTX("trepb_comm_allgather_dev");
    if (!c || !recv || (!send && bytes_per_rank > 0)) return cfail(TREPB_ERR_INVALID, "null argument");
    if (bytes_per_rank < 0) return cfail(TREPB_ERR_INVALID, "bytes_per_rank must be >= 0");
    if (bytes_per_rank == 0) return TREPB_OK;
    Nccl& n = nccl();
    CUC(cudaSetDevice(c->device));
    // 8-byte elements when possible (every slab of doubles): fewer elements for NCCL's element loops
    if (bytes_per_rank % 8 == 0) NC(n.AllGather(send, recv, (size_t)bytes_per_rank / 8, ncclFloat64, c->comm, (cudaStream_t)stream));
    else NC(n.AllGather(send, recv, (size_t)bytes_per_rank, ncclUint8, c->comm, (cudaStream_t)stream));
    return TREPB_OK;
}

int trepb_comm_gather_dev(trepb_comm* c, const void* send, void* recv, int64_t bytes_per_rank, int root, void* stream) {
    TREPB_NVTX("trepb_comm_gather_dev");
    if (!c || (!send && bytes_per_rank > 0)) return cfail(TREPB_ERR_INVALID, "null argument");
    if (root < 0 || root >= c->nranks) return cfail(TREPB_ERR_INVALID, "root out of range");
    if (c->rank == root && !recv) return cfail(TREPB_ERR_INVALID, "the root needs a receive buffer");
    if (bytes_per_rank < 0) return cfail(TREPB_ERR_INVALID, "bytes_per_rank must be >= 0");
    if (bytes_per_rank == 0) return TREPB_OK;
    Nccl& n = nccl();
    CUC(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nb = (size_t)bytes_per_rank;
    NC(n.GroupStart());
    if (c->rank == root) {
        for (int r = 0; r < c->nranks; ++r) {
            if (r == root) continue;
            ncclResult_t rr = n.Recv((char*)recv + (size_t)r * nb, nb, ncclUint8, r, c->comm, st);
            if (rr != ncclSuccess) { n.GroupEnd(); return nccl_fail(rr, "ncclRecv"); }
        }
    } else {
        ncclResult_t rr = n.Send(send, nb, ncclUint8, root, c->comm, st);
        if (rr != ncclSuccess) { n.GroupEnd(); return nccl_fail(rr, "ncclSend"); }
    }
    NC(n.GroupEnd());
    if (c->rank == root && (const char*)send != (char*)recv + (size_t)root * nb)
        CUC(cudaMemcpyAsync((char*)recv + (size_t)root * nb, send, nb, cudaMemcpyDeviceToDevice, st));
    return TREPB_OK;
}

// ---- peer-mapped slabs (CUDA IPC) ---------------------------------------------------------------
int trepb_ipc_export(int device, const void* ptr, char* handle) {
    TREPB_NVTX("trepb_ipc_export");
    static_assert(sizeof(cudaIpcMemHandle_t) == TREPB_IPC_HANDLE_BYTES, "TREPB_IPC_HANDLE_BYTES must be sizeof(cudaIpcMemHandle_t)");
    if (!ptr || !handle) return cfail(TREPB_ERR_INVALID, "null argument");
    CUC(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    CUC(cudaIpcGetMemHandle(&h, (void*)ptr));
    memcpy(handle, &h, sizeof(h));
    return TREPB_OK;
}

int trepb_ipc_open(int device, const char* handle, void** ptr) {
    TREPB_NVTX("trepb_ipc_open");
    if (!handle || !ptr) return cfail(TREPB_ERR_INVALID, "null argument");
    CUC(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    CUC(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return TREPB_OK;
}

int trepb_ipc_close(int device, void* ptr) {
    TREPB_NVTX("trepb_ipc_close");
    if (!ptr) return TREPB_OK;
    CUC(cudaSetDevice(device));
    CUC(cudaIpcCloseMemHandle(ptr));
    return TREPB_OK;
}

// Event-timed helpers for callers without a CUDA binding of their own (bench.py): a pair of events on `stream`.
int trepb_event_pair_create(int device, void** pair) {
    if (!pair) return cfail(TREPB_ERR_INVALID, "null argument");
    CUC(cudaSetDevice(device));
    cudaEvent_t* e = new cudaEvent_t[2];
    cudaError_t r = cudaEventCreate(&e[0]);
    if (r == cudaSuccess) r = cudaEventCreate(&e[1]);
    if (r != cudaSuccess) { delete[] e; return cfail(TREPB_ERR_CUDA, cudaGetErrorString(r)); }
    *pair = e;
    return TREPB_OK;
}
int trepb_event_record(void* pair, int which, void* stream) {
    if (!pair || which < 0 || which > 1) return cfail(TREPB_ERR_INVALID, "bad arguments");
    CUC(cudaEventRecord(((cudaEvent_t*)pair)[which], (cudaStream_t)stream));
    return TREPB_OK;
}
int trepb_event_elapsed_ms(void* pair, float* ms) {
    if (!pair || !ms) return cfail(TREPB_ERR_INVALID, "null argument");
    cudaEvent_t* e = (cudaEvent_t*)pair;
    CUC(cudaEventSynchronize(e[1]));
    CUC(cudaEventElapsedTime(ms, e[0], e[1]));
    return TREPB_OK;
}
void trepb_event_pair_destroy(void* pair) {
    if (!pair) return;
    cudaEvent_t* e = (cudaEvent_t*)pair;
    cudaEventDestroy(e[0]);
    cudaEventDestroy(e[1]);
    delete[] e;
}

}  // extern "C"
