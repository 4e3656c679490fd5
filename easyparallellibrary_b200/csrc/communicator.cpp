// Native NCCL communicator (C ABI, loaded with ctypes).
//
// B200 counterpart of the reference's native layer (csrc/communicators/*: 12 TF ops, each an
// AsyncOpKernel owning one extra CUDA stream and one worker thread, tensorflow_cuda.h:50-136).
// Here a communicator is a plain object: one ncclComm_t + one dedicated high-priority side
// stream + two events.  Every verb is asynchronous and stream-ordered:
//
//     side stream waits on an event recorded on the caller's stream  ->  NCCL call on the side
//     stream  ->  event recorded on the side stream; epl_comm_wait() makes the caller's stream
//     (not the host) wait for it.
//
// No host thread, no host synchronisation — including the variable-length collectives, whose
// counts are exchanged on the device and handed back as a device tensor (the reference blocks
// the host there: nccl_all_gather.cc:150-204, nccl_all_to_all.cc:134-200).
//
// NCCL is resolved at run time from the libnccl that PyTorch ships (dlopen), so the wrapper
// and torch.distributed always agree on the NCCL version.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

extern "C" {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5,
               ncclFloat16 = 6, ncclFloat32 = 7, ncclFloat64 = 8, ncclBfloat16 = 9 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3, ncclAvg = 4 } ncclRedOp_t;
}

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi g_api;
std::mutex g_mu;
std::string g_error;

template <typename F> bool sym(void* h, const char* name, F& out) {
  out = reinterpret_cast<F>(dlsym(h, name));
  return out != nullptr;
}

struct Comm {
  ncclComm_t comm = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ready = nullptr, done = nullptr;
  int rank = 0, size = 1, device = 0;
  void* counts_dev = nullptr;      // size*size int64 scratch for variable-length verbs
};

size_t dtype_size(int dt) {
  switch (dt) {
    case ncclInt8: case ncclUint8: return 1;
    case ncclFloat16: case ncclBfloat16: return 2;
    case ncclInt32: case ncclUint32: case ncclFloat32: return 4;
    default: return 8;
  }
}

int fail(const char* what, int code) {
  std::lock_guard<std::mutex> l(g_mu);
  g_error = std::string(what) + " failed with code " + std::to_string(code);
  if (g_api.GetErrorString && code > 0) g_error += std::string(": ") + g_api.GetErrorString((ncclResult_t)code);
  return code ? code : -1;
}

#define NCCL_TRY(expr) do { ncclResult_t r_ = (expr); if (r_ != ncclSuccess) return fail(#expr, (int)r_); } while (0)
#define CUDA_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return fail(#expr, 1000 + (int)e_); } while (0)

// caller stream -> side stream
int begin(Comm* c, void* caller_stream) {
  CUDA_TRY(cudaEventRecord(c->ready, static_cast<cudaStream_t>(caller_stream)));
  CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ready, 0));
  return 0;
}
int end(Comm* c) {
  CUDA_TRY(cudaEventRecord(c->done, c->stream));
  return 0;
}

}  // namespace

extern "C" {

const char* epl_comm_last_error() { return g_error.c_str(); }

int epl_nccl_load(const char* path) {
  std::lock_guard<std::mutex> l(g_mu);
  if (g_api.handle) return 0;
  void* h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) { g_error = std::string("dlopen failed: ") + dlerror(); return -1; }
  bool ok = sym(h, "ncclGetUniqueId", g_api.GetUniqueId) && sym(h, "ncclCommInitRank", g_api.CommInitRank) &&
            sym(h, "ncclCommDestroy", g_api.CommDestroy) && sym(h, "ncclCommAbort", g_api.CommAbort) &&
            sym(h, "ncclAllReduce", g_api.AllReduce) && sym(h, "ncclReduce", g_api.Reduce) &&
            sym(h, "ncclBroadcast", g_api.Broadcast) && sym(h, "ncclAllGather", g_api.AllGather) &&
            sym(h, "ncclReduceScatter", g_api.ReduceScatter) && sym(h, "ncclSend", g_api.Send) &&
            sym(h, "ncclRecv", g_api.Recv) && sym(h, "ncclGroupStart", g_api.GroupStart) &&
            sym(h, "ncclGroupEnd", g_api.GroupEnd) && sym(h, "ncclGetErrorString", g_api.GetErrorString) &&
            sym(h, "ncclGetVersion", g_api.GetVersion);
  if (!ok) { g_error = "libnccl is missing required symbols"; dlclose(h); return -2; }
  g_api.handle = h;
  return 0;
}

int epl_nccl_version() { int v = 0; if (g_api.GetVersion) g_api.GetVersion(&v); return v; }

// id: 128 bytes, produced on group rank 0 and distributed by the control plane (TCPStore)
int epl_comm_get_unique_id(void* id128) {
  if (!g_api.handle) return fail("nccl not loaded", -3);
  NCCL_TRY(g_api.GetUniqueId(static_cast<ncclUniqueId*>(id128)));
  return 0;
}

int epl_comm_create(const void* id128, int size, int rank, int device, void** out) {
  if (!g_api.handle) return fail("nccl not loaded", -3);
  Comm* c = new Comm();
  c->rank = rank; c->size = size; c->device = device;
  CUDA_TRY(cudaSetDevice(device));
  int lo = 0, hi = 0;
  CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CUDA_TRY(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, hi));
  CUDA_TRY(cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming));
  CUDA_TRY(cudaMalloc(&c->counts_dev, sizeof(int64_t) * size * size));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NCCL_TRY(g_api.CommInitRank(&c->comm, size, id, rank));
  *out = c;
  return 0;
}

int epl_comm_destroy(void* h) {
  Comm* c = static_cast<Comm*>(h);
  if (!c) return 0;
  if (c->comm) g_api.CommDestroy(c->comm);
  if (c->counts_dev) cudaFree(c->counts_dev);
  if (c->ready) cudaEventDestroy(c->ready);
  if (c->done) cudaEventDestroy(c->done);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

int epl_comm_abort(void* h) {
  Comm* c = static_cast<Comm*>(h);
  if (c && c->comm) { g_api.CommAbort(c->comm); c->comm = nullptr; }
  return 0;
}

void* epl_comm_stream(void* h) { return static_cast<Comm*>(h)->stream; }

// make `caller_stream` wait (on the device) for everything issued on this communicator so far
int epl_comm_wait(void* h, void* caller_stream) {
  Comm* c = static_cast<Comm*>(h);
  CUDA_TRY(cudaStreamWaitEvent(static_cast<cudaStream_t>(caller_stream), c->done, 0));
  return 0;
}

int epl_comm_all_reduce(void* h, const void* send, void* recv, int64_t count, int dtype, int op, void* stream) {
  Comm* c = static_cast<Comm*>(h);
  if (int rc = begin(c, stream)) return rc;
  NCCL_TRY(g_api.AllReduce(send, recv, (size_t)count, (ncclDataType_t)dtype, (ncclRedOp_t)op, c->comm, c->stream));
  return end(c);
}

int epl_comm_reduce(void* h, const void* send, void* recv, int64_t count, int dtype, int op, int root, void* stream) {
  Comm* c = static_cast<Comm*>(h);
  if (int rc = begin(c, stream)) return rc;
  NCCL_TRY(g_api.Reduce(send, recv, (size_t)count, (ncclDataType_t)dtype, (ncclRedOp_t)op, root, c->comm, c->stream));
  return end(c);
}

int epl_comm_broadcast(void* h, const void* send, void* recv, int64_t count, int dtype, int root, void* stream) {
  Comm* c = static_cast<Comm*>(h);
  if (int rc = begin(c, stream)) return rc;
  NCCL_TRY(g_api.Broadcast(send, recv, (size_t)count, (ncclDataType_t)dtype, root, c->comm, c->stream));
  return end(c);
}

int epl_comm_all_gather(void* h, const void* send, void* recv, int64_t send_count, int dtype, void* stream) {
  Comm* c = static_cast<Comm*>(h);
  if (int rc = begin(c, stream)) return rc;
  NCCL_TRY(g_api.AllGather(send, recv, (size_t)send_count, (ncclDataType_t)dtype, c->comm, c->stream));
  return end(c);
}

int epl_comm_reduce_scatter(void* h, const void* send, void* recv, int64_t recv_count, int dtype, int op, void* stream) {
  Comm* c = static_cast<Comm*>(h);
  if (int rc = begin(c, stream)) return rc;
  NCCL_TRY(g_api.ReduceScatter(send, recv, (size_t)recv_count, (ncclDataType_t)dtype, (ncclRedOp_t)op, c->comm, c->stream));
  return end(c);
}

// equal segments: segment j of `send` goes to rank j
int epl_comm_all_to_all(void* h, const void* send, void* recv, int64_t count_per_rank, int dtype, void* stream) {
  Comm* c = static_cast<Comm*>(h);
  if (int rc = begin(c, stream)) return rc;
  const size_t bytes = (size_t)count_per_rank * dtype_size(dtype);
  NCCL_TRY(g_api.GroupStart());
  for (int r = 0; r < c->size; ++r) {
    NCCL_TRY(g_api.Send(static_cast<const char*>(send) + r * bytes, (size_t)count_per_rank, (ncclDataType_t)dtype, r, c->comm, c->stream));
    NCCL_TRY(g_api.Recv(static_cast<char*>(recv) + r * bytes, (size_t)count_per_rank, (ncclDataType_t)dtype, r, c->comm, c->stream));
  }
  NCCL_TRY(g_api.GroupEnd());
  return end(c);
}

// variable segments with host-known element counts/offsets (element units); counts for both directions supplied
int epl_comm_all_to_allv(void* h, const void* send, const int64_t* send_counts, const int64_t* send_offsets, void* recv,
                         const int64_t* recv_counts, const int64_t* recv_offsets, int dtype, void* stream) {
  Comm* c = static_cast<Comm*>(h);
  if (int rc = begin(c, stream)) return rc;
  const size_t es = dtype_size(dtype);
  NCCL_TRY(g_api.GroupStart());
  for (int r = 0; r < c->size; ++r) {
    if (send_counts[r] > 0)
      NCCL_TRY(g_api.Send(static_cast<const char*>(send) + send_offsets[r] * es, (size_t)send_counts[r], (ncclDataType_t)dtype, r, c->comm, c->stream));
    if (recv_counts[r] > 0)
      NCCL_TRY(g_api.Recv(static_cast<char*>(recv) + recv_offsets[r] * es, (size_t)recv_counts[r], (ncclDataType_t)dtype, r, c->comm, c->stream));
  }
  NCCL_TRY(g_api.GroupEnd());
  return end(c);
}

// all-gather of rows whose count differs per rank, into a padded [size, max_rows, row_elems] buffer; the
// per-rank counts are gathered into `counts_out` (device int64[size]) by the same stream — no host sync.
int epl_comm_all_gatherv_padded(void* h, const void* send, const void* my_count_dev, void* recv, void* counts_out,
                                int64_t max_rows, int64_t row_elems, int dtype, void* stream) {
  Comm* c = static_cast<Comm*>(h);
  if (int rc = begin(c, stream)) return rc;
  NCCL_TRY(g_api.GroupStart());
  NCCL_TRY(g_api.AllGather(my_count_dev, counts_out, 1, ncclInt64, c->comm, c->stream));
  NCCL_TRY(g_api.AllGather(send, recv, (size_t)(max_rows * row_elems), (ncclDataType_t)dtype, c->comm, c->stream));
  NCCL_TRY(g_api.GroupEnd());
  return end(c);
}

int epl_comm_send(void* h, const void* buf, int64_t count, int dtype, int peer, void* stream) {
  Comm* c = static_cast<Comm*>(h);
  if (int rc = begin(c, stream)) return rc;
  NCCL_TRY(g_api.Send(buf, (size_t)count, (ncclDataType_t)dtype, peer, c->comm, c->stream));
  return end(c);
}

int epl_comm_recv(void* h, void* buf, int64_t count, int dtype, int peer, void* stream) {
  Comm* c = static_cast<Comm*>(h);
  if (int rc = begin(c, stream)) return rc;
  NCCL_TRY(g_api.Recv(buf, (size_t)count, (ncclDataType_t)dtype, peer, c->comm, c->stream));
  return end(c);
}

// batched point-to-point: n operations {is_send, buf, count, dtype, peer} issued as one NCCL group
struct EplP2POp { int is_send; int dtype; int peer; int pad; void* buf; int64_t count; };
int epl_comm_batch_p2p(void* h, const EplP2POp* ops, int n, void* stream) {
  Comm* c = static_cast<Comm*>(h);
  if (int rc = begin(c, stream)) return rc;
  NCCL_TRY(g_api.GroupStart());
  for (int i = 0; i < n; ++i) {
    if (ops[i].is_send) NCCL_TRY(g_api.Send(ops[i].buf, (size_t)ops[i].count, (ncclDataType_t)ops[i].dtype, ops[i].peer, c->comm, c->stream));
    else NCCL_TRY(g_api.Recv(ops[i].buf, (size_t)ops[i].count, (ncclDataType_t)ops[i].dtype, ops[i].peer, c->comm, c->stream));
  }
  NCCL_TRY(g_api.GroupEnd());
  return end(c);
}

}  // extern "C"
