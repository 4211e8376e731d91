// cq_comm.cu -- the multi-GPU side of the C-ABI (include/cq_b200.h): the stash exchange of A.assemble()
// (reference README.md:95, demo/poisson.py:194: PETSc MatAssemblyBegin/End ships the contributions to rows another
// process owns) as NCCL send/recv between slab neighbours, and a neighbour fence through peer memory for the fused
// peer scatter.  One process per GPU; the communicator lives in the cq_ctx.
//
// NCCL is resolved at run time: dlopen of the libnccl.so.2 already mapped into the process (torch's), else the
// system's.  The library keeps linking against libcudart only.
#include <dlfcn.h>

#include <cstring>

#include "cq_internal.h"

namespace cq {

namespace {

// the part of nccl.h that is used (NCCL 2.x ABI: ncclUniqueId is 128 bytes passed by value)
struct NcclUniqueId { char internal[CQ_COMM_ID_BYTES]; };
using NcclComm = void*;
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;

struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api = [] {
    NcclApi a;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the process already uses (torch)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return a;
    auto sym = [&](const char* n) { return dlsym(h, n); };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
    a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.Send && a.Recv && a.AllReduce && a.GroupStart && a.GroupEnd && a.GetErrorString;
    return a;
  }();
  return api;
}

int check_nccl(int r, const char* what) {
  if (r == 0) return CQ_OK;
  set_error("%s: NCCL error %d (%s)", what, r, nccl().GetErrorString ? nccl().GetErrorString(r) : "?");
  return CQ_ERR_CUDA;
}

int need_comm(cq_ctx* ctx) {
  if (!ctx) { set_error("null ctx"); return CQ_ERR_INVALID; }
  if (!ctx->nccl_comm) { set_error("no communicator: call cq_comm_init first"); return CQ_ERR_INVALID; }
  return check_cuda(cudaSetDevice(ctx->device), "cudaSetDevice");
}

int ensure_recv(cq_ctx* ctx, size_t bytes) {
  if (ctx->recv_bytes >= bytes) return CQ_OK;
  if (ctx->recv_buf) cudaFree(ctx->recv_buf);
  ctx->recv_buf = nullptr;
  ctx->recv_bytes = 0;
  if (int rc = check_cuda(cudaMalloc(&ctx->recv_buf, bytes), "cudaMalloc(recv buffer)")) return rc;
  ctx->recv_bytes = bytes;
  return CQ_OK;
}

// One thread.  Everything enqueued on the stream before this kernel (the zeroing of this rank's matrix, or its
// assembly kernel with the REDs into the neighbour's memory) has completed at kernel boundaries; the release store
// publishes it to the neighbour, the acquire load makes the neighbour's side visible to what follows on this stream.
__global__ void peer_fence_kernel(unsigned int* to_down, unsigned int* to_up, const unsigned int* from_down, const unsigned int* from_up,
                                  const unsigned int epoch) {
  __threadfence_system();
  if (to_down) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(to_down), "r"(epoch) : "memory");
  if (to_up) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(to_up), "r"(epoch) : "memory");
  const unsigned int* wait_on[2] = {to_down ? from_down : nullptr, to_up ? from_up : nullptr};
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (!wait_on[k]) continue;
    unsigned int v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(wait_on[k]) : "memory");
      if ((int)(v - epoch) < 0) __nanosleep(200);
    } while ((int)(v - epoch) < 0);  // wrap-safe: epochs of the two sides differ by at most one
  }
  __threadfence_system();
}

}  // namespace

void comm_release(cq_ctx* ctx) {
  if (ctx->nccl_comm && nccl().ok) nccl().CommDestroy(ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
  if (ctx->recv_buf) cudaFree(ctx->recv_buf);
  ctx->recv_buf = nullptr;
  ctx->recv_bytes = 0;
  if (ctx->fence_flags) cudaFree(ctx->fence_flags);
  ctx->fence_flags = nullptr;
}

}  // namespace cq

using namespace cq;

extern "C" {

int cq_comm_unique_id(unsigned char id[CQ_COMM_ID_BYTES]) {
  if (!id) { set_error("null id"); return CQ_ERR_INVALID; }
  if (!nccl().ok) { set_error("libnccl.so.2 not found (dlopen)"); return CQ_ERR_UNSUPPORTED; }
  NcclUniqueId u;
  if (int rc = check_nccl(nccl().GetUniqueId(&u), "ncclGetUniqueId")) return rc;
  memcpy(id, u.internal, CQ_COMM_ID_BYTES);
  return CQ_OK;
}

int cq_comm_init(cq_ctx* ctx, int nranks, int rank, const unsigned char id[CQ_COMM_ID_BYTES]) {
  if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) { set_error("bad argument"); return CQ_ERR_INVALID; }
  if (!nccl().ok) { set_error("libnccl.so.2 not found (dlopen)"); return CQ_ERR_UNSUPPORTED; }
  if (ctx->nccl_comm) { set_error("communicator already initialised"); return CQ_ERR_INVALID; }
  if (int rc = check_cuda(cudaSetDevice(ctx->device), "cudaSetDevice")) return rc;
  NcclUniqueId u;
  memcpy(u.internal, id, CQ_COMM_ID_BYTES);
  NcclComm comm = nullptr;
  if (int rc = check_nccl(nccl().CommInitRank(&comm, nranks, u, rank), "ncclCommInitRank")) return rc;
  ctx->nccl_comm = comm;
  ctx->comm_rank = rank;
  ctx->comm_size = nranks;
  return CQ_OK;
}

int cq_comm_destroy(cq_ctx* ctx) {
  if (!ctx) return CQ_OK;
  cudaSetDevice(ctx->device);
  if (ctx->nccl_comm && nccl().ok) nccl().CommDestroy(ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
  ctx->comm_rank = -1;
  ctx->comm_size = 0;
  return CQ_OK;
}

int cq_exchange_offrank(cq_ctx* ctx, const cq_csr* A, int up, int64_t send_lo, int64_t send_hi, int down, int64_t recv_row0,
                        int64_t recv_nrows, const int64_t* recv_indptr, int64_t recv_count, void* stream) {
  if (int rc = need_comm(ctx)) return rc;
  if (!A || !A->values || !A->indptr) { set_error("CSR arrays null"); return CQ_ERR_INVALID; }
  if (up >= ctx->comm_size || down >= ctx->comm_size) { set_error("neighbour rank out of range"); return CQ_ERR_INVALID; }
  if (up >= 0 && (send_lo < 0 || send_hi < send_lo || send_hi > A->nnz)) { set_error("send range out of bounds"); return CQ_ERR_INVALID; }
  if (down >= 0 && (recv_count < 0 || recv_nrows < 0 || recv_row0 < 0 || recv_row0 + recv_nrows > A->num_rows || (recv_nrows > 0 && !recv_indptr))) {
    set_error("receive description out of bounds");
    return CQ_ERR_INVALID;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool snd = up >= 0 && send_hi > send_lo, rcv = down >= 0 && recv_count > 0;
  if (rcv)
    if (int rc = ensure_recv(ctx, sizeof(double) * (size_t)recv_count)) return rc;
  if (snd || rcv) {
    if (int rc = check_nccl(nccl().GroupStart(), "ncclGroupStart")) return rc;
    int r1 = 0, r2 = 0;
    if (snd) r1 = nccl().Send(A->values + send_lo, (size_t)(send_hi - send_lo), NCCL_FLOAT64, up, ctx->nccl_comm, s);
    if (rcv) r2 = nccl().Recv(ctx->recv_buf, (size_t)recv_count, NCCL_FLOAT64, down, ctx->nccl_comm, s);
    const int r3 = nccl().GroupEnd();
    if (int rc = check_nccl(r1 ? r1 : (r2 ? r2 : r3), "ncclSend/ncclRecv")) return rc;
  }
  if (rcv) return cq_csr_accumulate_rows(ctx, A, recv_row0, recv_nrows, recv_indptr, static_cast<const double*>(ctx->recv_buf), stream);
  return CQ_OK;
}

int cq_exchange_offrank_vector(cq_ctx* ctx, double* b, int up, int64_t send_lo, int64_t send_hi, int down, int64_t recv_lo, int64_t recv_count,
                               void* stream) {
  if (int rc = need_comm(ctx)) return rc;
  if (!b) { set_error("b null"); return CQ_ERR_INVALID; }
  if (up >= ctx->comm_size || down >= ctx->comm_size || (up >= 0 && (send_lo < 0 || send_hi < send_lo)) || (down >= 0 && (recv_lo < 0 || recv_count < 0))) {
    set_error("bad exchange description");
    return CQ_ERR_INVALID;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool snd = up >= 0 && send_hi > send_lo, rcv = down >= 0 && recv_count > 0;
  if (rcv)
    if (int rc = ensure_recv(ctx, sizeof(double) * (size_t)recv_count)) return rc;
  if (snd || rcv) {
    if (int rc = check_nccl(nccl().GroupStart(), "ncclGroupStart")) return rc;
    int r1 = 0, r2 = 0;
    if (snd) r1 = nccl().Send(b + send_lo, (size_t)(send_hi - send_lo), NCCL_FLOAT64, up, ctx->nccl_comm, s);
    if (rcv) r2 = nccl().Recv(ctx->recv_buf, (size_t)recv_count, NCCL_FLOAT64, down, ctx->nccl_comm, s);
    const int r3 = nccl().GroupEnd();
    if (int rc = check_nccl(r1 ? r1 : (r2 ? r2 : r3), "ncclSend/ncclRecv")) return rc;
  }
  if (rcv) return cq_axpy(ctx, recv_count, static_cast<const double*>(ctx->recv_buf), b + recv_lo, stream);
  return CQ_OK;
}

int cq_allreduce_sum(cq_ctx* ctx, double* x, int64_t n, void* stream) {
  if (int rc = need_comm(ctx)) return rc;
  if (!x || n < 0) { set_error("bad argument"); return CQ_ERR_INVALID; }
  if (n == 0) return CQ_OK;
  return check_nccl(nccl().AllReduce(x, x, (size_t)n, NCCL_FLOAT64, NCCL_SUM, ctx->nccl_comm, static_cast<cudaStream_t>(stream)), "ncclAllReduce");
}

void* cq_peer_flags(cq_ctx* ctx) {
  if (!ctx) return nullptr;
  if (!ctx->fence_flags) {
    if (cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
    if (cudaMalloc(&ctx->fence_flags, 2 * sizeof(unsigned int)) != cudaSuccess) { ctx->fence_flags = nullptr; return nullptr; }
    cudaMemset(ctx->fence_flags, 0, 2 * sizeof(unsigned int));
    cudaDeviceSynchronize();
  }
  return ctx->fence_flags;
}

int cq_peer_fence_setup(cq_ctx* ctx, void* flag_in_down, void* flag_in_up) {
  if (!ctx) { set_error("null ctx"); return CQ_ERR_INVALID; }
  if (!cq_peer_flags(ctx)) { set_error("cannot allocate the fence words"); return CQ_ERR_CUDA; }
  ctx->fence_down = static_cast<unsigned int*>(flag_in_down);
  ctx->fence_up = static_cast<unsigned int*>(flag_in_up);
  ctx->fence_epoch = 0;
  return CQ_OK;
}

int cq_peer_fence(cq_ctx* ctx, void* stream) {
  if (!ctx || !ctx->fence_flags) { set_error("cq_peer_fence_setup has not been called"); return CQ_ERR_INVALID; }
  if (int rc = check_cuda(cudaSetDevice(ctx->device), "cudaSetDevice")) return rc;
  const unsigned int epoch = ++ctx->fence_epoch;
  peer_fence_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(ctx->fence_down, ctx->fence_up, ctx->fence_flags, ctx->fence_flags + 1, epoch);
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "peer_fence_kernel launch");
}

}  // extern "C"
