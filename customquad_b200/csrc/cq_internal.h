// cq_internal.h -- host-side internals shared by the translation units of libcq_b200.so
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <unordered_map>

#include "../../include/cq_b200.h"
#include "cq_kernels.cuh"

struct cq_ctx {
  int device = 0;
  int num_sms = 0;
  int variant = 0;
  int* err_flag = nullptr;  // device: [0] sparsity violation seen by a kernel, [1] scratch word of the setup kernels
  unsigned int* work_counter = nullptr;  // device: dynamic stage counter of the staged kernels
  int max_smem_optin = 0;
  std::atomic<int64_t> launches{0};
  // end-to-end (host buffer) path: grow-only device staging + a copy stream
  cudaStream_t copy_stream = nullptr;
  cudaStream_t d2h_stream = nullptr;  // result rows go back while later chunks are still coming in
  cudaEvent_t ev[12] = {};
  cudaEvent_t kev[16] = {};           // "kernel of chunk c done"
  int32_t* chunk_rowmin = nullptr;    // device [16]
  int32_t* host_rowmin = nullptr;     // pinned [16]
  int64_t* host_vallo = nullptr;      // pinned [16]
  void* stage = nullptr;
  size_t stage_bytes = 0;
  int* host_flag = nullptr;  // pinned
  double* scalar_dev = nullptr;  // device accumulator of the host-buffer scalar path
  // fused zeroing (cq_zero.cuh): per-wave value ranges + completion counters, grow-only
  int coop_launch = 0;  // device supports cooperative launches
  int fused_zero = 2;   // cq_ctx_set_fused_zero: 0 off, 1 on, 2 where measured to pay (cq_launch.cuh:fused_zero_ok)
  void* zero_buf = nullptr;
  size_t zero_waves = 0;
  // cut-cell producers: compacted list of the cells the box test calls cut (+ its length), grow-only
  int32_t* cut_list = nullptr;
  size_t cut_list_cap = 0;
  unsigned int* cut_count = nullptr;
  // multi-GPU (cq_comm.cu): NCCL communicator (resolved at run time), receive buffer, neighbour fence through peer memory
  void* nccl_comm = nullptr;
  int comm_rank = -1, comm_size = 0;
  void* recv_buf = nullptr;
  size_t recv_bytes = 0;
  unsigned int* fence_flags = nullptr;  // device [2]: [0] written by the down neighbour, [1] by the up neighbour
  unsigned int* fence_down = nullptr;   // mapped: the down neighbour's word [1]
  unsigned int* fence_up = nullptr;     // mapped: the up neighbour's word [0]
  unsigned int fence_epoch = 0;
  // owner-computes path (cq_owner.cuh): element blocks + dof -> entry lists, grow-only
  void* owner_ws = nullptr;
  size_t owner_ws_bytes = 0;
  // dynamic shared memory opt-in already granted on this device, per kernel (cq_launch.cuh:ensure_smem_optin)
  std::unordered_map<const void*, size_t> smem_optin;
};

namespace cq {

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

void comm_release(cq_ctx* ctx);  // cq_comm.cu: frees the communicator, the receive buffer and the fence words

// launchers implemented in cq_launch_*.cu; return cq_status
int launch_matrix(cq_ctx* ctx, const cq_form_desc& f, const KParams& P, cudaStream_t s);
int launch_vector(cq_ctx* ctx, const cq_form_desc& f, const KParams& P, cudaStream_t s);
int launch_scalar(cq_ctx* ctx, const cq_form_desc& f, const KParams& P, cudaStream_t s);

// grid size for a grid-strided kernel: enough blocks for the work, at most a full wave
template <auto Kernel>
int grid_for(cq_ctx* ctx, int threads, int64_t work_per_block_unit, int64_t units, size_t smem = 0) {
  static thread_local int blocks_per_sm = 0;  // one per kernel instantiation (an occupancy, the same on every B200)
  if (blocks_per_sm == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, Kernel, threads, smem) != cudaSuccess || blocks_per_sm < 1)
      blocks_per_sm = 1;
  }
  const int64_t need = (units + work_per_block_unit - 1) / work_per_block_unit;
  const int64_t full = (int64_t)ctx->num_sms * blocks_per_sm;
  return (int)(need < 1 ? 1 : (need < full ? need : full));
}

}  // namespace cq
