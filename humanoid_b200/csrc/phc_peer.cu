// RunningNorm.update across GPUs in ONE kernel: all-reduce over NVLink peer memory + the blend.
//
// The only exchange on the path: RunningNorm.update (policies/running_norm.py:23-34)
// over the concatenated batch of all ranks needs the sum over ranks of [sum x (C) | sum x^2 (C) | rows]
// in fp64 — 14 952 B for C = 934 — and then applies the reference's 1/count blend.  With NCCL that is
// an all-reduce launch (10-20 us of latency for 15 KB) between two small kernels.  Here every rank
// owns a mailbox in device memory that its peers map through CUDA IPC (NVLink / NVSwitch P2P); one
// launch per rank
//     1. copies the rank's partials into its own mailbox slot and publishes an epoch flag
//        (st.release.sys after __threadfence_system),
//     2. waits (ld.acquire.sys, bounded) until every rank's flag shows this epoch,
//     3. reads all W payloads straight out of the peers' memory and adds them IN RANK ORDER — every
//        rank computes the identical fp64 sums, so the statistics are bit-identical on all ranks,
//     4. applies mean / var(unbiased=False) and the 1/count blend to its fp32 running buffers,
//        bumps count, zeroes the caller's partials for the next rollout and advances the epoch.
// Two slots alternate by epoch parity: a rank can only reach epoch e+2 (and overwrite slot e&1) after
// every peer has published e+1, which each does (stream order) after its epoch-e kernel finished
// reading.  The epoch lives in device memory, so the launch is CUDA-graph capturable.  A peer that
// never shows up makes the wait time out: ONE block decides for the whole launch, so on a timeout the kernel
// records PHC_PEER_TIMEOUT in the context's status word and touches nothing — not the running buffers, not the
// caller's partials, not count / epoch — instead of hanging the GPU.  A timed-out rank has published its flag for
// an epoch its peers may or may not complete, so the group is out of step afterwards: every rank calls
// phc_peer_reduce_resync (between two host barriers) before the next update.
#include <cstdint>
#include <cstring>
#include <new>

#include <cuda_runtime.h>

#include "../../include/phc_b200.h"

namespace phc {
int record_cuda_error(int cuda_error);  // phc_kernels.cu
}

namespace {

constexpr int kThreads = 256;

// mailbox layout (doubles): slot s at s * slot_stride: [payload 2C+1 | pad]; flags after both slots
struct PeerParams {
  double* peers[PHC_PEER_MAX_WORLD];  // every rank's mailbox (own included), device-visible addresses
  int world, rank;
  int64_t cols;         // C
  int64_t slot_stride;  // doubles
  double* sums;         // caller's partials [2C], zeroed at the end
  const double* rows_dev;  // device scalar or null
  double rows;             // used when rows_dev is null
  float *mean, *var, *count;
  unsigned long long* epoch;  // device: number of completed reductions
  unsigned int* ticket;       // device: [0] publish ticket, [1] finish ticket
  int* status;                // device: 0 ok, PHC_PEER_TIMEOUT
  unsigned long long* decision;  // device: (launch sequence number << 1) | ok, written by block 0
  unsigned long long* seq;       // device: launches completed (timed-out ones included)
  unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned long long* flag_ptr(double* mailbox, int64_t slot_stride, int slot) {
  return reinterpret_cast<unsigned long long*>(mailbox + 2 * slot_stride) + slot;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) { st_release_sys(p, v); }

__global__ void __launch_bounds__(kThreads) peer_reduce_update_kernel(const PeerParams p) {
  __shared__ int s_ok;
  const int64_t C = p.cols, L = 2 * C + 1;
  // *p.epoch and *p.seq are only advanced by the last block of a launch, at its very end: every block reads the same
  const unsigned long long e = *p.epoch + 1;
  const unsigned long long seq = *p.seq + 1;  // launches so far + 1, timed-out ones included
  const int slot = (int)(e & 1);
  double* mine = p.peers[p.rank] + slot * p.slot_stride;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;

  // 1. partials -> own mailbox slot; the last block to finish publishes the flag
  for (int64_t i = i0; i < L; i += stride) mine[i] = i < 2 * C ? p.sums[i] : (p.rows_dev ? *p.rows_dev : p.rows);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(&p.ticket[0], 1u) == gridDim.x - 1) {
      p.ticket[0] = 0;
      __threadfence_system();
      st_release_sys(flag_ptr(p.peers[p.rank], p.slot_stride, slot), e);
    }
    // 2. ONE decision per launch: block 0 waits for every rank's flag of this epoch (own included: other blocks of
    // this grid wrote part of it) and publishes ok / timed out; the other blocks take its verdict, so either every
    // block blends and clears its columns or none does.  All blocks are co-resident (<= 32 blocks).
    int ok = 1;
    if (blockIdx.x == 0) {
      const unsigned long long t0 = global_ns();
      for (int r = 0; r < p.world && ok; ++r) {
        const unsigned long long* f = flag_ptr(p.peers[r], p.slot_stride, slot);
        while (ld_acquire_sys(f) != e) {
          if (global_ns() - t0 > p.timeout_ns) {
            ok = 0;
            break;
          }
          __nanosleep(200);
        }
      }
      st_release_sys_u64(p.decision, seq * 2 + (unsigned long long)ok);
    } else {
      unsigned long long d;
      while (((d = ld_acquire_sys(p.decision)) >> 1) != seq) __nanosleep(100);
      ok = (int)(d & 1ull);
    }
    s_ok = ok;
  }
  __syncthreads();
  const bool ok = s_ok != 0;

  // 3 + 4. sum in rank order, blend (policies/running_norm.py:23-34).  On a timeout nothing is touched: not the
  // running buffers, not the caller's partials, not count / epoch.
  if (ok) {
    double n = 0.0;
    for (int r = 0; r < p.world; ++r) n += ld_sys(p.peers[r] + slot * p.slot_stride + 2 * C);
    const float cnt = *p.count;
    for (int64_t c = i0; c < C; c += stride) {
      double s1 = 0.0, s2 = 0.0;
      for (int r = 0; r < p.world; ++r) {
        const double* q = p.peers[r] + slot * p.slot_stride;
        s1 += ld_sys(q + c);
        s2 += ld_sys(q + C + c);
      }
      if (n > 0) {
        const double m = s1 / n;
        double v = s2 / n - m * m;  // var(unbiased=False)
        if (v < 0) v = 0;
        const float w = 1.0f / cnt;  // weight = 1 / self.count
        p.mean[c] = p.mean[c] * (1.0f - w) + (float)m * w;
        p.var[c] = p.var[c] * (1.0f - w) + (float)v * w;
      }
    }
    for (int64_t i = i0; i < 2 * C; i += stride) p.sums[i] = 0.0;  // next rollout accumulates from zero
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&p.ticket[1], 1u) == gridDim.x - 1) {  // every block has read count, the epoch and the sequence number
      p.ticket[1] = 0;
      *p.seq = seq;
      if (ok) {
        *p.count = *p.count + 1.0f;
        *p.epoch = e;
      } else {
        *p.status = PHC_PEER_TIMEOUT;
      }
    }
  }
}

}  // namespace

struct PhcPeerReduce {
  int rank = 0, world = 1;
  int64_t cols = 0, slot_stride = 0;
  double* mailbox = nullptr;  // own, cudaMalloc
  size_t mailbox_bytes = 0;
  double* peers[PHC_PEER_MAX_WORLD] = {};
  bool opened[PHC_PEER_MAX_WORLD] = {};  // mapped through cudaIpcOpenMemHandle (to be closed)
  unsigned long long* epoch = nullptr;   // device scratch: epoch | tickets | status
  unsigned long long timeout_ns = 5000000000ull;
  bool connected = false;
};

extern "C" {

int phc_peer_reduce_create(int32_t rank, int32_t world, int64_t cols, int64_t timeout_ms, PhcPeerReduce** out) {
  if (!out) return PHC_ERR_NULL;
  if (world < 1 || world > PHC_PEER_MAX_WORLD || rank < 0 || rank >= world || cols < 1 || cols > (1 << 24))
    return PHC_ERR_SHAPE;
  PhcPeerReduce* c = new (std::nothrow) PhcPeerReduce;
  if (!c) return PHC_ERR_ALLOC;
  c->rank = rank;
  c->world = world;
  c->cols = cols;
  c->slot_stride = (2 * cols + 1 + 15) / 16 * 16;  // 128-B aligned slots
  c->mailbox_bytes = (size_t)(2 * c->slot_stride + 16) * sizeof(double);
  if (timeout_ms > 0) c->timeout_ns = (unsigned long long)timeout_ms * 1000000ull;
  cudaError_t e = cudaMalloc((void**)&c->mailbox, c->mailbox_bytes);
  if (e == cudaSuccess) e = cudaMemset(c->mailbox, 0, c->mailbox_bytes);
  if (e == cudaSuccess) e = cudaMalloc((void**)&c->epoch, 64);
  if (e == cudaSuccess) e = cudaMemset(c->epoch, 0, 64);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    if (c->mailbox) cudaFree(c->mailbox);
    if (c->epoch) cudaFree(c->epoch);
    delete c;
    return e == cudaErrorMemoryAllocation ? PHC_ERR_ALLOC : phc::record_cuda_error((int)e);
  }
  c->peers[rank] = c->mailbox;
  c->connected = world == 1;
  *out = c;
  return PHC_OK;
}

int phc_peer_reduce_handle(const PhcPeerReduce* c, void* handle_out) {
  if (!c || !handle_out) return PHC_ERR_NULL;
  static_assert(sizeof(cudaIpcMemHandle_t) == PHC_PEER_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, c->mailbox);
  if (e != cudaSuccess) return phc::record_cuda_error((int)e);
  memcpy(handle_out, &h, sizeof(h));
  return PHC_OK;
}

int phc_peer_reduce_connect(PhcPeerReduce* c, const void* handles) {
  if (!c || !handles) return PHC_ERR_NULL;
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank || c->peers[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const unsigned char*)handles + (size_t)r * PHC_PEER_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return phc::record_cuda_error((int)e);
    c->peers[r] = (double*)p;
    c->opened[r] = true;
  }
  c->connected = true;
  return PHC_OK;
}

int phc_peer_reduce_connect_local(PhcPeerReduce* c, PhcPeerReduce* const* all) {
  if (!c || !all) return PHC_ERR_NULL;
  for (int r = 0; r < c->world; ++r) {
    if (!all[r] || all[r]->cols != c->cols || all[r]->world != c->world || all[r]->rank != r) return PHC_ERR_SHAPE;
    c->peers[r] = all[r]->mailbox;
  }
  c->connected = true;
  return PHC_OK;
}

int phc_running_norm_update_peers(PhcPeerReduce* c, double* sums, const double* rows_dev, double rows,
                                  float* running_mean, float* running_var, float* count, phc_stream_t stream) {
  if (!c || !sums || !running_mean || !running_var || !count) return PHC_ERR_NULL;
  if (!c->connected) return PHC_ERR_UNSUPPORTED;
  PeerParams p{};
  for (int r = 0; r < c->world; ++r) p.peers[r] = c->peers[r];
  p.world = c->world;
  p.rank = c->rank;
  p.cols = c->cols;
  p.slot_stride = c->slot_stride;
  p.sums = sums;
  p.rows_dev = rows_dev;
  p.rows = rows;
  p.mean = running_mean;
  p.var = running_var;
  p.count = count;
  p.epoch = c->epoch;
  p.ticket = reinterpret_cast<unsigned int*>(c->epoch + 1);
  p.status = reinterpret_cast<int*>(c->epoch + 3);
  p.decision = c->epoch + 4;
  p.seq = c->epoch + 5;
  p.timeout_ns = c->timeout_ns;
  const int64_t L = 2 * c->cols + 1;
  int blocks = (int)((L + kThreads - 1) / kThreads);
  if (blocks > 32) blocks = 32;  // all blocks must be co-resident: one of them publishes, all of them wait
  peer_reduce_update_kernel<<<blocks, kThreads, 0, stream>>>(p);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? PHC_OK : phc::record_cuda_error((int)e);
}

int phc_peer_reduce_status(PhcPeerReduce* c, int64_t* epoch_out) {
  if (!c) return PHC_ERR_NULL;
  unsigned long long host[4] = {};
  const cudaError_t e = cudaMemcpy(host, c->epoch, sizeof(host), cudaMemcpyDeviceToHost);  // synchronises
  if (e != cudaSuccess) return phc::record_cuda_error((int)e);
  if (epoch_out) *epoch_out = (int64_t)host[0];
  int status;
  memcpy(&status, &host[3], sizeof(status));
  if (status != 0) {  // sticky until read
    (void)cudaMemset(reinterpret_cast<int*>(c->epoch + 3), 0, sizeof(int));
    return status;
  }
  return PHC_OK;
}

int phc_peer_reduce_resync(PhcPeerReduce* c) {
  if (!c) return PHC_ERR_NULL;
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemset(c->mailbox, 0, c->mailbox_bytes);  // both slots and both epoch flags
  if (e == cudaSuccess) e = cudaMemset(c->epoch, 0, 64);                  // epoch, tickets, status, decision, sequence
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  return e == cudaSuccess ? PHC_OK : phc::record_cuda_error((int)e);
}

void phc_peer_reduce_destroy(PhcPeerReduce* c) {
  if (!c) return;
  for (int r = 0; r < c->world; ++r)
    if (c->opened[r] && c->peers[r]) cudaIpcCloseMemHandle(c->peers[r]);
  if (c->mailbox) cudaFree(c->mailbox);
  if (c->epoch) cudaFree(c->epoch);
  delete c;
}

}  // extern "C"
