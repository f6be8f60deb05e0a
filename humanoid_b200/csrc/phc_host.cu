// Host-buffer pipeline around phc_step_fused (the end-to-end call of include/phc_b200.h).
//
// The env batch is cut into chunks; chunk c runs H2D(sim state + clock) -> fused step ->
// D2H(obs, reward, flags) on stream c % 3, so the host->device copy of one chunk, the kernel
// of another and the device->host copy of a third overlap (PCIe is full duplex; the B200 has
// separate copy engines per direction).  The call returns when every output is in host memory.
#include <cstdint>
#include <cstdlib>
#include <new>

#include <cuda_runtime.h>

#include "../../include/phc_b200.h"

namespace {

constexpr int kStreams = 3;
constexpr int kStateFloats = PHC_NUM_BODIES * 13;

struct DevBuf {
  float* state = nullptr;
  int16_t* progress = nullptr;
  float* start = nullptr;
  float* start_off = nullptr;
  float* goff = nullptr;
  int64_t* ids = nullptr;
  float* obs = nullptr;
  float* rew = nullptr;
  float* raw = nullptr;
  uint8_t* reset = nullptr;
  uint8_t* term = nullptr;
};

}  // namespace

struct PhcHostStep {
  const PhcLib* lib = nullptr;
  int64_t max_envs = 0;
  int32_t T = 1;
  int32_t chunks = 1;
  int64_t obs_dim = 0;
  float* term_dist = nullptr;
  uint32_t reset_mask = 0xFFFFFF;
  int32_t use_mean = 0, early = 1;
  float dt = 0.f;
  PhcRewardSpec rwd{};
  DevBuf d;
  cudaStream_t streams[kStreams] = {};
  int last_cuda = 0;
};

#define HOST_CUDA(ctx, call)            \
  do {                                  \
    cudaError_t e__ = (call);           \
    if (e__ != cudaSuccess) {           \
      if (ctx) (ctx)->last_cuda = e__;  \
      return PHC_ERR_CUDA;              \
    }                                   \
  } while (0)

extern "C" {

void phc_host_step_destroy(PhcHostStep* c) {
  if (!c) return;
  cudaFree(c->d.state);
  cudaFree(c->d.progress);
  cudaFree(c->d.start);
  cudaFree(c->d.start_off);
  cudaFree(c->d.goff);
  cudaFree(c->d.ids);
  cudaFree(c->d.obs);
  cudaFree(c->d.rew);
  cudaFree(c->d.raw);
  cudaFree(c->d.reset);
  cudaFree(c->d.term);
  cudaFree(c->term_dist);
  for (auto& s : c->streams)
    if (s) cudaStreamDestroy(s);
  delete c;
}

int phc_host_step_create(const PhcLib* lib, int64_t max_envs, int32_t time_steps, int32_t num_chunks,
                         const float* termination_distances_host, uint32_t reset_body_mask, int32_t use_mean,
                         int32_t enable_early_termination, float dt, const PhcRewardSpec* rwd, PhcHostStep** out) {
  if (!lib || !termination_distances_host || !rwd || !out) return PHC_ERR_NULL;
  if (max_envs <= 0 || time_steps < 1 || time_steps > PHC_MAX_TIME_STEPS || num_chunks < 1 || num_chunks > 1024)
    return PHC_ERR_SHAPE;
  PhcHostStep* c = new (std::nothrow) PhcHostStep;
  if (!c) return PHC_ERR_ALLOC;
  c->lib = lib;
  c->max_envs = max_envs;
  c->T = time_steps;
  c->chunks = num_chunks;
  c->obs_dim = PHC_SELF_OBS_DIM + (int64_t)PHC_TASK_OBS_DIM * time_steps;
  c->reset_mask = reset_body_mask;
  c->use_mean = use_mean;
  c->early = enable_early_termination;
  c->dt = dt;
  c->rwd = *rwd;
  const size_t n = (size_t)max_envs;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** p, size_t bytes) {
    if (e == cudaSuccess) e = cudaMalloc(p, bytes);
  };
  alloc((void**)&c->d.state, n * kStateFloats * sizeof(float));
  alloc((void**)&c->d.progress, n * sizeof(int16_t));
  alloc((void**)&c->d.start, n * sizeof(float));
  alloc((void**)&c->d.start_off, n * sizeof(float));
  alloc((void**)&c->d.goff, n * 3 * sizeof(float));
  alloc((void**)&c->d.ids, n * sizeof(int64_t));
  alloc((void**)&c->d.obs, n * c->obs_dim * sizeof(float));
  alloc((void**)&c->d.rew, n * sizeof(float));
  alloc((void**)&c->d.raw, n * 4 * sizeof(float));
  alloc((void**)&c->d.reset, n);
  alloc((void**)&c->d.term, n);
  alloc((void**)&c->term_dist, PHC_NUM_BODIES * sizeof(float));
  if (e == cudaSuccess)
    e = cudaMemcpy(c->term_dist, termination_distances_host, PHC_NUM_BODIES * sizeof(float), cudaMemcpyHostToDevice);
  for (int i = 0; i < kStreams && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    phc_host_step_destroy(c);
    return e == cudaErrorMemoryAllocation ? PHC_ERR_ALLOC : PHC_ERR_CUDA;
  }
  *out = c;
  return PHC_OK;
}

int64_t phc_host_step_h2d_bytes(const PhcHostStep* c, int64_t n) {
  (void)c;
  return n * (int64_t)(kStateFloats * sizeof(float) + sizeof(int16_t) + 2 * sizeof(float) + 3 * sizeof(float) +
                       sizeof(int64_t));
}
int64_t phc_host_step_d2h_bytes(const PhcHostStep* c, int64_t n) {
  return n * (int64_t)(c->obs_dim * sizeof(float) + sizeof(float) + 4 * sizeof(float) + 2 + sizeof(int16_t));
}

int phc_host_step(PhcHostStep* c, const PhcHostStepArgs* a, int64_t n) {
  if (!c || !a) return PHC_ERR_NULL;
  if (n == 0) return PHC_OK;
  if (n < 0 || n > c->max_envs) return PHC_ERR_SHAPE;
  if (!a->state || !a->progress_buf || !a->motion_start_times || !a->motion_start_times_offset ||
      !a->global_offset || !a->sampled_motion_ids || !a->obs_buf || !a->rew_buf || !a->reward_raw || !a->reset_buf ||
      !a->terminate_buf)
    return PHC_ERR_NULL;
  // chunk boundaries on multiples of 8 envs (one kernel block) so obs rows stay 16-B friendly
  int64_t per = (n + c->chunks - 1) / c->chunks;
  per = (per + 7) / 8 * 8;
  int ci = 0;
  for (int64_t lo = 0; lo < n; lo += per, ++ci) {
    const int64_t m = (n - lo) < per ? (n - lo) : per;
    cudaStream_t s = c->streams[ci % kStreams];
    const cudaMemcpyKind H2D = cudaMemcpyHostToDevice, D2H = cudaMemcpyDeviceToHost;
    HOST_CUDA(c, cudaMemcpyAsync(c->d.state + lo * kStateFloats, a->state + lo * kStateFloats,
                                 (size_t)m * kStateFloats * sizeof(float), H2D, s));
    HOST_CUDA(c, cudaMemcpyAsync(c->d.progress + lo, a->progress_buf + lo, (size_t)m * sizeof(int16_t), H2D, s));
    HOST_CUDA(c, cudaMemcpyAsync(c->d.start + lo, a->motion_start_times + lo, (size_t)m * sizeof(float), H2D, s));
    HOST_CUDA(c, cudaMemcpyAsync(c->d.start_off + lo, a->motion_start_times_offset + lo, (size_t)m * sizeof(float),
                                 H2D, s));
    HOST_CUDA(c, cudaMemcpyAsync(c->d.goff + lo * 3, a->global_offset + lo * 3, (size_t)m * 3 * sizeof(float), H2D, s));
    HOST_CUDA(c, cudaMemcpyAsync(c->d.ids + lo, a->sampled_motion_ids + lo, (size_t)m * sizeof(int64_t), H2D, s));

    PhcStepArgs k{};
    float* st = c->d.state + lo * kStateFloats;
    k.body.pos = PhcView{st, kStateFloats, 13};
    k.body.rot = PhcView{st + 3, kStateFloats, 13};
    k.body.vel = PhcView{st + 7, kStateFloats, 13};
    k.body.ang_vel = PhcView{st + 10, kStateFloats, 13};
    k.body.num_bodies = PHC_NUM_BODIES;
    k.progress_buf = c->d.progress + lo;
    k.motion_start_times = c->d.start + lo;
    k.motion_start_times_offset = c->d.start_off + lo;
    k.global_offset = c->d.goff + lo * 3;
    k.sampled_motion_ids = c->d.ids + lo;
    k.termination_distances = c->term_dist;
    k.reset_body_mask = c->reset_mask;
    k.use_mean = c->use_mean;
    k.enable_early_termination = c->early;
    k.advance_progress = 1;
    k.time_steps = c->T;
    k.dt = c->dt;
    k.rwd = c->rwd;
    k.obs_buf = c->d.obs + lo * c->obs_dim;
    k.obs_stride = c->obs_dim;
    k.rew_buf = c->d.rew + lo;
    k.reward_raw = c->d.raw + lo * 4;
    k.reward_raw_stride = 4;
    k.reset_buf = c->d.reset + lo;
    k.terminate_buf = c->d.term + lo;
    k.obs_moments = nullptr;
    int rc = phc_step_fused(c->lib, &k, m, s);
    if (rc) return rc;

    HOST_CUDA(c, cudaMemcpyAsync(a->obs_buf + lo * c->obs_dim, c->d.obs + lo * c->obs_dim,
                                 (size_t)m * c->obs_dim * sizeof(float), D2H, s));
    HOST_CUDA(c, cudaMemcpyAsync(a->rew_buf + lo, c->d.rew + lo, (size_t)m * sizeof(float), D2H, s));
    HOST_CUDA(c, cudaMemcpyAsync(a->reward_raw + lo * 4, c->d.raw + lo * 4, (size_t)m * 4 * sizeof(float), D2H, s));
    HOST_CUDA(c, cudaMemcpyAsync(a->reset_buf + lo, c->d.reset + lo, (size_t)m, D2H, s));
    HOST_CUDA(c, cudaMemcpyAsync(a->terminate_buf + lo, c->d.term + lo, (size_t)m, D2H, s));
    HOST_CUDA(c, cudaMemcpyAsync(a->progress_buf + lo, c->d.progress + lo, (size_t)m * sizeof(int16_t), D2H, s));
  }
  for (int i = 0; i < kStreams; ++i) HOST_CUDA(c, cudaStreamSynchronize(c->streams[i]));
  return PHC_OK;
}

}  // extern "C"
