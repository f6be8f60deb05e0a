// Host-buffer entry point around phc_step_fused (the end-to-end call of include/phc_b200.h).
//
// Direct path (all caller buffers pinned, hence mapped into the device address space by UVA): the
// sim state of each chunk is copied in by the host->device copy engine, and the chunk's fused
// kernel writes obs rows, rewards and flags straight into the caller's host buffers with its
// bulk stores (no device->host DMA descriptors, no device staging of outputs), so the two
// directions of the link overlap.  Used whenever cudaPointerGetAttributes says every buffer is
// pinned / managed host memory mapped at its own address (device pointers are rejected: PHC_ERR_UNSUPPORTED).
//
// Staged path (any pageable buffer): chunked copy pipeline, below.
//
// The env batch is cut into chunks; chunk c runs
//     H2D(sim state) + H2D(packed clock)  ->  fused step  ->  D2H(obs) + D2H(packed scalars)
// on stream c % 3, so the host->device copy of one chunk, the kernel of another and the
// device->host copy of a third overlap (PCIe is full duplex; the B200 has copy engines per
// direction).  The five small clock arrays and the five small outputs of a chunk travel as ONE
// transfer each through pinned staging buffers owned by the context (a DMA descriptor costs a
// few microseconds regardless of size; 22 of them per chunk were half of the step time).  The
// call returns when every output is in the caller's host memory.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include <cuda_runtime.h>

#include "../../include/phc_b200.h"

// phc_kernels.cu: phc_step_fused that also writes the advanced progress to a second address
int step_fused_mirrored(const PhcLib* lib, const PhcStepArgs* args, int64_t n, phc_stream_t stream, int16_t* progress_mirror);

namespace {

constexpr int kStreams = 3;
constexpr int kStateFloats = PHC_NUM_BODIES * 13;
// packed clock, per env: ids i64 | goff 3 f32 | start f32 | soff f32 | progress i16  (SoA per chunk)
constexpr size_t kClockBytes = 8 + 12 + 4 + 4 + 2;
// packed scalars out, per env: rew f32 | raw 4 f32 | progress i16 | reset u8 | term u8
constexpr size_t kOutBytes = 4 + 16 + 2 + 1 + 1;

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

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
  float* d_state = nullptr;
  float* d_obs = nullptr;
  unsigned char* d_clock = nullptr;  // packed clock, device
  unsigned char* d_out = nullptr;    // packed scalar outputs, device
  unsigned char* h_clock = nullptr;  // pinned staging
  unsigned char* h_out = nullptr;    // pinned staging
  size_t pack_capacity = 0;
  cudaStream_t streams[kStreams] = {};
  int last_cuda = 0;
  // direct path: verdict cached per set of caller pointers
  const void* seen[11] = {};
  int seen_direct = -1;
  // which output path pinned callers take: 0 = auto (time both over the first calls, keep the faster), 1 = direct
  // (kernels post into the mapped host buffers), 2 = staged (copy-engine D2H).  PHC_HOST_PATH=auto|direct|staged.
  int mode = 0;
  // the schedule (output path, chunk count) of pinned callers is tuned over the first calls: every candidate is timed
  // kTuneReps times, interleaved, and the fastest mean stays.  num_chunks > 0 at create and PHC_HOST_PATH narrow the
  // candidates (down to one: nothing to tune).
  struct Cand { int path, chunks; double t; int n; };  // t: fastest timed call so far
  Cand cand[8] = {};
  int ncand = 0;
  int chosen = -1;  // index into cand once decided
  int calls = 0;
};
constexpr int kTuneWarm = 2, kTuneReps = 4;

// After the first enqueue an early return must not leave kernels writing into the caller's (mapped) buffers or
// into the staging buffers the next call reuses: every error path drains the three streams first.
static int host_fail(PhcHostStep* c, int rc, cudaError_t e);
#define HOST_CUDA(ctx, call)                                        \
  do {                                                              \
    cudaError_t e__ = (call);                                       \
    if (e__ != cudaSuccess) return host_fail((ctx), PHC_ERR_CUDA, e__); \
  } while (0)

static int host_fail(PhcHostStep* c, int rc, cudaError_t e) {
  if (c) {
    if (e != cudaSuccess) c->last_cuda = e;
    for (auto& s : c->streams)
      if (s) (void)cudaStreamSynchronize(s);
    (void)cudaGetLastError();
  }
  return rc;
}

extern "C" {

void phc_host_step_destroy(PhcHostStep* c) {
  if (!c) return;
  cudaFree(c->d_state);
  cudaFree(c->d_obs);
  cudaFree(c->d_clock);
  cudaFree(c->d_out);
  cudaFree(c->term_dist);
  if (c->h_clock) cudaFreeHost(c->h_clock);
  if (c->h_out) cudaFreeHost(c->h_out);
  for (auto& s : c->streams)
    if (s) cudaStreamDestroy(s);
  delete c;
}

int phc_host_step_create(const PhcLib* lib, int64_t max_envs, int32_t time_steps, int32_t num_chunks,
                         const float* termination_distances_host, uint32_t reset_body_mask, int32_t use_mean,
                         int32_t enable_early_termination, float dt, const PhcRewardSpec* rwd, PhcHostStep** out) {
  if (!lib || !termination_distances_host || !rwd || !out) return PHC_ERR_NULL;
  if (max_envs <= 0 || time_steps < 1 || time_steps > PHC_MAX_TIME_STEPS || num_chunks < 0 || num_chunks > 1024)
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
  if (const char* m = getenv("PHC_HOST_PATH")) c->mode = !strcmp(m, "direct") ? 1 : !strcmp(m, "staged") ? 2 : 0;
  if (getenv("PHC_HOST_STAGED")) c->mode = 2;
  const size_t n = (size_t)max_envs;
  // every chunk's packed region is padded so each sub-array starts 16-B aligned
  c->pack_capacity = n * 32 + (size_t)(num_chunks > 8 ? num_chunks : 8) * 6 * 64 + 4096;
  {  // candidates: direct with 2 / 3 / 4 / 6 equal chunks, staged with 3 / 4 doubling chunks
    static const int direct_chunks[] = {3, 4, 2, 6}, staged_chunks[] = {3, 4};
    auto add = [&](int path, int chunks) {
      if (c->mode && c->mode != path) return;
      if (num_chunks > 0 && chunks != num_chunks) return;
      c->cand[c->ncand++] = {path, chunks, 0.0, 0};
    };
    if (num_chunks > 0) {
      add(1, num_chunks);
      add(2, num_chunks);
    } else {
      for (int ch : direct_chunks) add(1, ch);
      for (int ch : staged_chunks) add(2, ch);
    }
    if (c->ncand == 1) c->chosen = 0;
  }
  cudaError_t e = cudaSuccess;
  auto dmalloc = [&](void** p, size_t bytes) {
    if (e == cudaSuccess) e = cudaMalloc(p, bytes);
  };
  dmalloc((void**)&c->d_state, n * kStateFloats * sizeof(float));
  dmalloc((void**)&c->d_obs, n * c->obs_dim * sizeof(float));
  dmalloc((void**)&c->d_clock, c->pack_capacity);
  dmalloc((void**)&c->d_out, c->pack_capacity);
  dmalloc((void**)&c->term_dist, PHC_NUM_BODIES * sizeof(float));
  if (e == cudaSuccess) e = cudaMallocHost((void**)&c->h_clock, c->pack_capacity);
  if (e == cudaSuccess) e = cudaMallocHost((void**)&c->h_out, c->pack_capacity);
  if (e == cudaSuccess)
    e = cudaMemcpy(c->term_dist, termination_distances_host, PHC_NUM_BODIES * sizeof(float), cudaMemcpyHostToDevice);
  for (int i = 0; i < kStreams && e == cudaSuccess; ++i)
    e = cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    phc_host_step_destroy(c);
    return e == cudaErrorMemoryAllocation ? PHC_ERR_ALLOC : PHC_ERR_CUDA;
  }
  *out = c;
  return PHC_OK;
}

int64_t phc_host_step_h2d_bytes(const PhcHostStep* c, int64_t n) {
  (void)c;
  return n * (int64_t)(kStateFloats * sizeof(float) + kClockBytes);
}
int64_t phc_host_step_d2h_bytes(const PhcHostStep* c, int64_t n) {
  return n * (int64_t)(c->obs_dim * sizeof(float) + kOutBytes);
}

int phc_host_step(PhcHostStep* c, const PhcHostStepArgs* a, int64_t n) {
  if (!c || !a) return PHC_ERR_NULL;
  if (n == 0) return PHC_OK;
  if (n < 0 || n > c->max_envs) return PHC_ERR_SHAPE;
  if (!a->state || !a->progress_buf || !a->motion_start_times || !a->motion_start_times_offset ||
      !a->global_offset || !a->sampled_motion_ids || !a->obs_buf || !a->rew_buf || !a->reward_raw || !a->reset_buf ||
      !a->terminate_buf)
    return PHC_ERR_NULL;
  // Chunk sizes double (n/2^(C-1), n/2^(C-1), n/2^(C-2), .., n/2): the device->host direction
  // carries 3x the bytes and is the bottleneck, so the first chunk is small to start it early
  // while the later, larger chunks keep the per-transfer overhead low.  Boundaries fall on
  // multiples of 8 envs so every chunk's obs rows start 16-B aligned.
  int64_t bounds[1026];
  int nchunks = 0;
  auto make_bounds = [&](const int C) {
    nchunks = 0;
    int64_t lo = 0;
    bounds[0] = 0;
    for (int i = 0; i < C && lo < n; ++i) {
      const int shift = (i == 0) ? C - 1 : C - i;
      int64_t sz = shift < 62 ? (n >> shift) : 0;
      sz = (sz + 7) / 8 * 8;
      if (sz < 8) sz = 8;
      int64_t hi = (i == C - 1) ? n : (lo + sz < n ? lo + sz : n);
      bounds[++nchunks] = hi;
      lo = hi;
    }
    bounds[nchunks] = n;
  };
  // sub-array offsets inside a chunk's packed regions (each 16-B aligned)
  struct Layout {
    size_t ids, goff, start, soff, prog, clock_bytes;
    size_t rew, raw, oprog, reset, term, out_bytes;
  };
  auto layout = [](int64_t m) {
    Layout L;
    size_t o = 0;
    L.ids = o, o = align_up(o + (size_t)m * 8, 16);
    L.goff = o, o = align_up(o + (size_t)m * 12, 16);
    L.start = o, o = align_up(o + (size_t)m * 4, 16);
    L.soff = o, o = align_up(o + (size_t)m * 4, 16);
    L.prog = o, o = align_up(o + (size_t)m * 2, 16);
    L.clock_bytes = o;
    o = 0;
    L.rew = o, o = align_up(o + (size_t)m * 4, 16);
    L.raw = o, o = align_up(o + (size_t)m * 16, 16);
    L.oprog = o, o = align_up(o + (size_t)m * 2, 16);
    L.reset = o, o = align_up(o + (size_t)m, 16);
    L.term = o, o = align_up(o + (size_t)m, 16);
    L.out_bytes = o;
    return L;
  };

  {  // ---- can the direct path be used: every buffer is pinned / managed host memory mapped at its own address ----------------------
    const void* ptrs[11] = {a->state, a->progress_buf, a->motion_start_times, a->motion_start_times_offset,
                            a->global_offset, a->sampled_motion_ids, a->obs_buf, a->rew_buf, a->reward_raw,
                            a->reset_buf, a->terminate_buf};
    if (c->seen_direct < 0 || memcmp(ptrs, c->seen, sizeof(ptrs)) != 0) {
      int direct = 1;
      for (int i = 0; i < 11 && direct; ++i) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, ptrs[i]) != cudaSuccess) {
          (void)cudaGetLastError();
          direct = 0;
        } else if (at.type == cudaMemoryTypeDevice) {
          return PHC_ERR_UNSUPPORTED;  // this entry point takes HOST buffers (the clock arrays are memcpy'd by the CPU)
        } else if ((at.type != cudaMemoryTypeHost && at.type != cudaMemoryTypeManaged) || at.devicePointer != ptrs[i]) {
          direct = 0;  // pageable, or mapped at a different device address
        }
      }
      memcpy(c->seen, ptrs, sizeof(ptrs));
      c->seen_direct = direct;
    }
  }
  auto direct_path = [&](const int chunks) -> int {
      // Hybrid: everything INBOUND rides the host->device copy engine in a few chunks — the sim
      // rows directly from the caller's buffer, the five small clock arrays packed into one
      // transfer (SM-issued reads of system memory are 32-B PCIe round trips: 6 per block, ~70 us
      // per 4096-env step when the kernel read the clock in place) — while each chunk's kernel
      // WRITES obs rows, rewards and flags straight into the mapped host buffers.  The copy
      // engine pulling chunk c+1 and the kernel pushing chunk c's rows use opposite directions of
      // the link.  The advanced progress is written to the caller's buffer by the kernel as well.
      // equal chunks here: measured, a small first chunk does not help this path (the step is
      // bound by the kernels' posted writes, ~300 us per 4096 envs, plus the first copy)
      const int C = chunks < 1 ? 1 : chunks;
      int64_t per = (n + C - 1) / C;
      per = (per + 7) / 8 * 8;
      size_t coff = 0;
      int ci = 0;
      static const bool trace_on = getenv("PHC_HOST_TRACE") != nullptr;
      static int trace_calls = 0;
      const bool trace = trace_on && ++trace_calls == 20;
      cudaEvent_t ev0 = nullptr, evh[64], evk[64];
      double host_us[64];
      const auto th0 = std::chrono::steady_clock::now();
      if (trace) {
        cudaEventCreate(&ev0);
        cudaEventRecord(ev0, c->streams[0]);
      }
      for (int64_t lo = 0; lo < n; lo += per, ++ci) {
        const int64_t m = (n - lo) < per ? (n - lo) : per;
        const Layout L = layout(m);
        if (coff + L.clock_bytes > c->pack_capacity) return host_fail(c, PHC_ERR_SHAPE, cudaSuccess);
        cudaStream_t s = c->streams[ci % kStreams];
        float* st = c->d_state + lo * kStateFloats;
        unsigned char* hc = c->h_clock + coff;
        unsigned char* dc = c->d_clock + coff;
        memcpy(hc + L.ids, a->sampled_motion_ids + lo, (size_t)m * 8);
        memcpy(hc + L.goff, a->global_offset + lo * 3, (size_t)m * 12);
        memcpy(hc + L.start, a->motion_start_times + lo, (size_t)m * 4);
        memcpy(hc + L.soff, a->motion_start_times_offset + lo, (size_t)m * 4);
        memcpy(hc + L.prog, a->progress_buf + lo, (size_t)m * 2);
        HOST_CUDA(c, cudaMemcpyAsync(dc, hc, L.clock_bytes, cudaMemcpyHostToDevice, s));
        HOST_CUDA(c, cudaMemcpyAsync(st, a->state + lo * kStateFloats, (size_t)m * kStateFloats * sizeof(float),
                                     cudaMemcpyHostToDevice, s));
        PhcStepArgs k{};
        if (trace && ci < 64) {
          cudaEventCreate(&evh[ci]);
          cudaEventRecord(evh[ci], s);
        }
        k.body.pos = PhcView{st, kStateFloats, 13};
        k.body.rot = PhcView{st + 3, kStateFloats, 13};
        k.body.vel = PhcView{st + 7, kStateFloats, 13};
        k.body.ang_vel = PhcView{st + 10, kStateFloats, 13};
        k.body.num_bodies = PHC_NUM_BODIES;
        k.progress_buf = (int16_t*)(dc + L.prog);
        k.motion_start_times = (const float*)(dc + L.start);
        k.motion_start_times_offset = (const float*)(dc + L.soff);
        k.global_offset = (const float*)(dc + L.goff);
        k.sampled_motion_ids = (const int64_t*)(dc + L.ids);
        k.termination_distances = c->term_dist;
        k.reset_body_mask = c->reset_mask;
        k.use_mean = c->use_mean;
        k.enable_early_termination = c->early;
        k.advance_progress = 1;
        k.time_steps = c->T;
        k.dt = c->dt;
        k.rwd = c->rwd;
        k.obs_buf = a->obs_buf + lo * c->obs_dim;
        k.obs_stride = c->obs_dim;
        k.rew_buf = a->rew_buf + lo;
        k.reward_raw = a->reward_raw + lo * 4;
        k.reward_raw_stride = 4;
        k.reset_buf = a->reset_buf + lo;
        k.terminate_buf = a->terminate_buf + lo;
        k.flags = PHC_STEP_MAPPED_HOST_IO;
        k.obs_moments = nullptr;
        // the advanced progress goes straight into the caller's buffer too (no device->host copy at the tail)
        int rc = step_fused_mirrored(c->lib, &k, m, s, a->progress_buf + lo);
        if (rc) return host_fail(c, rc, cudaSuccess);
        if (trace && ci < 64) {
          cudaEventCreate(&evk[ci]);
          cudaEventRecord(evk[ci], s);
          host_us[ci] = std::chrono::duration<double>(std::chrono::steady_clock::now() - th0).count() * 1e6;
        }
        coff += L.clock_bytes;
      }
      for (int i = 0; i < kStreams; ++i) HOST_CUDA(c, cudaStreamSynchronize(c->streams[i]));
      if (trace) {
        const double tot = std::chrono::duration<double>(std::chrono::steady_clock::now() - th0).count() * 1e6;
        fprintf(stderr, "[phc_host trace] %d chunks of %lld envs, host total %.1f us\n", ci, (long long)per, tot);
        for (int i = 0; i < ci && i < 64; ++i) {
          float a_ms = 0, b_ms = 0;
          cudaEventElapsedTime(&a_ms, ev0, evh[i]);
          cudaEventElapsedTime(&b_ms, ev0, evk[i]);
          fprintf(stderr, "  chunk %d: submitted at %.1f us (host), H2D done %.1f us, kernel done %.1f us\n", i, host_us[i],
                  a_ms * 1e3, b_ms * 1e3);
        }
      }
      return PHC_OK;
  };
  auto staged_path = [&](const int chunks) -> int {
  const cudaMemcpyKind H2D = cudaMemcpyHostToDevice, D2H = cudaMemcpyDeviceToHost;
  make_bounds(chunks < 1 ? 1 : chunks);

  size_t coff = 0, ooff = 0;
  static const bool strace_on = getenv("PHC_HOST_TRACE") != nullptr;
  static int strace_calls = 0;
  const bool strace = strace_on && ++strace_calls == 20;
  cudaEvent_t sev0 = nullptr, sevh[16], sevk[16], sevd[16];
  if (strace) {
    cudaEventCreate(&sev0);
    cudaEventRecord(sev0, c->streams[0]);
  }
  for (int ci = 0; ci < nchunks; ++ci) {
    const int64_t lo = bounds[ci], m = bounds[ci + 1] - bounds[ci];
    if (m <= 0) continue;
    const Layout L = layout(m);
    if (coff + L.clock_bytes > c->pack_capacity || ooff + L.out_bytes > c->pack_capacity)
      return host_fail(c, PHC_ERR_SHAPE, cudaSuccess);
    cudaStream_t s = c->streams[ci % kStreams];
    // pack the chunk's clock on the host (tens of KB), then ONE transfer
    unsigned char* hc = c->h_clock + coff;
    memcpy(hc + L.ids, a->sampled_motion_ids + lo, (size_t)m * 8);
    memcpy(hc + L.goff, a->global_offset + lo * 3, (size_t)m * 12);
    memcpy(hc + L.start, a->motion_start_times + lo, (size_t)m * 4);
    memcpy(hc + L.soff, a->motion_start_times_offset + lo, (size_t)m * 4);
    memcpy(hc + L.prog, a->progress_buf + lo, (size_t)m * 2);
    unsigned char* dc = c->d_clock + coff;
    unsigned char* dout = c->d_out + ooff;
    HOST_CUDA(c, cudaMemcpyAsync(c->d_state + lo * kStateFloats, a->state + lo * kStateFloats,
                                 (size_t)m * kStateFloats * sizeof(float), H2D, s));
    HOST_CUDA(c, cudaMemcpyAsync(dc, hc, L.clock_bytes, H2D, s));

    PhcStepArgs k{};
    float* st = c->d_state + lo * kStateFloats;
    k.body.pos = PhcView{st, kStateFloats, 13};
    k.body.rot = PhcView{st + 3, kStateFloats, 13};
    k.body.vel = PhcView{st + 7, kStateFloats, 13};
    k.body.ang_vel = PhcView{st + 10, kStateFloats, 13};
    k.body.num_bodies = PHC_NUM_BODIES;
    k.progress_buf = (int16_t*)(dc + L.prog);
    k.motion_start_times = (const float*)(dc + L.start);
    k.motion_start_times_offset = (const float*)(dc + L.soff);
    k.global_offset = (const float*)(dc + L.goff);
    k.sampled_motion_ids = (const int64_t*)(dc + L.ids);
    k.termination_distances = c->term_dist;
    k.reset_body_mask = c->reset_mask;
    k.use_mean = c->use_mean;
    k.enable_early_termination = c->early;
    k.advance_progress = 1;
    k.time_steps = c->T;
    k.dt = c->dt;
    k.rwd = c->rwd;
    k.obs_buf = c->d_obs + lo * c->obs_dim;
    k.obs_stride = c->obs_dim;
    k.rew_buf = (float*)(dout + L.rew);
    k.reward_raw = (float*)(dout + L.raw);
    k.reward_raw_stride = 4;
    k.reset_buf = dout + L.reset;
    k.terminate_buf = dout + L.term;
    k.obs_moments = nullptr;
    if (strace && ci < 16) {
      cudaEventCreate(&sevh[ci]);
      cudaEventRecord(sevh[ci], s);
    }
    int rc = phc_step_fused(c->lib, &k, m, s);
    if (rc) return host_fail(c, rc, cudaSuccess);
    if (strace && ci < 16) {
      cudaEventCreate(&sevk[ci]);
      cudaEventRecord(sevk[ci], s);
    }
    // the advanced progress rides back with the scalar outputs
    HOST_CUDA(c, cudaMemcpyAsync(dout + L.oprog, dc + L.prog, (size_t)m * 2, cudaMemcpyDeviceToDevice, s));
    HOST_CUDA(c, cudaMemcpyAsync(a->obs_buf + lo * c->obs_dim, c->d_obs + lo * c->obs_dim,
                                 (size_t)m * c->obs_dim * sizeof(float), D2H, s));
    HOST_CUDA(c, cudaMemcpyAsync(c->h_out + ooff, dout, L.out_bytes, D2H, s));
    if (strace && ci < 16) {
      cudaEventCreate(&sevd[ci]);
      cudaEventRecord(sevd[ci], s);
    }
    coff += L.clock_bytes;
    ooff += L.out_bytes;
  }
  for (int i = 0; i < kStreams; ++i) HOST_CUDA(c, cudaStreamSynchronize(c->streams[i]));
  if (strace) {
    fprintf(stderr, "[phc_host trace] staged, %d chunks\n", nchunks);
    for (int i = 0; i < nchunks && i < 16; ++i) {
      float a_ms = 0, b_ms = 0, d_ms = 0;
      cudaEventElapsedTime(&a_ms, sev0, sevh[i]);
      cudaEventElapsedTime(&b_ms, sev0, sevk[i]);
      cudaEventElapsedTime(&d_ms, sev0, sevd[i]);
      fprintf(stderr, "  chunk %d (%lld envs): H2D done %.1f us, kernel done %.1f us, D2H done %.1f us\n", i,
              (long long)(bounds[i + 1] - bounds[i]), a_ms * 1e3, b_ms * 1e3, d_ms * 1e3);
    }
  }
  // unpack the scalar outputs
  ooff = 0;
  for (int ci = 0; ci < nchunks; ++ci) {
    const int64_t lo = bounds[ci], m = bounds[ci + 1] - bounds[ci];
    if (m <= 0) continue;
    const Layout L = layout(m);
    const unsigned char* ho = c->h_out + ooff;
    memcpy(a->rew_buf + lo, ho + L.rew, (size_t)m * 4);
    memcpy(a->reward_raw + lo * 4, ho + L.raw, (size_t)m * 16);
    memcpy(a->progress_buf + lo, ho + L.oprog, (size_t)m * 2);
    memcpy(a->reset_buf + lo, ho + L.reset, (size_t)m);
    memcpy(a->terminate_buf + lo, ho + L.term, (size_t)m);
    ooff += L.out_bytes;
  }
  return PHC_OK;
  };

  // ---- which schedule ------------------------------------------------------------------------
  const int fallback_chunks = c->chunks > 0 ? c->chunks : 3;
  if (c->seen_direct != 1) return staged_path(fallback_chunks);  // pageable caller: copy-engine path only
  auto run = [&](const PhcHostStep::Cand& k) { return k.path == 1 ? direct_path(k.chunks) : staged_path(k.chunks); };
  if (c->chosen >= 0) return run(c->cand[c->chosen]);
  // Tuning: calls 0-1 warm up, then the candidates take turns (kTuneReps timed calls each) and the fastest mean
  // stays.  What decides is the machine, under whatever load the other GPUs of the box put on the host at that moment:
  // one GPU posts its obs rows into host memory faster than a copy engine drains a staging buffer, eight GPUs posting
  // into one root complex may not (profiles/r2_e2e_floor.md); and on some hosts the inbound copy of the next chunk
  // crawls while a kernel is posting (its read requests queue behind the posted writes), which moves the best chunk
  // count (profiles/r2_host_timeline.md).
  const int k = c->calls++;
  const int which = k < kTuneWarm ? k % c->ncand : (k - kTuneWarm) % c->ncand;
  const auto t0 = std::chrono::steady_clock::now();
  const int rc = run(c->cand[which]);
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (rc == PHC_OK && k >= kTuneWarm) {
    // the first timed call of a candidate warms its own launch configuration: not counted.  A candidate's figure is
    // its FASTEST call (what the schedule can do; a mean over three calls carries the host's jitter), and the first
    // candidate — three chunks, direct: the best schedule on most hosts — is only given up for one that is 2 % faster.
    if (k - kTuneWarm >= c->ncand) {
      c->cand[which].t = c->cand[which].n == 0 || dt < c->cand[which].t ? dt : c->cand[which].t;
      c->cand[which].n += 1;
    }
    bool done = true;
    for (int i = 0; i < c->ncand; ++i) done = done && c->cand[i].n >= kTuneReps;
    if (done) {
      int best = 0;
      double best_t = c->cand[0].t * 0.98;
      for (int i = 1; i < c->ncand; ++i)
        if (c->cand[i].t < best_t) best = i, best_t = c->cand[i].t;
      c->chosen = best;
      if (getenv("PHC_HOST_TRACE"))
        for (int i = 0; i < c->ncand; ++i)
          fprintf(stderr, "[phc_host tune] %s, %d chunks: %.1f us%s\n", c->cand[i].path == 1 ? "direct" : "staged",
                  c->cand[i].chunks, c->cand[i].t * 1e6, i == best ? "  <- kept" : "");
    }
  }
  return rc;
}

int phc_host_step_path(const PhcHostStep* c) {
  if (!c) return 0;
  if (c->seen_direct == 0) return 2;
  return c->chosen >= 0 ? c->cand[c->chosen].path : 0;
}

int phc_host_step_chunks(const PhcHostStep* c) {
  if (!c) return 0;
  if (c->seen_direct == 0) return c->chunks > 0 ? c->chunks : 3;
  return c->chosen >= 0 ? c->cand[c->chosen].chunks : 0;
}

int phc_host_step_tuning_calls(const PhcHostStep* c) {
  if (!c || c->ncand <= 1) return 0;
  return kTuneWarm + c->ncand * (kTuneReps + 1);
}

}  // extern "C"
