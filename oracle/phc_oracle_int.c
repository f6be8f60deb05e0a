/*
 * phc_oracle_int.c — plain-C restatement of the INTEGER / FLAG outputs of the PHC step path.
 * TEST INFRASTRUCTURE ONLY (see oracle/phc_oracle.py for the full float oracle and the rules on
 * who may call into oracle/).  A second, independent pin for the outputs that must be bit-exact:
 * frame indices and blend (motion_lib.py:655-665), the reset start time (motion_lib.py:526-535),
 * and the reset / termination flags (envs/common.py:325-364).  Checked against the
 * reference-generated fixtures in tests/test_oracle_c.py.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: every + - * / is rounded separately, as in
 * ATen's one-kernel-per-op evaluation; the one place ATen's CPU kernel fuses — the 3-vector
 * norm — calls fmaf explicitly.)
 */
#include <math.h>
#include <stdint.h>

static float clip01(float x) { return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x); }

/* MotionLibBase._calc_frame_blend(time, len, num_frames, dt), motion_lib.py:655-665 */
void phc_oracle_frame_blend(int64_t n, const float* time, const float* len, const int64_t* num_frames,
                            const float* dt, int64_t* idx0, int64_t* idx1, float* blend) {
  for (int64_t i = 0; i < n; ++i) {
    float t = time[i];
    float phase = clip01(t / len[i]); /* from the un-clamped time (:656-657) */
    if (t < 0.0f) t = 0.0f;            /* time[time < 0] = 0 (:658)          */
    const int64_t nf = num_frames[i];
    const int64_t a = (int64_t)(phase * (float)(nf - 1)); /* .long() truncates (:660) */
    const int64_t b = a + 1 < nf - 1 ? a + 1 : nf - 1;
    idx0[i] = a;
    idx1[i] = b;
    blend[i] = clip01((t - (float)a * dt[i]) / dt[i]);
  }
}

/* MotionLibBase.sample_time_interval, motion_lib.py:526-535, with the uniform numbers passed in:
 * ((phase * motion_len) / curr_fps).long() * curr_fps, curr_fps = 1/30 */
void phc_oracle_sample_time(int64_t n, const float* phase, const float* motion_len, float* out) {
  const float curr_fps = (float)(1.0 / 30.0);
  for (int64_t i = 0; i < n; ++i) {
    const int64_t k = (int64_t)((phase[i] * motion_len[i]) / curr_fps);
    out[i] = (float)k * curr_fps;
  }
}

/* torch.sum(x, dim=-1) of a contiguous row of n floats on ATen CPU (aten/src/ATen/native/cpu/SumKernel.cpp, the
 * AVX2 build torch dispatches to on AVX2 and AVX-512 hosts):
 *   n >= 8: vectorized_inner_sum — the n / 8 full 8-lane vectors are added lane-wise (row_sum: four interleaved
 *           vector partials, combined 0+1+2+3), the scalar accumulator then takes the n % 8 tail elements in order
 *           and the 8 lanes after them;
 *   n <  8: scalar_inner_sum — row_sum with four scalar partials. */
float phc_oracle_row_sum(const float* v, int32_t n) {
  if (n < 8) {
    float p[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const int32_t si = n / 4;
    for (int32_t i = 0; i < si; ++i)
      for (int k = 0; k < 4; ++k) p[k] += v[i * 4 + k];
    for (int32_t i = si * 4; i < n; ++i) p[0] += v[i];
    for (int k = 1; k < 4; ++k) p[0] += p[k];
    return p[0];
  }
  const int32_t nv = n / 8, si = nv / 4;
  float part[4][8] = {{0}};
  for (int32_t i = 0; i < si; ++i)
    for (int k = 0; k < 4; ++k)
      for (int l = 0; l < 8; ++l) part[k][l] += v[(i * 4 + k) * 8 + l];
  for (int32_t i = si * 4; i < nv; ++i)
    for (int l = 0; l < 8; ++l) part[0][l] += v[i * 8 + l];
  for (int k = 1; k < 4; ++k)
    for (int l = 0; l < 8; ++l) part[0][l] += part[k][l];
  float acc = 0.0f;
  for (int32_t k = nv * 8; k < n; ++k) acc += v[k];
  for (int l = 0; l < 8; ++l) acc += part[0][l];
  return acc;
}

void phc_oracle_row_sums(int64_t rows, int32_t n, const float* x, float* out) {
  for (int64_t i = 0; i < rows; ++i) out[i] = phc_oracle_row_sum(x + i * n, n);
}

/* compute_humanoid_im_reset, envs/common.py:325-364.  pos / ref are dense [n, R, 3] (the caller's gather of the
 * reset bodies, humanoid_phc.py:1321-1322); pass_time, reset, terminated are bytes.
 *   use_mean == 0: any_b(|p_b - r_b| > term_dist[b])                       (:348-350)
 *   use_mean != 0: mean_b |p_b - r_b| > term_dist[0], mean = row sum / R   (:343-346) */
void phc_oracle_im_reset(int64_t n, int32_t R, const float* pos, const float* ref, const int16_t* progress,
                         const uint8_t* pass_time, const float* term_dist, int32_t early, int32_t use_mean,
                         uint8_t* reset, uint8_t* terminated, float* dist_out /* [n,R] or NULL */) {
  float d[64];
  for (int64_t i = 0; i < n; ++i) {
    int fallen = 0;
    for (int32_t b = 0; b < R; ++b) {
      const float* p = pos + (i * R + b) * 3;
      const float* r = ref + (i * R + b) * 3;
      const float x = p[0] - r[0], y = p[1] - r[1], z = p[2] - r[2];
      /* torch.norm(dim=-1) over 3 on ATen CPU: sqrt(fma(z,z,fma(y,y,x*x))) */
      d[b] = sqrtf(fmaf(z, z, fmaf(y, y, x * x)));
      if (dist_out) dist_out[i * R + b] = d[b];
      if (!use_mean && d[b] > term_dist[b]) fallen = 1;
    }
    if (use_mean) fallen = (phc_oracle_row_sum(d, R) / (float)R) > term_dist[0];
    if (!early) fallen = 0;
    if (!(progress[i] > 1)) fallen = 0; /* has_fallen *= progress_buf > 1 (:353) */
    terminated[i] = (uint8_t)fallen;
    reset[i] = pass_time[i] ? 1 : (uint8_t)fallen; /* :362 */
  }
}
