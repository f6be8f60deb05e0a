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

/* compute_humanoid_im_reset, envs/common.py:325-364 (use_mean = False branch and the use_mean
 * branch's distances).  pos / ref are dense [n, R, 3]; pass_time, reset, terminated are bytes. */
void phc_oracle_im_reset(int64_t n, int32_t R, const float* pos, const float* ref, const int16_t* progress,
                         const uint8_t* pass_time, const float* term_dist, int32_t early, uint8_t* reset,
                         uint8_t* terminated, float* dist_out /* [n,R] or NULL */) {
  for (int64_t i = 0; i < n; ++i) {
    int fallen = 0;
    for (int32_t b = 0; b < R; ++b) {
      const float* p = pos + (i * R + b) * 3;
      const float* r = ref + (i * R + b) * 3;
      const float x = p[0] - r[0], y = p[1] - r[1], z = p[2] - r[2];
      /* torch.norm(dim=-1) over 3 on ATen CPU: sqrt(fma(z,z,fma(y,y,x*x))) */
      const float d = sqrtf(fmaf(z, z, fmaf(y, y, x * x)));
      if (dist_out) dist_out[i * R + b] = d;
      if (d > term_dist[b]) fallen = 1;
    }
    if (!early) fallen = 0;
    if (!(progress[i] > 1)) fallen = 0; /* has_fallen *= progress_buf > 1 (:353) */
    terminated[i] = (uint8_t)fallen;
    reset[i] = pass_time[i] ? 1 : (uint8_t)fallen; /* :362 */
  }
}
