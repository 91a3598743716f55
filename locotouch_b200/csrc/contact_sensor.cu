// K18: ContactSensor bookkeeping -- net-force history ring and the air / contact-time state machine of every sensor body.
// Replaces the per-step tensor code of IsaacLab's ContactSensor._update_buffers_impl / reset ([IL], IsaacLab 2.2.1, not under the
// reference tree; restated from SURVEY.md App. B and used by the reference through `locomotion_base_env_cfg.py:35-39,358-359`:
// history_length 3, track_air_time on, update every step) whose outputs feed a2 / a4 / a7 / a8 / a9 (`current_air_time`,
// `current_contact_time`, `last_air_time`, `last_contact_time`, `net_forces_w_history`; reference locotouch/mdp/rewards.py:116-156,
// 596-604, observations.py:60-66).  ~10 boolean-index / where / clone launches per sensor and step become one pass:
//   history[:, 1:] = history[:, :-1]; history[:, 0] = F                              (ring shift + insert)
//   is_contact       = |F| > force_threshold
//   first_contact    = current_air_time > 0  and is_contact;   first_detached = current_contact_time > 0 and not is_contact
//   last_air_time     = first_contact  ? current_air_time + dt     : last_air_time
//   current_air_time  = is_contact     ? 0                         : current_air_time + dt
//   last_contact_time = first_detached ? current_contact_time + dt : last_contact_time
//   current_contact_time = is_contact  ? current_contact_time + dt : 0
// and, for envs flagged in `reset_mask`, ContactSensor.reset(env_ids): forces, history and the four timers cleared instead.
// One thread per (env, body); 12 B x (H + 1) + 16 B read, 12 B x H + 16 B written per unit.
#include "lt_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
contact_sensor_kernel(const float* __restrict__ forces, float* __restrict__ net_forces_w, float* __restrict__ history, int H, int N, int Bd,
                      float* __restrict__ cur_air, float* __restrict__ last_air, float* __restrict__ cur_contact, float* __restrict__ last_contact,
                      const float* __restrict__ dt_per_env, float dt_scalar, float threshold, const uint8_t* __restrict__ reset_mask) {
  const int64_t total = (int64_t)N * Bd;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / Bd), b = (int)(i - (int64_t)n * Bd);
    const bool reset = reset_mask != nullptr && reset_mask[n] != 0;
    float fx = 0.f, fy = 0.f, fz = 0.f;
    if (!reset) {
      fx = forces[3 * i];
      fy = forces[3 * i + 1];
      fz = forces[3 * i + 2];
    }
    if (net_forces_w != nullptr && (net_forces_w != forces || reset)) {
      net_forces_w[3 * i] = fx;
      net_forces_w[3 * i + 1] = fy;
      net_forces_w[3 * i + 2] = fz;
    }
    if (history != nullptr) {
      float* h = history + ((size_t)n * H * Bd + b) * 3;      // [N, H, Bd, 3]
      const size_t stride = (size_t)Bd * 3;
      for (int k = H - 1; k > 0; --k) {
        float* dst = h + k * stride;
        const float* src = h + (k - 1) * stride;
        dst[0] = reset ? 0.f : src[0];
        dst[1] = reset ? 0.f : src[1];
        dst[2] = reset ? 0.f : src[2];
      }
      h[0] = fx;
      h[1] = fy;
      h[2] = fz;
    }
    if (cur_air == nullptr) continue;
    if (reset) {
      cur_air[i] = last_air[i] = cur_contact[i] = last_contact[i] = 0.f;
      continue;
    }
    const float dt = dt_per_env != nullptr ? dt_per_env[n] : dt_scalar;
    // torch.norm(F, dim=-1): sqrt of the sum of squares, accumulated x, y, z
    const float norm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)), __fmul_rn(fz, fz)));
    const bool is_contact = norm > threshold;
    const float ca = cur_air[i], cc = cur_contact[i];
    const bool first_contact = ca > 0.f && is_contact;
    const bool first_detached = cc > 0.f && !is_contact;
    if (first_contact) last_air[i] = __fadd_rn(ca, dt);
    cur_air[i] = is_contact ? 0.f : __fadd_rn(ca, dt);
    if (first_detached) last_contact[i] = __fadd_rn(cc, dt);
    cur_contact[i] = is_contact ? __fadd_rn(cc, dt) : 0.f;
  }
}

}  // namespace

extern "C" int lt_contact_sensor_update(const float* forces, float* net_forces_w, float* net_forces_w_history, int history_length, int N, int num_bodies,
                                        float* current_air_time, float* last_air_time, float* current_contact_time, float* last_contact_time,
                                        const float* dt_per_env, float dt, float force_threshold, const uint8_t* reset_mask, void* stream) {
  if (!forces || N <= 0 || num_bodies <= 0 || history_length < 0) return LT_ERR_INVALID_ARG;
  if (net_forces_w_history && history_length == 0) return LT_ERR_INVALID_ARG;
  const bool any_timer = current_air_time || last_air_time || current_contact_time || last_contact_time;
  if (any_timer && !(current_air_time && last_air_time && current_contact_time && last_contact_time)) return LT_ERR_INVALID_ARG;
  const int64_t total = (int64_t)N * num_bodies;
  int64_t blocks = lt::ceil_div(total, 256);
  const int64_t cap = 8 * (int64_t)lt::sm_count();
  if (blocks > cap) blocks = cap;
  contact_sensor_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(forces, net_forces_w, net_forces_w_history, history_length, N, num_bodies,
                                                                      current_air_time, last_air_time, current_contact_time, last_contact_time,
                                                                      dt_per_env, dt, force_threshold, reset_mask);
  return lt::check_launch();
}
