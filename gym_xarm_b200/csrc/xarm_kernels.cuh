// xarm_kernels.cuh - per-env bodies of the kernels (init / step / reset / obs).  The __global__ wrappers live in
// xarm_lib.cu; tests/hostsim compiles these same bodies for the host to check the kernel logic without a GPU.
#pragma once
#include "xarm_env.cuh"

struct KArgs {
  float* state;          // [S][N] SoA slab
  float* ep_return;      // [N] running undiscounted return of the current episode
  uint8_t* need_reset;   // [N] set by the step kernel (auto-reset) / consumed by the reset kernel
  double* stats;         // [5] episodes, sum return, sum length, sum success, diverged
  XarmBuffers b;
  ResetCfg rc;
  int64_t n;
  int auto_reset;
  // load balancing (DESIGN.md "warp homogeneity"): thread t of the step kernel simulates env perm[t]
  int* perm;             // [N] envs ordered by predicted contact load (nearest gripper-object distance first)
  int* bin_slot;         // [N] (bin << 24 | slot within bin) scratch of the classify pass
  int* bin_counts;       // [XARM_LOAD_BINS]
  int* reset_list;       // [N] compacted list of envs the step kernel finished (auto-reset)
  int* reset_count;      // [1]
};
#define XARM_LOAD_BINS 8

// predicted contact load of env i for the coming step: bin 0 = gripper closest to an object (contacts likely)
template <class T>
XD int body_load_bin(const KArgs& a, int64_t i) {
  using MD = typename T::MD;
  Env<T> e;
  env_load<T>(e, a.state, a.n, i);
  const int limit = a.rc.max_episode_steps > 0 ? a.rc.max_episode_steps : T::MAX_STEPS;
  if (a.auto_reset && e.step_count + 1 >= limit) return 0;  // will hit the time limit: step + reset in this launch
  if (T::NOBJ == 0 || !MD::HAS_BOXES) return 1;
  float dmin = 1e30f;
#pragma unroll
  for (int arm = 0; arm < T::NARM; arm++) {
    M3 Re; V3 pe, org[7], axs[7];
    arm_fk7<T>(arm, e.arm[arm].q, Re, pe, org, axs);
    V3 gc = pe + Re * v3(0.f, 0.f, 0.043f);  // middle of the hand + finger hulls along the hand axis
    for (int o = 0; o < T::NOBJ; o++) dmin = fminf(dmin, norm(gc - e.obj[o].pos));
  }
  int bin = 1 + (int)((dmin - 0.10f) * 20.f);
  return bin < 1 ? 1 : (bin >= XARM_LOAD_BINS ? XARM_LOAD_BINS - 1 : bin);
}
struct StepStats {
  float eps, ret, len, suc, div;
};

template <class T>
XD void write_obs(const KArgs& a, int64_t i, const Obs<T>& o) {
  if (a.b.observation)
#pragma unroll
    for (int k = 0; k < T::O; k++) a.b.observation[i * T::O + k] = o.obs[k];
  if (a.b.achieved_goal)
#pragma unroll
    for (int k = 0; k < T::G; k++) a.b.achieved_goal[i * T::G + k] = o.ag[k];
  if (a.b.desired_goal)
#pragma unroll
    for (int k = 0; k < T::G; k++) a.b.desired_goal[i * T::G + k] = o.dg[k];
}

template <class T>
XD bool env_finite(const Env<T>& e) {
  float s = 0.f;
#pragma unroll
  for (int a = 0; a < T::NARM; a++)
#pragma unroll
    for (int k = 0; k < T::MD::N; k++) s += e.arm[a].q[k] + e.arm[a].qd[k];
#pragma unroll
  for (int o = 0; o < T::NOBJ; o++) s += e.obj[o].pos.x + e.obj[o].pos.y + e.obj[o].pos.z + e.obj[o].v.x + e.obj[o].v.y + e.obj[o].v.z + e.obj[o].quat.w;
  return isfinite(s);
}

template <class T>
XD void body_init(const KArgs& a, int64_t i) {
  Env<T> e;
  env_construct<T>(e, a.rc, a.rc.env_index_base + i);
  env_store<T>(e, a.state, a.n, i);
  a.ep_return[i] = 0.f;
  a.need_reset[i] = 0;
}

// Env.step for env i
template <class T>
XD void body_step(const KArgs& a, int64_t i, StepStats& st, bool block_sync = false, bool valid = true) {
  Env<T> e;
  env_load<T>(e, a.state, a.n, i);
  float act[T::A];
#pragma unroll
  for (int k = 0; k < T::A; k++) act[k] = a.b.actions[i * T::A + k];
  Obs<T> o;
  StepOut so;
  env_step<T>(e, act, a.rc, o, so);
  if (!valid) return;  // padding lane of a phase-synchronised block: simulated a copy, stores nothing
  if (!env_finite<T>(e)) {  // NaN guard: rebuild the env, end the episode
    uint32_t ep = e.episode;
    env_construct<T>(e, a.rc, a.rc.env_index_base + i);
    e.episode = ep;
    get_obs<T>(e, o);
    so.reward = 0.f; so.success = 0.f; so.done = true; so.truncated = true;
    st.div = 1.f;
  }
  write_obs<T>(a, i, o);
  a.b.reward[i] = so.reward;
  a.b.done[i] = so.done;
  a.b.success[i] = so.success;
  if (a.b.truncated) a.b.truncated[i] = so.truncated;
  float ret = a.ep_return[i] + so.reward;
  if (so.done) {
    if (a.b.terminal_observation) {
      float* t = a.b.terminal_observation + i * (T::O + 2 * T::G);
#pragma unroll
      for (int k = 0; k < T::O; k++) t[k] = o.obs[k];
#pragma unroll
      for (int k = 0; k < T::G; k++) { t[T::O + k] = o.ag[k]; t[T::O + T::G + k] = o.dg[k]; }
    }
    st.eps = 1.f; st.ret = ret; st.len = (float)e.step_count; st.suc = so.success;
    ret = 0.f;
  }
  a.ep_return[i] = ret;
  a.need_reset[i] = 0;
  if (so.done && a.auto_reset) {
    // VecEnv auto-reset, fused: the finishing lane runs Env.reset() right away (its warp finishes later, the other
    // warps of the grid keep the SMs busy meanwhile - a separate reset kernel over the few finished envs would add its
    // whole 6-sim-step latency to every step).  Envs that reach their time limit are grouped into the same warps by
    // the load-balancing permutation.
    env_reset<T>(e, a.rc, a.rc.env_index_base + i);
    get_obs<T>(e, o);
    e.d_old = np_dist(o.ag, o.dg, T::G);
    write_obs<T>(a, i, o);
  }
  env_store<T>(e, a.state, a.n, i);
}

// Env.reset for env i
template <class T>
XD void body_reset(const KArgs& a, int64_t i, bool clear_return) {
  Env<T> e;
  env_load<T>(e, a.state, a.n, i);
  env_reset<T>(e, a.rc, a.rc.env_index_base + i);
  Obs<T> o;
  get_obs<T>(e, o);
  e.d_old = np_dist(o.ag, o.dg, T::G);  // [REF xarm_reach.py:100]
  write_obs<T>(a, i, o);
  env_store<T>(e, a.state, a.n, i);
  a.need_reset[i] = 0;
  if (clear_return) a.ep_return[i] = 0.f;
}

template <class T>
XD void body_obs(const KArgs& a, int64_t i) {
  Env<T> e;
  env_load<T>(e, a.state, a.n, i);
  Obs<T> o;
  get_obs<T>(e, o);
  write_obs<T>(a, i, o);
}
