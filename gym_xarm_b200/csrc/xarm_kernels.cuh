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
  int* reset_list;       // [N] compacted list of envs that finished this step (auto-reset) / were selected by xarm_reset
  int* reset_count;      // [1]
  // step pipeline (xarm_pipeline.cuh)
  float* scratch;        // [pipe_scratch_words][N] rows of the current substep
  int* form;             // [N] light / heavy classification of the current substep (+ manifold point count)
  int* heavy_list;       // [N] envs the setup kernel classified as heavy: entry t at heavy_list[heavy_dir * t]
  int heavy_dir;         // +1: the list grows up from heavy_list; -1: down from it (the two concurrent branches of a step share one array)
  int* heavy_count;      // [XARM_PIPE_COUNTERS] one counter per simulate() pass and substep (zeroed by k_pipe_begin)
  int* rng_draw;         // [N] Philox draw counter carried from the placement stage of a reset to its goal stage
  const int* list;       // optional: thread t works on env list[t] (auto-reset tail); NULL = identity
  const int* list_count;
  // development timeline (XARM_TIMELINE=1): first block start / last block end of every pipeline launch, %globaltimer ns
  unsigned long long* tl;  // [5 * XARM_TL_SLOTS + 8] or NULL
  int tl_slot;
  int tl_branch;           // 0 main, 1 early, 2 late tail: env-substep counters at tl[5 * XARM_TL_SLOTS + {branch: heavy, 3 + branch: all}]
  // SM partition of a split step (xarm_lib.cu): launches of the main branch carry a work counter (blocks claim their
  // chunks dynamically) and leave at once when they land on an SM of sm_mask - those SMs belong to the early branch
  int spread4_max;                 // setup kernel: lists up to this size run 4 lanes per env (xarm_lib.cu: pipe_spread_setup)
  int light_dual;                  // list launches of the light kernel come in two register budgets (xarm_lib.cu)
  int heavy_dual;                  // heavy kernels come in two forms: 1 = this launch works only on lists <= XARM_FUSED_MAX (fused), 2 = only on longer ones
  int* work;                       // NULL: static block-stride loop, every SM
  unsigned long long sm_mask[4];   // bit smid set: reserved for the early branch
};
#define XARM_TL_SLOTS 4096
#define XARM_MAX_SUBSTEPS 32
#define XARM_PIPE_PASSES 16  /* simulate() passes of one xarm_step: per branch the step itself + up to 6 of the auto-reset tail */
#define XARM_PIPE_COUNTERS (XARM_PIPE_PASSES * XARM_MAX_SUBSTEPS)
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

// Env.step for env i in ONE piece (step + inline auto-reset).  The CUDA path runs the step as the kernel pipeline of
// xarm_pipeline.cuh; this fused form is what tests/hostsim checks the pipeline against (same results, bit for bit on
// the host) and the plain statement of the step's semantics.
template <class T>
XD void body_step(const KArgs& a, int64_t i, StepStats& st) {
  Env<T> e;
  env_load<T>(e, a.state, a.n, i);
  float act[T::A];
#pragma unroll
  for (int k = 0; k < T::A; k++) act[k] = a.b.actions[i * T::A + k];
  Obs<T> o;
  StepOut so;
  env_step<T>(e, act, a.rc, o, so);
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
    // VecEnv auto-reset: the finished env runs Env.reset() right away
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
  if (clear_return) e.step_count = initial_step_count<T>(a.rc, a.rc.env_index_base + i);
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
