// hostsim.cpp - TEST HARNESS ONLY: compiles the per-env bodies of the CUDA kernels (gym_xarm_b200/csrc/*.cuh) for the
// host with XARM_HOST_SIM, so that the kernel logic can be checked against the oracle on a machine without a GPU
// (`pytest -m "not gpu"`).  It is not part of the product, is never imported by gym_xarm_b200 and is not a fallback:
// the product library has no CPU path.  float arithmetic, same code as the device path minus FMA contraction.
#define XARM_HOST_SIM 1
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <math.h>
#include <stdint.h>

#include "../../gym_xarm_b200/csrc/xarm_math.cuh"  // first: in the double-precision diagnostic build it redefines float
#include "../../gym_xarm_b200/csrc/xarm_pipeline.cuh"

struct Ops {
  void (*init)(const KArgs&);
  void (*step)(const KArgs&);
  void (*reset)(const KArgs&, const uint8_t*, int);
  void (*obs)(const KArgs&);
  void (*step_pipeline)(const KArgs&);
  void (*reset_pipeline)(const KArgs&, const uint8_t*);
  int A, O, G, S, scratch_words;
};
template <class T>
struct OpsT {
  static constexpr bool HAS_LIGHT = task_has_light<T>();
  static void init(const KArgs& a) { for (int64_t i = 0; i < a.n; i++) body_init<T>(a, i); }
  static void step(const KArgs& a) {  // the fused form
    for (int64_t i = 0; i < a.n; i++) {
      StepStats st = {0, 0, 0, 0, 0};
      body_step<T>(a, i, st);
      a.stats[0] += st.eps; a.stats[1] += st.ret; a.stats[2] += st.len; a.stats[3] += st.suc; a.stats[4] += st.div;
    }
  }
  static void reset(const KArgs& a, const uint8_t* mask, int use_flags) {
    for (int64_t i = 0; i < a.n; i++) {
      if (use_flags) { if (!a.need_reset[i]) continue; }
      else if (mask && !mask[i]) continue;
      body_reset<T>(a, i, !use_flags);
    }
  }
  static void obs(const KArgs& a) { for (int64_t i = 0; i < a.n; i++) body_obs<T>(a, i); }
  // ---- the kernel pipeline of the CUDA path, one "launch" = one loop over the envs (same order of operations as
  // OpsT<T>::step / ::reset in xarm_lib.cu)
  template <class F>
  static void launch(const KArgs& a, F f) {
    if (a.list) { for (int t = 0; t < *a.list_count; t++) f((int64_t)a.list[t]); }
    else for (int64_t i = 0; i < a.n; i++) f(i);
  }
  static void simulate(const KArgs& a) {
    for (int sub = 0; sub < T::NSUB; sub++) {
      if constexpr (!HAS_LIGHT) {
        launch(a, [&](int64_t i) { Contacts<T> C; pipe_heavy<T>(a, i, sub, C); });
      } else {
        int nh = 0;
        launch(a, [&](int64_t i) { if (pipe_setup<T>(a, i, sub)) a.heavy_list[nh++] = (int)i; });
        launch(a, [&](int64_t i) { float mrows[XARM_MROW_WORDS]; pipe_light<T>(a, i, mrows, 1); });
        for (int t = 0; t < nh; t++) { Contacts<T> C; pipe_heavy<T>(a, a.heavy_list[t], sub, C); }
      }
    }
  }
  static void reset_passes(const KArgs& r, bool clear_return) {
    if (reset_has_servo<T>())
      for (int rep = 0; rep < 5; rep++) { launch(r, [&](int64_t i) { pipe_reset_stage<T>(r, i, rep, false); }); simulate(r); }
    launch(r, [&](int64_t i) { pipe_reset_stage<T>(r, i, XARM_RESET_PLACE, false); });
    simulate(r);
    launch(r, [&](int64_t i) { pipe_reset_stage<T>(r, i, XARM_RESET_FINISH, clear_return); });
  }
  static void step_pipeline(const KArgs& a0) {
    KArgs a = a0; a.list = nullptr; a.list_count = nullptr;
    *a.reset_count = 0;
    launch(a, [&](int64_t i) { pipe_action<T>(a, i); });
    simulate(a);
    launch(a, [&](int64_t i) {
      StepStats st = {0, 0, 0, 0, 0};
      if (pipe_finish<T>(a, i, st)) a.reset_list[(*a.reset_count)++] = (int)i;
      a.stats[0] += st.eps; a.stats[1] += st.ret; a.stats[2] += st.len; a.stats[3] += st.suc; a.stats[4] += st.div;
    });
    if (a.auto_reset) { KArgs r = a; r.list = a.reset_list; r.list_count = a.reset_count; reset_passes(r, false); }
  }
  static void reset_pipeline(const KArgs& a0, const uint8_t* mask) {
    KArgs a = a0; a.list = nullptr; a.list_count = nullptr;
    *a.reset_count = 0;
    if (mask) {
      for (int64_t i = 0; i < a.n; i++) if (mask[i]) a.reset_list[(*a.reset_count)++] = (int)i;
      a.list = a.reset_list; a.list_count = a.reset_count;
    }
    reset_passes(a, true);
  }
  static Ops make() { Ops o = {init, step, reset, obs, step_pipeline, reset_pipeline, T::A, T::O, T::G, state_words<T>(), pipe_scratch_words<T>()}; return o; }
};
static bool get_ops(int task, Ops* out) {
  switch (task) {
    case XARM_TASK_REACH: *out = OpsT<TaskT<XARM_TASK_REACH, 0>>::make(); return true;
    case XARM_TASK_PICK_AND_PLACE: *out = OpsT<TaskT<XARM_TASK_PICK_AND_PLACE, 1>>::make(); return true;
    case XARM_TASK_STACK_TOWER: *out = OpsT<TaskT<XARM_TASK_STACK_TOWER, 3>>::make(); return true;
    case XARM_TASK_PUSH_WITH_DOOR: *out = OpsT<TaskT<XARM_TASK_PUSH_WITH_DOOR, 1>>::make(); return true;
    case XARM_TASK_HANDOVER: *out = OpsT<TaskT<XARM_TASK_HANDOVER, 1>>::make(); return true;
  }
  return false;
}

struct HS {
  XarmConfig cfg; Ops ops; KArgs k;
  std::vector<float> state, ep_return, io, scratch; std::vector<uint8_t> need_reset, flags; std::vector<int> ints; double stats[5];
  int pipeline = 0;  // 0: fused bodies (body_step / body_reset), 1: the kernel pipeline of the CUDA path
};

extern "C" {
HS* hs_create(int task, int reward_type, int num_obj, int goal_shape, double init_grasp_rate, double goal_ground_rate,
              double same_side_rate, int max_episode_steps, int auto_reset, long long num_envs, long long env_index_base,
              unsigned long long seed) {
  HS* h = new HS();
  XarmConfig c0; memset(&c0, 0, sizeof(c0));
  c0.task = task; c0.reward_type = reward_type; c0.num_obj = num_obj; c0.goal_shape = goal_shape;
  c0.init_grasp_rate = init_grasp_rate; c0.goal_ground_rate = goal_ground_rate; c0.same_side_rate = same_side_rate;
  c0.max_episode_steps = max_episode_steps; c0.auto_reset = auto_reset; c0.num_envs = num_envs; c0.env_index_base = env_index_base; c0.seed = seed;
  const XarmConfig* cfg = &c0;
  h->cfg = c0;
  if (!get_ops(cfg->task, &h->ops)) { delete h; return nullptr; }
  const int64_t n = cfg->num_envs; const Ops& o = h->ops;
  h->state.assign((size_t)n * o.S, 0.f); h->ep_return.assign(n, 0.f); h->need_reset.assign(n, 0); h->flags.assign(2 * n, 0);
  h->io.assign((size_t)n * (o.A + o.O + 2 * o.G + 2), 0.f);
  memset(h->stats, 0, sizeof(h->stats));
  KArgs& k = h->k; memset(&k, 0, sizeof(k));
  k.state = h->state.data(); k.ep_return = h->ep_return.data(); k.need_reset = h->need_reset.data(); k.stats = h->stats;
  k.n = n; k.auto_reset = cfg->auto_reset; k.heavy_dir = 1;
  h->scratch.assign((size_t)n * o.scratch_words, 0.f); h->ints.assign((size_t)4 * n + XARM_PIPE_COUNTERS + 4, 0);
  k.scratch = h->scratch.data(); k.reset_list = h->ints.data(); k.heavy_list = k.reset_list + n; k.form = k.reset_list + 2 * n;
  k.rng_draw = k.reset_list + 3 * n; k.heavy_count = k.reset_list + 4 * n; k.reset_count = k.heavy_count + XARM_PIPE_COUNTERS;
  k.rc.seed = cfg->seed; k.rc.env_index_base = cfg->env_index_base; k.rc.reward_type = cfg->reward_type;
  k.rc.goal_shape = cfg->goal_shape; k.rc.max_episode_steps = cfg->max_episode_steps;
  k.rc.init_grasp_rate = cfg->init_grasp_rate; k.rc.goal_ground_rate = cfg->goal_ground_rate; k.rc.same_side_rate = cfg->same_side_rate;
  k.rc.stagger = cfg->stagger_phases != 0;
  float* p = h->io.data();
  k.b.actions = p; p += n * o.A; k.b.observation = p; p += n * o.O; k.b.achieved_goal = p; p += n * o.G;
  k.b.desired_goal = p; p += n * o.G; k.b.reward = p; p += n; k.b.success = p; p += n;
  k.b.done = h->flags.data(); k.b.truncated = h->flags.data() + n; k.b.terminal_observation = nullptr;
  h->ops.init(k);
  return h;
}
void hs_destroy(HS* h) { delete h; }
void hs_set_pipeline(HS* h, int on) { h->pipeline = on; }
void hs_dims(HS* h, int* A, int* O, int* G, int* S) { *A = h->ops.A; *O = h->ops.O; *G = h->ops.G; *S = h->ops.S; }
static void copy_out(HS* h, float* obs, float* ag, float* dg) {
  const int64_t n = h->cfg.num_envs; const Ops& o = h->ops;
  if (obs) memcpy(obs, h->k.b.observation, sizeof(float) * n * o.O);
  if (ag) memcpy(ag, h->k.b.achieved_goal, sizeof(float) * n * o.G);
  if (dg) memcpy(dg, h->k.b.desired_goal, sizeof(float) * n * o.G);
}
void hs_reset(HS* h, const uint8_t* mask, float* obs, float* ag, float* dg) {
  if (h->pipeline) h->ops.reset_pipeline(h->k, mask); else h->ops.reset(h->k, mask, 0);
  copy_out(h, obs, ag, dg);
}
void hs_get_obs(HS* h, float* obs, float* ag, float* dg) { h->ops.obs(h->k); copy_out(h, obs, ag, dg); }
void hs_step(HS* h, const float* actions, float* obs, float* ag, float* dg, float* reward, uint8_t* done, float* success, uint8_t* truncated) {
  const int64_t n = h->cfg.num_envs; const Ops& o = h->ops;
  memcpy((void*)h->k.b.actions, actions, sizeof(float) * n * o.A);
  if (h->pipeline) h->ops.step_pipeline(h->k); else h->ops.step(h->k);  // fused: auto-reset inside the step body
  copy_out(h, obs, ag, dg);
  if (reward) memcpy(reward, h->k.b.reward, sizeof(float) * n);
  if (success) memcpy(success, h->k.b.success, sizeof(float) * n);
  if (done) memcpy(done, h->k.b.done, n);
  if (truncated) memcpy(truncated, h->k.b.truncated, n);
}
void hs_get_state(HS* h, float* out) {
  const int64_t n = h->cfg.num_envs; const int S = h->ops.S;
  for (int64_t i = 0; i < n; i++) for (int w = 0; w < S; w++) out[i * S + w] = h->state[(size_t)w * n + i];
}
// light / heavy classification the setup pass of the LAST substep left (pipeline mode): 1 = heavy
void hs_get_forms(HS* h, int* out) { for (int64_t i = 0; i < h->cfg.num_envs; i++) out[i] = h->k.form[i] & 0xff; }
void hs_set_state(HS* h, const float* in) {
  const int64_t n = h->cfg.num_envs; const int S = h->ops.S;
  for (int64_t i = 0; i < n; i++) for (int w = 0; w < S; w++) h->state[(size_t)w * n + i] = in[i * S + w];
}
void hs_compute_reward(int task, int reward_type, int num_obj, const float* ag, const float* dg, int64_t n, float* out) {
  int G = task == XARM_TASK_REACH ? 3 : (task == XARM_TASK_STACK_TOWER ? 9 : 3 * (task == XARM_TASK_PUSH_WITH_DOOR ? 1 : num_obj));
  for (int64_t i = 0; i < n; i++) out[i] = reward_stateless(task, reward_type, num_obj, task_threshold(task), ag + i * G, dg + i * G, G);
}
int hs_box_box(const float* A, const float* B, float* out) {
  Box a, b; CPoint c[4];
  a.c = v3(A[0], A[1], A[2]); for (int i = 0; i < 9; i++) a.R.m[i] = A[3 + i]; a.h = v3(A[12], A[13], A[14]);
  b.c = v3(B[0], B[1], B[2]); for (int i = 0; i < 9; i++) b.R.m[i] = B[3 + i]; b.h = v3(B[12], B[13], B[14]);
  int n = box_box(a, b, c, 4);
  for (int i = 0; i < n; i++) {
    float* o = out + 10 * i;
    o[0] = c[i].pa.x; o[1] = c[i].pa.y; o[2] = c[i].pa.z; o[3] = c[i].pb.x; o[4] = c[i].pb.y; o[5] = c[i].pb.z;
    o[6] = c[i].n.x; o[7] = c[i].n.y; o[8] = c[i].n.z; o[9] = c[i].depth;
  }
  return n;
}
}

// debug exports: joint-space inverse inertia and unconstrained velocity update of arm 0 (first substep semantics)
template <class T>
static int dyn_t(const float* q, const float* qd, int damping, float* Minv, float* qdu) {
  using MD = typename T::MD;
  ArmState<MD> st; ArmDyn<MD> D;
  for (int i = 0; i < MD::N; i++) { st.q[i] = q[i]; st.qd[i] = qd[i]; st.qt[i] = q[i]; }
  arm_dynamics<T>(0, st, damping != 0, D);
  for (int i = 0; i < MD::N; i++) { qdu[i] = D.qdu[i]; for (int j = 0; j < MD::N; j++) Minv[i * MD::N + j] = D.Minv[tri(i, j)]; }
  return MD::N;
}
extern "C" int hs_dynamics(int task, const float* q, const float* qd, int damping, float* Minv, float* qdu) {
  if (task == XARM_TASK_REACH) return dyn_t<TaskT<XARM_TASK_REACH, 0>>(q, qd, damping, Minv, qdu);
  return dyn_t<TaskT<XARM_TASK_PICK_AND_PLACE, 1>>(q, qd, damping, Minv, qdu);
}
template <class T>
static void ik_t(const float* q, const float* target, float* out) { arm_ik<T>(0, q, v3(target[0], target[1], target[2]), out); }
extern "C" void hs_ik(int task, const float* q, const float* target, float* out) {
  if (task == XARM_TASK_REACH) ik_t<TaskT<XARM_TASK_REACH, 0>>(q, target, out); else ik_t<TaskT<XARM_TASK_PICK_AND_PLACE, 1>>(q, target, out);
}
