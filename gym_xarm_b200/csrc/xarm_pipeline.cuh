// xarm_pipeline.cuh - Env.step as a pipeline of small kernels (per-env bodies; the __global__ wrappers are in
// xarm_lib.cu, tests/hostsim runs the same bodies on the host).
//
// Why: the fused one-kernel step (body_step) is ~270 KB of SASS and instruction-fetch bound - ncu shows 70 % of the
// warp cycles in "no instruction" stalls, 0.5 IPC per SM and 8 warps per SM at 255 registers (profiles/).  Each piece
// below fits the 32 KB L1.5 instruction cache, needs far fewer registers, and all warps of a launch run the same loop:
//
//   pipe_action            _set_action: FK, IK, motor targets (once per env step)
//   NSUB x { pipe_setup    collide -> unconstrained velocities -> rows; classifies the env as light / heavy
//            pipe_light    light envs : PGS over the arm rows + one object manifold, integrate      } run concurrently
//            pipe_heavy    heavy envs : the whole substep with the generic coupled solver             } (graph fork)
//   pipe_finish            _get_obs, reward, success, done, TimeLimit, episode statistics, auto-reset list
//
// Intermediates travel through a per-handle scratch slab in the same [word][env] layout as the state (coalesced; it
// stays in the 126 MB L2 at the benchmark sizes): 72 + 22 + 28 words per env and substep for PickAndPlace.
// Every body takes the env index; with a list (auto-reset tail) thread t works on env list[t].
#pragma once
#include "xarm_kernels.cuh"

#define XARM_FORM_LIGHT 0
#define XARM_FORM_HEAVY 1

template <class T>
XHD int pipe_ar_words() { return ar_words<T>(); }
template <class T>
XHD int pipe_sb_words() { return T::NARM * T::MD::N + 6 * (T::NOBJ > 0 ? T::NOBJ : 1) + 1; }
XHD int pipe_mi_words() { return mi_words(); }
template <class T>
XHD int pipe_dyn_words() { return T::MD::N * 6 + T::MD::N * (T::MD::N + 1) / 2 + T::MD::N + 9 + 9; }  // ArmDyn of a heavy env
template <class T>
XHD int pipe_scratch_words() {
  const int light = pipe_ar_words<T>() + pipe_sb_words<T>() + pipe_mi_words() * (T::NOBJ > 0 ? T::NOBJ : 1), heavy = pipe_dyn_words<T>();
  return light > heavy ? light : heavy;
}

// ---- scratch records: word w of env i at base[w * n + i] (ar_store / ar_load / mi_store / mi_load: xarm_sim.cuh)
template <class T>
XD void sb_store(const SubBase<T>& B, float* __restrict__ s, int64_t n, int64_t i) {
  int w = 0;
#pragma unroll
  for (int a = 0; a < T::NARM; a++)
#pragma unroll
    for (int k = 0; k < T::MD::N; k++) s[(w++) * n + i] = B.qdu[a][k];
#pragma unroll
  for (int o = 0; o < (T::NOBJ > 0 ? T::NOBJ : 1); o++) {
    s[(w++) * n + i] = B.vu[o].x; s[(w++) * n + i] = B.vu[o].y; s[(w++) * n + i] = B.vu[o].z;
    s[(w++) * n + i] = B.wu[o].x; s[(w++) * n + i] = B.wu[o].y; s[(w++) * n + i] = B.wu[o].z;
  }
  s[(w++) * n + i] = B.door_qdu;
}
template <class T>
XD void sb_load(SubBase<T>& B, const float* __restrict__ s, int64_t n, int64_t i) {
  int w = 0;
#pragma unroll
  for (int a = 0; a < T::NARM; a++)
#pragma unroll
    for (int k = 0; k < T::MD::N; k++) B.qdu[a][k] = s[(w++) * n + i];
#pragma unroll
  for (int o = 0; o < (T::NOBJ > 0 ? T::NOBJ : 1); o++) {
    B.vu[o].x = s[(w++) * n + i]; B.vu[o].y = s[(w++) * n + i]; B.vu[o].z = s[(w++) * n + i];
    B.wu[o].x = s[(w++) * n + i]; B.wu[o].y = s[(w++) * n + i]; B.wu[o].z = s[(w++) * n + i];
  }
  B.door_qdu = s[(w++) * n + i];
}
// the dynamics pass of a heavy env (joint subspaces, inverse inertia, unconstrained velocities, gripper frames): the
// heavy path reuses it instead of running arm_dynamics again
template <class T>
XD void dyn_store(const ArmDyn<typename T::MD>& D, float* __restrict__ s, int64_t n, int64_t i) {
  constexpr int N = T::MD::N, NT = N * (N + 1) / 2;
  int w = 0;
#pragma unroll
  for (int k = 0; k < N; k++) {
    s[(w++) * n + i] = D.S[k].a.x; s[(w++) * n + i] = D.S[k].a.y; s[(w++) * n + i] = D.S[k].a.z;
    s[(w++) * n + i] = D.S[k].l.x; s[(w++) * n + i] = D.S[k].l.y; s[(w++) * n + i] = D.S[k].l.z;
  }
#pragma unroll
  for (int k = 0; k < NT; k++) s[(w++) * n + i] = D.Minv[k];
#pragma unroll
  for (int k = 0; k < N; k++) s[(w++) * n + i] = D.qdu[k];
#pragma unroll
  for (int k = 0; k < 9; k++) s[(w++) * n + i] = D.Rh.m[k];
  s[(w++) * n + i] = D.ph.x; s[(w++) * n + i] = D.ph.y; s[(w++) * n + i] = D.ph.z;
  s[(w++) * n + i] = D.pf1.x; s[(w++) * n + i] = D.pf1.y; s[(w++) * n + i] = D.pf1.z;
  s[(w++) * n + i] = D.pf2.x; s[(w++) * n + i] = D.pf2.y; s[(w++) * n + i] = D.pf2.z;
}
template <class T>
XD void dyn_load(ArmDyn<typename T::MD>& D, const float* __restrict__ s, int64_t n, int64_t i) {
  constexpr int N = T::MD::N, NT = N * (N + 1) / 2;
  int w = 0;
#pragma unroll
  for (int k = 0; k < N; k++) {
    D.S[k].a.x = s[(w++) * n + i]; D.S[k].a.y = s[(w++) * n + i]; D.S[k].a.z = s[(w++) * n + i];
    D.S[k].l.x = s[(w++) * n + i]; D.S[k].l.y = s[(w++) * n + i]; D.S[k].l.z = s[(w++) * n + i];
  }
#pragma unroll
  for (int k = 0; k < NT; k++) D.Minv[k] = s[(w++) * n + i];
#pragma unroll
  for (int k = 0; k < N; k++) D.qdu[k] = s[(w++) * n + i];
#pragma unroll
  for (int k = 0; k < 9; k++) D.Rh.m[k] = s[(w++) * n + i];
  D.ph.x = s[(w++) * n + i]; D.ph.y = s[(w++) * n + i]; D.ph.z = s[(w++) * n + i];
  D.pf1.x = s[(w++) * n + i]; D.pf1.y = s[(w++) * n + i]; D.pf1.z = s[(w++) * n + i];
  D.pf2.x = s[(w++) * n + i]; D.pf2.y = s[(w++) * n + i]; D.pf2.z = s[(w++) * n + i];
}

// ---- _set_action (once per env step): FK, IK, motor targets; Handover's lego clamp
template <class T>
XD void pipe_action(const KArgs& a, int64_t i) {
  Env<T> e;
  env_load<T>(e, a.state, a.n, i);
  float act[T::A];
#pragma unroll
  for (int k = 0; k < T::A; k++) act[k] = a.b.actions[i * T::A + k];
  e.step_count += 1;
  set_action<T>(e, act);
  env_store<T>(e, a.state, a.n, i);
}

// ---- substep, part 1: rows.  Light envs leave their rows in the scratch slab; the others join the heavy list.
// Returns true when env i is heavy (the wrapper appends it to the list, warp-aggregated).
template <class T, bool FLAT = false>
XD bool pipe_setup(const KArgs& a, int64_t i, int sub) {
  Env<T> e;
  env_load<T>(e, a.state, a.n, i);
  ArmRows<T> AR;
  SubBase<T> B;
  const bool last = sub == T::NSUB - 1;
  if constexpr (task_single_island_pair<T>()) {
    ManifoldIn MI;
    const int g0 = e.grasp[0];
    int nc = 0;
    ArmDyn<typename T::MD> D[1];
    // heavy: a gripper link touches the object - or (rare) an arm joint sits on a limit: the light solver keeps no rows for those
    if (!sub_setup_lean<T, FLAT>(e, T::DAMP_EACH || sub == 0, last, AR, B, MI, nc, D) || arm_joint_on_limit<T>(AR)) {
      a.form[i] = XARM_FORM_HEAVY;
      dyn_store<T>(D[0], a.scratch, a.n, i);  // the heavy path continues from this dynamics pass
      return true;
    }
    float* s = a.scratch;
    ar_store<T>(AR, s, a.n, i); s += (int64_t)pipe_ar_words<T>() * a.n;
    sb_store<T>(B, s, a.n, i); s += (int64_t)pipe_sb_words<T>() * a.n;
    if (nc > 0) mi_store(MI, s, a.n, i);
    a.form[i] = XARM_FORM_LIGHT | (nc << 8);
    if (last && e.grasp[0] != g0) a.state[(int64_t)(state_words<T>() - 2) * a.n + i] = grasp_word(e.grasp[0], e.grasp_cmd[0]);  // grasp flag of the last collision pass
    return false;
  } else {
    // several islands (two arms, several objects, the door): heavy = any contact outside "object on one static box"
    constexpr int NO = T::NOBJ > 0 ? T::NOBJ : 1;
    ManifoldIn MI[NO];
    int nc[NO];
    ArmDyn<typename T::MD> D[T::NARM];
    int g0[2] = {e.grasp[0], e.grasp[1]};
    if (!sub_setup_lean_multi<T, FLAT>(e, T::DAMP_EACH || sub == 0, last, AR, B, MI, nc, D) || arm_joint_on_limit<T>(AR)) {
      a.form[i] = XARM_FORM_HEAVY;   // (the generic heavy path makes its own dynamics pass)
      return true;
    }
    float* s = a.scratch;
    ar_store<T>(AR, s, a.n, i); s += (int64_t)pipe_ar_words<T>() * a.n;
    sb_store<T>(B, s, a.n, i); s += (int64_t)pipe_sb_words<T>() * a.n;
    int packed = 0;
#pragma unroll
    for (int o = 0; o < T::NOBJ; o++) {
      if (nc[o] > 0) mi_store(MI[o], s + (int64_t)(o * pipe_mi_words()) * a.n, a.n, i);
      packed |= nc[o] << (3 * o);
    }
    a.form[i] = XARM_FORM_LIGHT | (packed << 8);
    if (last) {
#pragma unroll
      for (int arm = 0; arm < T::NARM; arm++)
        if (e.grasp[arm] != g0[arm]) a.state[(int64_t)(state_words<T>() - 2 + arm) * a.n + i] = grasp_word(e.grasp[arm], e.grasp_cmd[arm]);
    }
    return false;
  }
}

// ---- substep, part 2 (light envs): PGS over the arm rows and the object's manifold, then stepPositionsMultiDof.
// mrows: per-thread manifold rows, word w at mrows[w * stride] (shared memory in the kernel)
template <class T>
XD void pipe_light(const KArgs& a, int64_t i, float* mrows, int stride) {
  const int f = a.form[i];
  if ((f & 0xff) != XARM_FORM_LIGHT) return;
  SubBase<T> B;
  SubSol<T> S;
  const float* s = a.scratch;
  if constexpr (task_single_island_pair<T>()) {
    const int nc = f >> 8;
    ArmRows<T> AR;
    ManifoldIn MI;
    ar_load<T>(AR, s, a.n, i); s += (int64_t)pipe_ar_words<T>() * a.n;
    sb_load<T>(B, s, a.n, i); s += (int64_t)pipe_sb_words<T>() * a.n;
    if (nc > 0) mi_load(MI, s, a.n, i);
    sub_solve_light<T>(AR, nc, MI, mrows, stride, S);
  } else {
    constexpr int NO = T::NOBJ > 0 ? T::NOBJ : 1;
    int nc[NO];
#pragma unroll
    for (int o = 0; o < NO; o++) nc[o] = (f >> (8 + 3 * o)) & 7;
    const float* s_sb = s + (int64_t)pipe_ar_words<T>() * a.n;
    const float* s_mi = s_sb + (int64_t)pipe_sb_words<T>() * a.n;
    sub_solve_light_multi<T>(s, s_mi, a.n, i, nc, mrows, stride, S);
    sb_load<T>(B, s_sb, a.n, i);
  }
  Env<T> e;
  env_load_dyn<T>(e, a.state, a.n, i);
  sub_integrate<T>(e, B, S);
  env_store_dyn<T>(e, a.state, a.n, i);
}

// ---- substep for heavy envs (gripper contacts, several manifolds, two arms, door): setup + generic coupled PGS + integrate
template <class T>
XD void pipe_heavy(const KArgs& a, int64_t i, int sub, Contacts<T>& C) {
  Env<T> e;
  env_load<T>(e, a.state, a.n, i);
  substep_generic<T>(e, T::DAMP_EACH || sub == 0, sub == T::NSUB - 1, C);
  env_store<T>(e, a.state, a.n, i);
}

// ---- end of the env step: _get_obs, _is_success, compute_reward, done / TimeLimit, statistics, auto-reset bookkeeping
// Returns true when the env finished and auto-reset is on (the wrapper appends it to the reset list).
template <class T>
XD bool pipe_finish(const KArgs& a, int64_t i, StepStats& st) {
  Env<T> e;
  env_load<T>(e, a.state, a.n, i);
  Obs<T> o;
  StepOut so;
  env_step_outputs<T>(e, a.rc, o, so);
  bool rebuilt = false;
  if (!env_finite<T>(e)) {  // NaN guard: rebuild the env, end the episode
    uint32_t ep = e.episode;
    env_construct<T>(e, a.rc, a.rc.env_index_base + i);
    e.episode = ep;
    get_obs<T>(e, o);
    so.reward = 0.f; so.success = 0.f; so.done = true; so.truncated = true;
    st.div = 1.f;
    rebuilt = true;
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
  if (rebuilt || (T::TASK == XARM_TASK_REACH && a.rc.reward_type == XARM_REWARD_DENSE_DIFF)) env_store<T>(e, a.state, a.n, i);  // d_old / rebuilt state
  return so.done && a.auto_reset;
}

// ---- which envs may finish in the coming step (time limit reached, or success within reach): the "early" branch of
// xarm_step runs them first, so that their auto-reset passes overlap the step of all the other envs.  A wrong "no"
// is only slower (the env is reset by the late tail after both branches), never wrong.
template <class T>
XD bool pipe_may_finish(const KArgs& a, int64_t i) {
  if (!a.auto_reset) return false;
  const int64_t n = a.n;
  const int w_goal = state_words<T>() - 5 - T::G;
  const int step_count = (int)a.state[(int64_t)(w_goal + T::G) * n + i];
  const int limit = a.rc.max_episode_steps > 0 ? a.rc.max_episode_steps : T::MAX_STEPS;
  if (step_count + 1 >= limit) return true;
  if (T::TASK == XARM_TASK_PICK_AND_PLACE || T::TASK == XARM_TASK_HANDOVER) {  // success ends the episode
    const int w_obj = T::NARM * 3 * T::MD::N;
    bool all_near = true;
    for (int o = 0; o < T::NOBJ; o++) {
      float d2 = 0.f, v2 = 0.f;
      for (int c = 0; c < 3; c++) {
        const float d = a.state[(int64_t)(w_obj + 13 * o + c) * n + i] - a.state[(int64_t)(w_goal + 3 * o + c) * n + i];
        const float v = a.state[(int64_t)(w_obj + 13 * o + 7 + c) * n + i];
        d2 += d * d; v2 += v * v;
      }
      // an object no gripper touches travels |v| dt (+ what gravity adds while it falls); one in the gripper travels with it:
      // at most max_vel * dt per axis and step
      const bool held = (a.form[i] & 0xff) != XARM_FORM_LIGHT;
      const float dt_step = (float)(T::TIME_STEP * (T::TASK == XARM_TASK_HANDOVER ? T::NSUB : 1));
      // (generous factors: a wrong "no" costs a whole serial reset tail after both branches)
      const float free_travel = 3.f * sqrtf(v2) * dt_step + (float)XARM_GRAVITY * dt_step * dt_step + 0.02f;
      const float held_travel = 3.f * (float)(T::MAX_VEL * T::DT_CMD) + 0.02f;
      float reach = T::THRESHOLD + (held ? held_travel : fminf(free_travel, held_travel));
      if (T::TASK == XARM_TASK_PICK_AND_PLACE && a.b.observation) {
        // Tighter, by the hand - lego distance the last step (or reset) left in the caller's observation buffer
        // [REF xarm_pick_and_place.py:224-239: obs[8 + 16 o + 13 ..] = lego pos - hand pos].  The commanded hand target moves at
        // most max_vel * dt per axis and step (0.108 m along the diagonal; measured over 12 766 success finishers of 100 staggered
        // steps at 131 072 envs: lego travel <= 0.0955 m, tools/predictor_stats.py), so: a lego within H1 of the hand may be carried
        // R; a lego farther than H2 from the hand cannot be touched in this step and flies free; in between the old generous
        // bound.  (None of the finishers was missed by this rule in that run; candidates 9.6 k -> 7.2 k per step, and those in
        // gripper contact - the cooperative solver's load in the early branch's first pass - 2.06 k -> ~1 k.)  A stale or
        // overwritten buffer only costs a late tail, never a wrong result.
        const float* ob = a.b.observation + i * T::O + 8 + 16 * o + 13;
        const float hd2 = ob[0] * ob[0] + ob[1] * ob[1] + ob[2] * ob[2];
        const float R = 0.11f, H1 = 0.13f, H2 = 0.24f;
        // between H1 and H2 the hand can still reach the lego within the step and push it for the rest of its travel: at most
        // ~(H2 - hd).  Round 2 end (tools/late_tail_diag.py, 200 staggered steps): the ~15 finishers the predictor missed were (a)
        // legos just outside H1 at rest (d 0.075-0.079 against a free-travel reach of 0.073-0.076) and (b) legos INSIDE H1 that no
        // gripper link touched yet - their reach had stayed the free-travel bound because it was min()-ed with the untouched-lego bound
        // above; a lego inside H1 may be carried R whether or not it is touched now.
        const float hd = sqrtf(hd2);
        const float mid = fminf(fmaxf(free_travel, H2 - hd + 0.01f), R), far = fminf(1.5f * sqrtf(v2) * dt_step + (float)XARM_GRAVITY * dt_step * dt_step + 0.003f, R);
        const float tight = hd2 < H1 * H1 ? R : (hd2 > H2 * H2 ? far : mid);   // (NaN: mid)
        reach = hd2 < H1 * H1 ? T::THRESHOLD + R : fminf(reach, T::THRESHOLD + tight);
      }
      all_near = all_near && d2 < reach * reach;
    }
    return all_near;
  }
  return false;
}

// ---- Env.reset() through the same pipeline (auto-reset tail and xarm_reset): env_reset() cut at its stepSimulation
// calls.  Stages: 0..4 = the five servo repetitions of PickAndPlace / Handover (IK to the start pose, finger command;
// stage 0 also opens the episode), XARM_RESET_PLACE = object placement (teleport tasks: also joint teleport and episode
// bookkeeping), XARM_RESET_FINISH = _sample_goal, _get_obs, d_old.  A simulate() pass follows every stage but the last.
// The Philox draw counter travels from PLACE to FINISH in a.rng_draw (Handover's placement loop consumes a variable
// number of draws).  [REF xarm_pick_and_place.py:250-287; xarm_handover.py:338-393; xarm_reach.py:163-173;
// xarm_stack_tower.py:201-219; xarm_push_with_door.py:195-212]
#define XARM_RESET_PLACE 5
#define XARM_RESET_FINISH 6
template <class T>
XHD bool reset_has_servo() { return T::TASK == XARM_TASK_PICK_AND_PLACE || T::TASK == XARM_TASK_HANDOVER; }

template <class T>
XD void pipe_reset_stage(const KArgs& a, int64_t i, int stage, bool clear_return) {
  using MD = typename T::MD;
  Env<T> e;
  env_load<T>(e, a.state, a.n, i);
  const int64_t genv = a.rc.env_index_base + i;
  if (stage < XARM_RESET_PLACE) {  // servo repetition
    if (stage == 0) { e.episode += 1; e.step_count = 0; }
    const bool pp = T::TASK == XARM_TASK_PICK_AND_PLACE;
    const V3 t0 = pp ? v3(0.4f, 0.f, 0.12f) : v3(-0.15f, 0.f, 0.15f), t1 = pp ? v3(0, 0, 0) : v3(0.15f, 0.f, 0.15f);
#pragma unroll
    for (int arm = 0; arm < T::NARM; arm++) {
      float qn[7];
      arm_ik<T>(arm, e.arm[arm].q, arm == 0 ? t0 : t1, qn);
#pragma unroll
      for (int k = 0; k < 7; k++) e.arm[arm].qt[k] = qn[k];
      if (pp && MD::F2 >= 0) { e.arm[arm].qt[MD::F1] = 0.02f; e.arm[arm].qt[MD::F2 < 0 ? 0 : MD::F2] = 0.02f; }
    }
    env_store<T>(e, a.state, a.n, i);
    return;
  }
  if (stage == XARM_RESET_PLACE) {
    if (!reset_has_servo<T>()) { e.episode += 1; e.step_count = 0; }
    Rng rng = {a.rc.seed, (uint64_t)genv, e.episode, 0u};
    env_reset_place<T>(e, rng, a.rc);
    a.rng_draw[i] = (int)rng.draw;
    env_store<T>(e, a.state, a.n, i);
    return;
  }
  Rng rng = {a.rc.seed, (uint64_t)genv, e.episode, (uint32_t)a.rng_draw[i]};
  sample_goal<T>(e, rng, a.rc);
  Obs<T> o;
  get_obs<T>(e, o);
  e.d_old = np_dist(o.ag, o.dg, T::G);  // [REF xarm_reach.py:100]
  if (clear_return) e.step_count = initial_step_count<T>(a.rc, genv);   // explicit reset (xarm_reset): staggered phases, if configured
  write_obs<T>(a, i, o);
  env_store<T>(e, a.state, a.n, i);
  a.need_reset[i] = 0;
  if (clear_return) a.ep_return[i] = 0.f;
}
