// xarm_env.cuh - gym-level logic of the five tasks on the device: _set_action, _get_obs, compute_reward,
// _is_success, done, _reset_sim, _sample_goal.  Each function cites the reference method it replaces.
#pragma once
#include "xarm_sim.cuh"

// ------------------------------------------------------------------------------------------------ rewards (K6)
// Bit-exact float32 evaluation of the reference's NumPy expressions: every operation individually rounded
// (no FMA contraction), np.add.reduce order (left-to-right below 8 terms, 8-way pairwise otherwise).
XD float np_sum_sq_diff(const float* a, const float* b, int n) {
  if (n < 8) {
    float d0 = __fsub_rn(a[0], b[0]);
    float s = __fmul_rn(d0, d0);
    for (int i = 1; i < n; i++) { float d = __fsub_rn(a[i], b[i]); s = __fadd_rn(s, __fmul_rn(d, d)); }
    return s;
  }
  float r[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { float d = __fsub_rn(a[i], b[i]); r[i] = __fmul_rn(d, d); }
  int i = 8;
  for (; i + 8 <= n; i += 8)
#pragma unroll
    for (int j = 0; j < 8; j++) { float d = __fsub_rn(a[i + j], b[i + j]); r[j] = __fadd_rn(r[j], __fmul_rn(d, d)); }
  float s = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])), __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; i++) { float d = __fsub_rn(a[i], b[i]); s = __fadd_rn(s, __fmul_rn(d, d)); }
  return s;
}
XD float np_dist(const float* a, const float* b, int n) { return __fsqrt_rn(np_sum_sq_diff(a, b, n)); }
// IEEE negation (sign-bit flip, -(+0) = -0).  nvcc 12.9 / ptxas turn select(c, -1.0f, -0.0f) - however it is spelled:
// -(c ? 1.f : 0.f), an xor with 0x80000000, an explicit select of the two constants - into I2FP(c ? -1 : 0), which
// yields +0.0 and loses the reference's -0.0 sparse reward.  The sign bit therefore comes from constant
// memory, which the compiler cannot fold.
#if defined(__CUDACC__) && !defined(XARM_HOST_SIM)
__constant__ float c_neg_zero = -0.0f;
#endif
XD float neg_exact(float x) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(__float_as_int(x) ^ __float_as_int(c_neg_zero));
#else
  return -x;
#endif
}

// state-free rewards (the ones SB3's HER may call with batches)
XD float reward_stateless(int task, int reward_type, int num_obj, float thr, const float* ag, const float* dg, int G) {
  switch (task) {
    case XARM_TASK_REACH: {  // [REF xarm_reach.py:107-112]
      float d = np_dist(ag, dg, 3);
      return reward_type == XARM_REWARD_SPARSE ? (d < thr ? 1.0f : 0.0f) : neg_exact(d);
    }
    case XARM_TASK_PICK_AND_PLACE: {  // [REF xarm_pick_and_place.py:163-165,176-177]
      float d = np_dist(ag, dg, G);
      return reward_type == XARM_REWARD_SPARSE ? (d < thr ? 1.0f : 0.0f) : neg_exact(d);
    }
    case XARM_TASK_STACK_TOWER:
    case XARM_TASK_PUSH_WITH_DOOR: {  // [REF xarm_stack_tower.py:124-129]: -(d > thr) as float32 => -1.0 or -0.0
      float d = np_dist(ag, dg, G);
      return reward_type == XARM_REWARD_SPARSE ? neg_exact(d > thr ? 1.0f : 0.0f) : neg_exact(d);
    }
    default: {  // Handover [REF xarm_handover.py:177-183]
      float s = 0.f;
      for (int i = 0; i < num_obj; i++) s = __fadd_rn(s, np_dist(ag + 3 * i, dg + 3 * i, 3) > thr ? 1.0f : 0.0f);
      return neg_exact(s);
    }
  }
}
XD float task_threshold(int task) {
  return task == XARM_TASK_STACK_TOWER ? (float)(0.03 * 3) : (task == XARM_TASK_PUSH_WITH_DOOR ? (float)(0.03 * 1) : 0.05f);
}

// ------------------------------------------------------------------------------------------------ _get_obs
template <class T>
struct Obs {
  float obs[T::O];
  float ag[T::G];
  float dg[T::G];
};

template <class T>
XD void get_obs(const Env<T>& e, Obs<T>& o) {
  using MD = typename T::MD;
  V3 hp[T::NARM], hv[T::NARM];
#pragma unroll
  for (int a = 0; a < T::NARM; a++) hand_state<T>(a, e.arm[a], hp[a], hv[a]);
  int n = 0;
  if (T::TASK == XARM_TASK_REACH || T::TASK == XARM_TASK_PICK_AND_PLACE) {
    // [REF xarm_reach.py:144-161] / [REF xarm_pick_and_place.py:220-248]
    o.obs[n++] = hp[0].x; o.obs[n++] = hp[0].y; o.obs[n++] = hp[0].z;
    o.obs[n++] = hv[0].x; o.obs[n++] = hv[0].y; o.obs[n++] = hv[0].z;
    o.obs[n++] = e.arm[0].q[MD::F1]; o.obs[n++] = e.arm[0].qd[MD::F1];
    if (T::TASK == XARM_TASK_REACH) {
      o.ag[0] = hp[0].x; o.ag[1] = hp[0].y; o.ag[2] = hp[0].z;
    } else {
#pragma unroll
      for (int i = 0; i < T::NOBJ; i++) {
        const ObjState& b = e.obj[i];
        o.obs[n++] = b.pos.x; o.obs[n++] = b.pos.y; o.obs[n++] = b.pos.z;
        o.obs[n++] = b.quat.x; o.obs[n++] = b.quat.y; o.obs[n++] = b.quat.z; o.obs[n++] = b.quat.w;
        o.obs[n++] = b.v.x - hv[0].x; o.obs[n++] = b.v.y - hv[0].y; o.obs[n++] = b.v.z - hv[0].z;
        o.obs[n++] = b.w.x; o.obs[n++] = b.w.y; o.obs[n++] = b.w.z;
        o.obs[n++] = b.pos.x - hp[0].x; o.obs[n++] = b.pos.y - hp[0].y; o.obs[n++] = b.pos.z - hp[0].z;
        o.ag[3 * i] = b.pos.x; o.ag[3 * i + 1] = b.pos.y; o.ag[3 * i + 2] = b.pos.z;
      }
    }
  } else {
    // [REF xarm_stack_tower.py:164-199; xarm_push_with_door.py:162-193; xarm_handover.py:299-336]
#pragma unroll
    for (int i = 0; i < T::NOBJ; i++) { o.obs[n++] = e.obj[i].pos.x; o.obs[n++] = e.obj[i].pos.y; o.obs[n++] = e.obj[i].pos.z; }
#pragma unroll
    for (int i = 0; i < T::NOBJ; i++) { o.obs[n++] = e.obj[i].quat.x; o.obs[n++] = e.obj[i].quat.y; o.obs[n++] = e.obj[i].quat.z; o.obs[n++] = e.obj[i].quat.w; }
#pragma unroll
    for (int i = 0; i < T::NOBJ; i++) { o.obs[n++] = e.obj[i].v.x; o.obs[n++] = e.obj[i].v.y; o.obs[n++] = e.obj[i].v.z; }
#pragma unroll
    for (int i = 0; i < T::NOBJ; i++) { o.obs[n++] = e.obj[i].w.x; o.obs[n++] = e.obj[i].w.y; o.obs[n++] = e.obj[i].w.z; }
#pragma unroll
    for (int a = 0; a < T::NARM; a++) {
      o.obs[n++] = hp[a].x; o.obs[n++] = hp[a].y; o.obs[n++] = hp[a].z - T::HAND_OFF_Z;
      o.obs[n++] = hv[a].x; o.obs[n++] = hv[a].y; o.obs[n++] = hv[a].z;
      if (T::TASK != XARM_TASK_PUSH_WITH_DOOR) { o.obs[n++] = e.arm[a].q[MD::F1]; o.obs[n++] = e.arm[a].qd[MD::F1]; }
    }
#pragma unroll
    for (int i = 0; i < T::NOBJ; i++) { o.ag[3 * i] = e.obj[i].pos.x; o.ag[3 * i + 1] = e.obj[i].pos.y; o.ag[3 * i + 2] = e.obj[i].pos.z; }
  }
#pragma unroll
  for (int c = 0; c < T::G; c++) o.dg[c] = e.goal[c];
}

// staged dense rewards that read the live simulator [REF xarm_pick_and_place.py:166-175; xarm_handover.py:185-199 (D2)]
template <class T>
XD float reward_dense_staged(const Env<T>& e, const Obs<T>& o) {
  const float z3[3] = {0.f, 0.f, 0.f};
  if (T::TASK == XARM_TASK_PICK_AND_PLACE) {
    float d3[3] = {o.obs[0] - o.ag[0] + 0.06f, o.obs[1] - o.ag[1], (o.obs[2] - (float)(0.088 - 0.021)) - o.ag[2]};
    float d_ao = np_dist(d3, z3, 3), d_og = np_dist(o.ag, o.dg, T::G);
    if (!e.grasp[0]) return 0.25f * (1.f - tanhf(d_ao));
    if (o.ag[2] > 0.05f) return 1.0f + 0.25f * (1.f - tanhf(d_og));
    return 0.5f;
  }
  const int n0 = 13 * T::NOBJ;
  float p1[3] = {o.obs[n0] - o.ag[0] + 0.06f, o.obs[n0 + 1] - o.ag[1], o.obs[n0 + 2] - o.ag[2]};
  float p2[3] = {o.obs[n0 + 8] - o.ag[0] - 0.06f, o.obs[n0 + 9] - o.ag[1], o.obs[n0 + 10] - o.ag[2]};
  float d1 = np_dist(p1, z3, 3), d2 = np_dist(p2, z3, 3);
  // if_xarm1_grasp / if_xarm2_grasp are set by _set_action, BEFORE the 15 stepSimulation calls, and compute_reward reads
  // those values [REF xarm_handover.py:262-263, 185-199] (PickAndPlace above queries getContactPoints afresh)
  const bool g1 = e.grasp_cmd[0], g2 = e.grasp_cmd[1];
  if (!g1 && !g2) return 0.25f * (1.f - tanhf(d1)) / 2.25f;
  if (g1 && !g2) return o.ag[2] > 0.05f ? (1.0f + 0.25f * (1.f - tanhf(d2))) / 2.25f : 0.5f / 2.25f;
  if (g1 && g2) return 1.5f / 2.25f;
  return (2.0f + 0.25f * (1.f - tanhf(np_dist(o.ag, o.dg, T::G)))) / 2.25f;
}

// ------------------------------------------------------------------------------------------------ _set_action
template <class T>
XD void set_action(Env<T>& e, const float* act_in) {
  using MD = typename T::MD;
  float act[T::A];
#pragma unroll
  for (int i = 0; i < T::A; i++) act[i] = fminf(1.f, fmaxf(-1.f, act_in[i]));  // np.clip(action, -1, 1)
  e.grasp_cmd[0] = e.grasp[0]; e.grasp_cmd[1] = e.grasp[1];  // the contact flags _set_action reads (friction switch, Handover's reward stages)
#pragma unroll
  for (int a = 0; a < T::NARM; a++) {
    const float* u = T::TASK == XARM_TASK_PUSH_WITH_DOOR ? act + 3 * a : act + 4 * a;
    M3 Re; V3 pe, org[7], axs[7];
    arm_fk7<T>(a, e.arm[a].q, Re, pe, org, axs);  // getLinkState(arm, 8)[0]
    // `a[:3] * max_vel * dt` is float32 array arithmetic in the reference: two separately rounded products
    const float mv = (float)T::MAX_VEL, dtc = (float)T::DT_CMD;
    V3 target;
    target.x = fminf(T::pos_hi(a, 0), fmaxf(T::pos_lo(a, 0), pe.x + __fmul_rn(__fmul_rn(u[0], mv), dtc)));
    target.y = fminf(T::pos_hi(a, 1), fmaxf(T::pos_lo(a, 1), pe.y + __fmul_rn(__fmul_rn(u[1], mv), dtc)));
    target.z = fminf(T::pos_hi(a, 2), fmaxf(T::pos_lo(a, 2), pe.z + __fmul_rn(__fmul_rn(u[2], mv), dtc)));
    float qn[7];
    arm_ik<T>(a, e.arm[a].q, target, qn);
#pragma unroll
    for (int i = 0; i < 7; i++) e.arm[a].qt[i] = qn[i];
    if (T::GRIP_CMD) {
      float g = e.arm[a].q[MD::F1] + u[3] * (float)(T::DT_CMD * T::MAX_GRIP_VEL);
      if (T::GRIP_CLIP) g = fminf(T::GRIP_HI, fmaxf(T::GRIP_LO, g));
      if (MD::F2 >= 0) { e.arm[a].qt[MD::F1] = g; e.arm[a].qt[MD::F2 < 0 ? 0 : MD::F2] = g; }
      else {
#pragma unroll
        for (int i = MD::F1; i < MD::N; i++) e.arm[a].qt[i] = g;  // Reach drives joints 10..16 [REF xarm_reach.py:141-142]
      }
    }
  }
  if (T::LEGO_CLAMP) {  // [REF xarm_handover.py:282-297] clamp xy, keep only pitch, velocity reset to zero
#pragma unroll
    for (int i = 0; i < T::NOBJ; i++) {
      ObjState& b = e.obj[i];
      float x = b.pos.x, y = b.pos.y;
      const bool neg = x < 0.f;
      if (neg) x = -x;
      x = fminf(0.28f, fmaxf(-0.28f, x)); y = fminf(0.2f, fmaxf(-0.2f, y));
      if (neg) x = -x;
      M3 R = quat_to_m3(b.quat);
      float sp = fminf(1.f, fmaxf(-1.f, -R.m[6]));
      float pitch = asinf(sp), s, c;
      sincosf(0.5f * pitch, &s, &c);
      b.pos.x = x; b.pos.y = y;
      b.quat.x = 0.f; b.quat.y = s; b.quat.z = 0.f; b.quat.w = c;
      b.v = v3(0, 0, 0); b.w = v3(0, 0, 0);
    }
  }
}

// ------------------------------------------------------------------------------------------------ reset
struct ResetCfg {
  uint64_t seed;
  int64_t env_index_base;
  int32_t reward_type, goal_shape, max_episode_steps;
  float init_grasp_rate, goal_ground_rate, same_side_rate;
  int32_t stagger;   // XarmConfig.stagger_phases
};
// step counter an env starts from at construction / after an explicit reset: 0, or (global index mod episode length) with
// stagger_phases - time-limit endings then spread evenly over the steps
template <class T>
XD int initial_step_count(const ResetCfg& cfg, int64_t genv) {
  if (!cfg.stagger) return 0;
  const int limit = cfg.max_episode_steps > 0 ? cfg.max_episode_steps : T::MAX_STEPS;
  return (int)(genv % limit);
}

template <class T>
XD void set_joint_init(Env<T>& e, int a, float finger) {
  using MD = typename T::MD;
#pragma unroll
  for (int i = 0; i < MD::N; i++) { e.arm[a].q[i] = i < 7 ? c_joint_init[i] : 0.f; e.arm[a].qd[i] = 0.f; }
  if (MD::F2 >= 0) { e.arm[a].q[MD::F1] = finger; e.arm[a].q[MD::F2 < 0 ? 0 : MD::F2] = finger; }
#pragma unroll
  for (int i = 0; i < MD::N; i++) e.arm[a].qt[i] = e.arm[a].q[i];
}
XD void place_obj(ObjState& b, float x, float y) {
  b.pos = v3(x, y, 0.025f);
  b.quat.x = 0; b.quat.y = 0; b.quat.z = 0; b.quat.w = 1;
  b.v = v3(0, 0, 0); b.w = v3(0, 0, 0);
}

// _sample_goal [REF xarm_reach.py:170-173; xarm_pick_and_place.py:269-287; xarm_stack_tower.py:213-219;
// xarm_push_with_door.py:207-212; xarm_handover.py:370-393]
template <class T>
XD void sample_goal(Env<T>& e, Rng& rng, const ResetCfg& cfg) {
  if (T::TASK == XARM_TASK_REACH) {
    e.goal[0] = rng.box(0.3f, 0.5f); e.goal[1] = rng.box(-0.25f, 0.25f); e.goal[2] = rng.box(0.3f, 0.4f);
  } else if (T::TASK == XARM_TASK_PICK_AND_PLACE) {
    const float lo[3] = {0.35f, -0.25f, 0.025f}, hi[3] = {0.45f, 0.25f, 0.27f};
    if (cfg.goal_shape == XARM_GOAL_AIR) {
      for (int i = 0; i < T::NOBJ; i++) {
        float* g = &e.goal[3 * i];
        for (int c = 0; c < 3; c++) g[c] = rng.box(lo[c], hi[c]);
        if (rng.uniform() < (double)cfg.goal_ground_rate) g[2] = lo[2];
        for (int tries = 0; i > 0 && tries < XARM_GOAL_RESAMPLE_TRIES; tries++) {
          double mn = 1e30;
          for (int j = 0; j < i; j++) {
            double s = 0;
            for (int c = 0; c < 3; c++) { double d = (double)g[c] - (double)e.goal[3 * j + c]; s += d * d; }
            mn = fmin(mn, sqrt(s));
          }
          if (!(mn < 0.05)) break;
          for (int c = 0; c < 3; c++) g[c] = rng.box(lo[c], hi[c]);
        }
      }
    } else {
      float xy[3];
      for (int c = 0; c < 3; c++) xy[c] = rng.box(lo[c], hi[c]);
      for (int i = 0; i < T::NOBJ; i++) { e.goal[3 * i] = xy[0]; e.goal[3 * i + 1] = xy[1]; e.goal[3 * i + 2] = (float)(0.025 * (2 * i + 1)); }
    }
  } else if (T::TASK == XARM_TASK_STACK_TOWER) {
    float x = rng.box(-0.3f, 0.3f), y = rng.box(-0.2f, 0.2f);
    for (int i = 0; i < T::NOBJ; i++) { e.goal[3 * i] = x; e.goal[3 * i + 1] = y; e.goal[3 * i + 2] = (float)(0.025 * (2 * i + 1)); }
  } else if (T::TASK == XARM_TASK_PUSH_WITH_DOOR) {
    e.goal[0] = rng.box(0.1f, 0.3f); e.goal[1] = rng.box(-0.2f, 0.2f); e.goal[2] = (float)0.025;
  } else {
    const float lo[3] = {0.1f, -0.18f, 0.025f}, hi[3] = {0.28f, 0.18f, 0.2f};
    for (int i = 0; i < T::NOBJ; i++) {
      float* g = &e.goal[3 * i];
      for (int c = 0; c < 3; c++) g[c] = rng.box(lo[c], hi[c]);
      for (int tries = 0; i > 0 && tries < XARM_GOAL_RESAMPLE_TRIES; tries++) {
        double mn_obj = 1e30, mn_g = 1e30;
        for (int k = 0; k < T::NOBJ; k++) {
          double dx = (double)g[0] - (double)e.obj[k].pos.x, dy = (double)g[1] - (double)e.obj[k].pos.y;
          mn_obj = fmin(mn_obj, sqrt(dx * dx + dy * dy));
        }
        for (int j = 0; j < i; j++) mn_g = fmin(mn_g, fabs((double)g[1] - (double)e.goal[3 * j + 1]));
        if (!(mn_g < 0.08 || mn_obj < 0.08)) break;
        for (int c = 0; c < 3; c++) g[c] = rng.box(lo[c], hi[c]);
      }
      bool same = rng.uniform() < (double)cfg.same_side_rate;
      if ((e.obj[i].pos.x > 0.f) != same) g[0] = -g[0];
      if (cfg.goal_shape == XARM_GOAL_GROUND) g[2] = (float)0.025;
    }
  }
}

// 5x { IK to the start pose, motor commands, stepSimulation } [REF xarm_pick_and_place.py:252-258; xarm_handover.py:347-353]
template <class T>
XD void servo_reset(Env<T>& e, V3 t0, V3 t1, bool cmd_fingers) {
  using MD = typename T::MD;
  for (int rep = 0; rep < 5; rep++) {
#pragma unroll
    for (int a = 0; a < T::NARM; a++) {
      float qn[7];
      arm_ik<T>(a, e.arm[a].q, a == 0 ? t0 : t1, qn);
#pragma unroll
      for (int i = 0; i < 7; i++) e.arm[a].qt[i] = qn[i];
      if (cmd_fingers && MD::F2 >= 0) { e.arm[a].qt[MD::F1] = 0.02f; e.arm[a].qt[MD::F2 < 0 ? 0 : MD::F2] = 0.02f; }
    }
    simulate<T>(e);
  }
}

// constructor state: what the env looks like before its first reset() (episode 0 draws)
template <class T>
XD void env_construct(Env<T>& e, const ResetCfg& cfg, int64_t genv) {
  using MD = typename T::MD;
  Rng rng = {cfg.seed, (uint64_t)genv, 0u, 0u};
  e.episode = 0; e.step_count = initial_step_count<T>(cfg, genv); e.d_old = 0.f; e.grasp[0] = 0; e.grasp[1] = 0; e.grasp_cmd[0] = 0; e.grasp_cmd[1] = 0; e.door_q = 0.f; e.door_qd = 0.f;
#pragma unroll
  for (int a = 0; a < T::NARM; a++) {
    if (T::TASK == XARM_TASK_PICK_AND_PLACE) {  // its ctor never calls resetJointState [REF xarm_pick_and_place.py:76-79]
#pragma unroll
      for (int i = 0; i < MD::N; i++) { e.arm[a].q[i] = 0.f; e.arm[a].qd[i] = 0.f; e.arm[a].qt[i] = 0.f; }
    } else {
      set_joint_init<T>(e, a, T::TASK == XARM_TASK_HANDOVER ? 0.04f : 0.f);
    }
  }
  for (int i = 0; i < T::NOBJ; i++) {
    float x = 0.f, y = 0.f;
    if (T::TASK == XARM_TASK_PICK_AND_PLACE) { x = rng.box(0.35f, 0.45f); y = rng.box(-0.25f, 0.25f); }
    else if (T::TASK == XARM_TASK_STACK_TOWER) { x = rng.box(-0.3f, 0.3f); y = rng.box(-0.2f, 0.2f); }
    else if (T::TASK == XARM_TASK_PUSH_WITH_DOOR) { x = rng.box(-0.3f, -0.1f); y = rng.box(-0.2f, 0.2f); }
    else { x = rng.box(0.11f, 0.28f); y = rng.box(-0.18f, 0.2f); }
    place_obj(e.obj[i], x, y);
  }
  sample_goal<T>(e, rng, cfg);
}

// The part of _reset_sim between the arm motion and the last stepSimulation: joint teleport (Reach, StackTower,
// PushWithDoor) and object placement [REF xarm_reach.py:163-168; xarm_pick_and_place.py:259-266;
// xarm_stack_tower.py:201-211; xarm_push_with_door.py:195-205; xarm_handover.py:354-367]
template <class T>
XD void env_reset_place(Env<T>& e, Rng& rng, const ResetCfg& cfg) {
  if (T::TASK == XARM_TASK_REACH) {
    set_joint_init<T>(e, 0, 0.f);
  } else if (T::TASK == XARM_TASK_PICK_AND_PLACE) {
    for (int i = 0; i < T::NOBJ; i++) {
      double ug = rng.uniform();
      float x = rng.box(0.35f, 0.45f), y = rng.box(-0.25f, 0.25f);
      if (ug < (double)cfg.init_grasp_rate) place_obj(e.obj[i], 0.4f, 0.0f); else place_obj(e.obj[i], x, y);
    }
  } else if (T::TASK == XARM_TASK_STACK_TOWER || T::TASK == XARM_TASK_PUSH_WITH_DOOR) {
#pragma unroll
    for (int a = 0; a < T::NARM; a++) set_joint_init<T>(e, a, 0.f);
    for (int i = 0; i < T::NOBJ; i++) {
      float x, y;
      if (T::TASK == XARM_TASK_STACK_TOWER) { x = rng.box(-0.3f, 0.3f); y = rng.box(-0.2f, 0.2f); }
      else { x = rng.box(-0.3f, -0.1f); y = rng.box(-0.2f, 0.2f); }
      place_obj(e.obj[i], x, y);
    }
  } else {
    float px[T::NOBJ > 0 ? T::NOBJ : 1], py[T::NOBJ > 0 ? T::NOBJ : 1];
    for (int i = 0; i < T::NOBJ; i++) {
      px[i] = rng.box(0.11f, 0.28f); py[i] = rng.box(-0.18f, 0.2f);
      for (int tries = 0; i > 0 && tries < XARM_GOAL_RESAMPLE_TRIES; tries++) {
        double mn = 1e30;
        for (int j = 0; j < i; j++) mn = fmin(mn, fabs((double)py[i] - (double)py[j]));
        if (!(mn < 0.05)) break;
        px[i] = rng.box(0.11f, 0.28f); py[i] = rng.box(-0.18f, 0.2f);
      }
      if (rng.uniform() < 0.5) px[i] = -px[i];
      place_obj(e.obj[i], px[i], py[i]);
    }
  }
}

// Env.reset() [REF xarm_reach.py:96-102,163-168; xarm_pick_and_place.py:121-127,250-267; xarm_stack_tower.py:115-119,201-211;
// xarm_push_with_door.py:117-121,195-205; xarm_handover.py:141-145,338-368]
template <class T>
XD void env_reset(Env<T>& e, const ResetCfg& cfg, int64_t genv) {
  e.episode += 1; e.step_count = 0;
  Rng rng = {cfg.seed, (uint64_t)genv, e.episode, 0u};
  if (T::TASK == XARM_TASK_PICK_AND_PLACE) servo_reset<T>(e, v3(0.4f, 0.f, 0.12f), v3(0, 0, 0), true);
  if (T::TASK == XARM_TASK_HANDOVER) servo_reset<T>(e, v3(-0.15f, 0.f, 0.15f), v3(0.15f, 0.f, 0.15f), false);
  env_reset_place<T>(e, rng, cfg);
  simulate<T>(e);
  sample_goal<T>(e, rng, cfg);
}

// ------------------------------------------------------------------------------------------------ step
struct StepOut {
  float reward, success;
  bool done, truncated;
};

// everything Env.step does after stepSimulation: _get_obs, _is_success, compute_reward, done, TimeLimit
template <class T>
XD void env_step_outputs(Env<T>& e, const ResetCfg& cfg, Obs<T>& o, StepOut& so);

template <class T>
XD void env_step(Env<T>& e, const float* action, const ResetCfg& cfg, Obs<T>& o, StepOut& so) {
  e.step_count += 1;
  set_action<T>(e, action);
  simulate<T>(e);
  env_step_outputs<T>(e, cfg, o, so);
}

template <class T>
XD void env_step_outputs(Env<T>& e, const ResetCfg& cfg, Obs<T>& o, StepOut& so) {
  get_obs<T>(e, o);
  const float thr = T::THRESHOLD;
  float succ;
  if (T::TASK == XARM_TASK_HANDOVER || (T::TASK == XARM_TASK_PICK_AND_PLACE && T::NOBJ > 1)) {  // [REF xarm_handover.py:395-402], D5
    succ = 1.f;
    for (int i = 0; i < T::NOBJ; i++) if (!(np_dist(o.ag + 3 * i, o.dg + 3 * i, 3) < thr)) succ = 0.f;
  } else {
    succ = np_dist(o.ag, o.dg, T::G) < thr ? 1.f : 0.f;
  }
  float rew;
  const int rt = cfg.reward_type;
  if (T::TASK == XARM_TASK_REACH && rt == XARM_REWARD_DENSE_DIFF) {  // [REF xarm_reach.py:113-116]
    float d = np_dist(o.ag, o.dg, 3);
    rew = __fsub_rn(e.d_old, d); e.d_old = d;
  } else if ((T::TASK == XARM_TASK_PICK_AND_PLACE || T::TASK == XARM_TASK_HANDOVER) && rt == XARM_REWARD_DENSE) {
    rew = reward_dense_staged<T>(e, o);
  } else {
    rew = reward_stateless(T::TASK, rt, T::NOBJ, thr, o.ag, o.dg, T::G);
  }
  const int limit = cfg.max_episode_steps > 0 ? cfg.max_episode_steps : T::MAX_STEPS;
  bool terminated = false;
  if (T::TASK == XARM_TASK_PICK_AND_PLACE) terminated = np_dist(o.ag, o.dg, T::G) < thr;  // [REF xarm_pick_and_place.py:117]
  if (T::TASK == XARM_TASK_HANDOVER) terminated = succ != 0.f;                            // [REF xarm_handover.py:138]
  const bool time_up = e.step_count >= limit;
  so.reward = rew; so.success = succ;
  so.done = terminated || time_up;
  so.truncated = time_up && !terminated;
}
