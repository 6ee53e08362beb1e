// xarm_sim.cuh - per-env device code of the batched gym-xarm step.  One thread owns one env; all per-env
// quantities live in registers / thread-local memory, the SoA state slab in HBM is touched once per kernel.
//
// What it replaces: N9 IK (arm_ik), N3 multibody dynamics (arm_dynamics: world-frame RNEA + CRBA +
// Cholesky, mathematically the ABA Bullet runs), N4-N6 motor/limit/gear rows, N7 box-box contacts (box_box),
// N8 the sequential-impulse PGS (solve), N10/N11 getters/setters (obs assembly, reset).
#pragma once
#include "xarm_tasks.cuh"

#define NOINL __device__ __noinline__

// ------------------------------------------------------------------------------------------------ state
template <class MD>
struct ArmState {
  float q[MD::N], qd[MD::N], qt[MD::N];
};
struct ObjState {
  V3 pos;
  Q4 quat;
  V3 v, w;
};
template <class T>
struct Env {
  ArmState<typename T::MD> arm[T::NARM];
  ObjState obj[T::NOBJ > 0 ? T::NOBJ : 1];
  float door_q, door_qd;
  float goal[T::G];
  int step_count;
  uint32_t episode;
  float d_old;
  int grasp[2];      // both fingers of arm a touch lego 0 in the LAST collision pass (what getContactPoints would report now)
  int grasp_cmd[2];  // the same flags as _set_action read them at the start of the current step [REF xarm_handover.py:262-263]
};
// the two grasp words of the state slab carry (live flag) | (flag snapshot of _set_action) << 1
XD float grasp_word(int live, int cmd) { return (float)((live ? 1 : 0) | (cmd ? 2 : 0)); }

// SoA slab: word w of env e lives at state[w * n + e] (coalesced when consecutive threads own consecutive envs)
template <class T>
XD void env_load(Env<T>& e, const float* __restrict__ s, int64_t n, int64_t i) {
  using MD = typename T::MD;
  int w = 0;
#pragma unroll
  for (int a = 0; a < T::NARM; a++) {
#pragma unroll
    for (int k = 0; k < MD::N; k++) e.arm[a].q[k] = s[(w++) * n + i];
#pragma unroll
    for (int k = 0; k < MD::N; k++) e.arm[a].qd[k] = s[(w++) * n + i];
#pragma unroll
    for (int k = 0; k < MD::N; k++) e.arm[a].qt[k] = s[(w++) * n + i];
  }
#pragma unroll
  for (int o = 0; o < T::NOBJ; o++) {
    ObjState& b = e.obj[o];
    b.pos.x = s[(w++) * n + i]; b.pos.y = s[(w++) * n + i]; b.pos.z = s[(w++) * n + i];
    b.quat.x = s[(w++) * n + i]; b.quat.y = s[(w++) * n + i]; b.quat.z = s[(w++) * n + i]; b.quat.w = s[(w++) * n + i];
    b.v.x = s[(w++) * n + i]; b.v.y = s[(w++) * n + i]; b.v.z = s[(w++) * n + i];
    b.w.x = s[(w++) * n + i]; b.w.y = s[(w++) * n + i]; b.w.z = s[(w++) * n + i];
  }
  if (T::HAS_DOOR) { e.door_q = s[(w++) * n + i]; e.door_qd = s[(w++) * n + i]; } else { e.door_q = 0; e.door_qd = 0; }
#pragma unroll
  for (int k = 0; k < T::G; k++) e.goal[k] = s[(w++) * n + i];
  e.step_count = (int)s[(w++) * n + i];
  e.episode = (uint32_t)s[(w++) * n + i];
  e.d_old = s[(w++) * n + i];
#pragma unroll
  for (int a = 0; a < 2; a++) {
    const int g = (int)s[(w++) * n + i];
    e.grasp[a] = g & 1; e.grasp_cmd[a] = (g >> 1) & 1;
  }
}
template <class T>
XD void env_store(const Env<T>& e, float* __restrict__ s, int64_t n, int64_t i) {
  using MD = typename T::MD;
  int w = 0;
#pragma unroll
  for (int a = 0; a < T::NARM; a++) {
#pragma unroll
    for (int k = 0; k < MD::N; k++) s[(w++) * n + i] = e.arm[a].q[k];
#pragma unroll
    for (int k = 0; k < MD::N; k++) s[(w++) * n + i] = e.arm[a].qd[k];
#pragma unroll
    for (int k = 0; k < MD::N; k++) s[(w++) * n + i] = e.arm[a].qt[k];
  }
#pragma unroll
  for (int o = 0; o < T::NOBJ; o++) {
    const ObjState& b = e.obj[o];
    s[(w++) * n + i] = b.pos.x; s[(w++) * n + i] = b.pos.y; s[(w++) * n + i] = b.pos.z;
    s[(w++) * n + i] = b.quat.x; s[(w++) * n + i] = b.quat.y; s[(w++) * n + i] = b.quat.z; s[(w++) * n + i] = b.quat.w;
    s[(w++) * n + i] = b.v.x; s[(w++) * n + i] = b.v.y; s[(w++) * n + i] = b.v.z;
    s[(w++) * n + i] = b.w.x; s[(w++) * n + i] = b.w.y; s[(w++) * n + i] = b.w.z;
  }
  if (T::HAS_DOOR) { s[(w++) * n + i] = e.door_q; s[(w++) * n + i] = e.door_qd; }
#pragma unroll
  for (int k = 0; k < T::G; k++) s[(w++) * n + i] = e.goal[k];
  s[(w++) * n + i] = (float)e.step_count;
  s[(w++) * n + i] = (float)e.episode;
  s[(w++) * n + i] = e.d_old;
  s[(w++) * n + i] = grasp_word(e.grasp[0], e.grasp_cmd[0]);
  s[(w++) * n + i] = grasp_word(e.grasp[1], e.grasp_cmd[1]);
}

// the dynamic part only (what a substep changes): joint positions / velocities, object poses / velocities, the door.
// Motor targets, goal, counters and grasp flags are not touched (the arms' qt words are skipped).
template <class T>
XD void env_load_dyn(Env<T>& e, const float* __restrict__ s, int64_t n, int64_t i) {
  using MD = typename T::MD;
  int w = 0;
#pragma unroll
  for (int a = 0; a < T::NARM; a++) {
#pragma unroll
    for (int k = 0; k < MD::N; k++) e.arm[a].q[k] = s[(w++) * n + i];
#pragma unroll
    for (int k = 0; k < MD::N; k++) e.arm[a].qd[k] = s[(w++) * n + i];
    w += MD::N;
  }
#pragma unroll
  for (int o = 0; o < T::NOBJ; o++) {
    ObjState& b = e.obj[o];
    b.pos.x = s[(w++) * n + i]; b.pos.y = s[(w++) * n + i]; b.pos.z = s[(w++) * n + i];
    b.quat.x = s[(w++) * n + i]; b.quat.y = s[(w++) * n + i]; b.quat.z = s[(w++) * n + i]; b.quat.w = s[(w++) * n + i];
    b.v.x = s[(w++) * n + i]; b.v.y = s[(w++) * n + i]; b.v.z = s[(w++) * n + i];
    b.w.x = s[(w++) * n + i]; b.w.y = s[(w++) * n + i]; b.w.z = s[(w++) * n + i];
  }
  if (T::HAS_DOOR) { e.door_q = s[(w++) * n + i]; e.door_qd = s[(w++) * n + i]; } else { e.door_q = 0; e.door_qd = 0; }
}
template <class T>
XD void env_store_dyn(const Env<T>& e, float* __restrict__ s, int64_t n, int64_t i) {
  using MD = typename T::MD;
  int w = 0;
#pragma unroll
  for (int a = 0; a < T::NARM; a++) {
#pragma unroll
    for (int k = 0; k < MD::N; k++) s[(w++) * n + i] = e.arm[a].q[k];
#pragma unroll
    for (int k = 0; k < MD::N; k++) s[(w++) * n + i] = e.arm[a].qd[k];
    w += MD::N;
  }
#pragma unroll
  for (int o = 0; o < T::NOBJ; o++) {
    const ObjState& b = e.obj[o];
    s[(w++) * n + i] = b.pos.x; s[(w++) * n + i] = b.pos.y; s[(w++) * n + i] = b.pos.z;
    s[(w++) * n + i] = b.quat.x; s[(w++) * n + i] = b.quat.y; s[(w++) * n + i] = b.quat.z; s[(w++) * n + i] = b.quat.w;
    s[(w++) * n + i] = b.v.x; s[(w++) * n + i] = b.v.y; s[(w++) * n + i] = b.v.z;
    s[(w++) * n + i] = b.w.x; s[(w++) * n + i] = b.w.y; s[(w++) * n + i] = b.w.z;
  }
  if (T::HAS_DOOR) { s[(w++) * n + i] = e.door_q; s[(w++) * n + i] = e.door_qd; }
}

// ------------------------------------------------------------------------------------------------ RNG (Appendix E)
XD void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(XARM_PHILOX_M0, c[0]), lo0 = XARM_PHILOX_M0 * c[0];
    uint32_t hi1 = __umulhi(XARM_PHILOX_M1, c[2]), lo1 = XARM_PHILOX_M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += XARM_PHILOX_W0; k1 += XARM_PHILOX_W1;
  }
}
struct Rng {
  uint64_t seed, genv;
  uint32_t episode, draw;
  XD double uniform() {
    uint32_t k = draw++;
    uint32_t c[4] = {(uint32_t)genv, (uint32_t)(genv >> 32), episode, k >> 2};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint32_t x = (k & 3) == 0 ? c[0] : ((k & 3) == 1 ? c[1] : ((k & 3) == 2 ? c[2] : c[3]));
    return (double)x * (1.0 / 4294967296.0);
  }
  // gym Box.sample(): float64 arithmetic on float32-rounded bounds, then cast to float32
  XD float box(float lo, float hi) {
    double u = uniform();
    return (float)((double)lo + ((double)hi - (double)lo) * u);
  }
};

// ------------------------------------------------------------------------------------------------ kinematics / dynamics
template <class MD>
struct ArmDyn {
  SV S[MD::N];                          // joint motion subspaces (world, about the origin)
  float Minv[MD::N * (MD::N + 1) / 2];  // packed symmetric inverse joint-space inertia
  float qdu[MD::N];                     // velocities after the unconstrained update, then the solved velocities
  M3 Rh;                                // link7 frame (eef / hand; the Panda fingers share its orientation)
  V3 ph, pf1, pf2;                      // origins of the link7 and finger frames
};

template <class T>
XD void arm_base(int a, M3& R, V3& p) {
  R = m3_identity();
  if (T::NARM == 2 && a == 1) { R.m[0] = -1.f; R.m[4] = -1.f; }  // yaw pi [REF xarm_handover.py:53]
  p = v3(T::base_x(a), 0.f, 0.f);
}

// FK of the first 7 joints only: eef frame, joint origins and axes (for IK and _set_action)
template <class T>
XD void arm_fk7(int a, const float* q, M3& Re, V3& pe, V3* org, V3* axs) {
  using MD = typename T::MD;
  M3 R; V3 p;
  arm_base<T>(a, R, p);
#pragma unroll 1
  for (int i = 0; i < 7; i++) {
    const float* r0 = MD::R0(i);
    M3 R0;
#pragma unroll
    for (int k = 0; k < 9; k++) R0.m[k] = r0[k];
    p = p + R * MD::t0(i);
    R = R * R0;
    axs[i] = R * MD::axis(i);
    org[i] = p;
    R = R * m3_axis_angle(MD::axis(i), q[i]);
  }
  Re = R; pe = p;
}

// calculateInverseKinematics (N9): n_ik damped-least-squares iterations on the 7 arm joints
template <class T>
NOINL void arm_ik(int a, const float* q_in, V3 target, float* q_out) {
  // (1) one FK per iteration - the FK that checks the residual at the end
  // of an iteration IS the FK the next iteration starts from - and (2) the 7 x 7 normal equations and their Cholesky solve
  // unrolled, so that A and b live in registers.
  float q[7];
#pragma unroll
  for (int i = 0; i < 7; i++) q[i] = q_in[i];
  M3 Re; V3 pe, org[7], axs[7];
  arm_fk7<T>(a, q, Re, pe, org, axs);
  for (int it = 0; it < T::NIK; it++) {
    float e[6];
    e[0] = target.x - pe.x; e[1] = target.y - pe.y; e[2] = target.z - pe.z;
    Q4 qc = m3_to_quat(Re);
    Q4 qT = {1.f, 0.f, 0.f, 0.f}, qci = {-qc.x, -qc.y, -qc.z, qc.w};
    Q4 dq = quat_mul(qT, qci);
    // rotation vector of dq: Bullet takes angle = 2 acos(w) and axis = xyz / sqrt(1 - w^2); for a unit quaternion
    // that is 2 atan2(|xyz|, w) * xyz / |xyz|, which stays accurate in float32 when the error is small
    float vn = sqrtf(dq.x * dq.x + dq.y * dq.y + dq.z * dq.z);
    float angle = 2.f * atan2f(vn, dq.w);
    if (angle > 3.14159265358979f) angle -= 6.28318530717959f;
    float an = vn > 1e-30f ? angle / vn : 0.f;
    e[3] = an * dq.x; e[4] = an * dq.y; e[5] = an * dq.z;
    float J[6][7];
#pragma unroll
    for (int j = 0; j < 7; j++) {
      V3 lin = cross(axs[j], pe - org[j]);
      J[0][j] = lin.x; J[1][j] = lin.y; J[2][j] = lin.z; J[3][j] = axs[j].x; J[4][j] = axs[j].y; J[5][j] = axs[j].z;
    }
    float A[28], b[7];
#pragma unroll
    for (int i = 0; i < 7; i++) {
#pragma unroll
      for (int j = 0; j <= i; j++) {
        float s = (i == j) ? (float)XARM_IK_DAMPING : 0.f;
#pragma unroll
        for (int r = 0; r < 6; r++) s += J[r][i] * J[r][j];
        A[tri(i, j)] = s;
      }
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < 6; r++) s += J[r][i] * e[r];
      b[i] = s;
    }
#pragma unroll
    for (int j = 0; j < 7; j++) {
      float s = A[tri(j, j)];
#pragma unroll
      for (int k = 0; k < j; k++) s -= A[tri(j, k)] * A[tri(j, k)];
      float d = sqrtf(s), di = 1.f / d;
      A[tri(j, j)] = di;
#pragma unroll
      for (int i = j + 1; i < 7; i++) {
        float t = A[tri(i, j)];
#pragma unroll
        for (int k = 0; k < j; k++) t -= A[tri(i, k)] * A[tri(j, k)];
        A[tri(i, j)] = t * di;
      }
    }
#pragma unroll
    for (int i = 0; i < 7; i++) {
      float s = b[i];
#pragma unroll
      for (int k = 0; k < i; k++) s -= A[tri(i, k)] * b[k];
      b[i] = s * A[tri(i, i)];
    }
#pragma unroll
    for (int i = 6; i >= 0; i--) {
      float s = b[i];
#pragma unroll
      for (int k = i + 1; k < 7; k++) s -= A[tri(k, i)] * b[k];
      b[i] = s * A[tri(i, i)];
    }
    float mx = 0.f;
#pragma unroll
    for (int i = 0; i < 7; i++) mx = fmaxf(mx, fabsf(b[i]));
    float sc = mx > (float)XARM_IK_MAX_STEP ? (float)XARM_IK_MAX_STEP / mx : 1.f;
#pragma unroll
    for (int i = 0; i < 7; i++) q[i] += sc * b[i];
    arm_fk7<T>(a, q, Re, pe, org, axs);
    if (norm(target - pe) < (float)XARM_IK_RESIDUAL) break;
  }
#pragma unroll
  for (int i = 0; i < 7; i++) q_out[i] = q[i];
}

// Full-arm pass: FK, spatial velocities, bias forces (world-frame RNEA), joint-space inertia (CRBA), its inverse
// (Cholesky) and the unconstrained velocity update qdu = qd + h * Minv (tau - bias).  Equivalent to the ABA +
// calcAccelerationDeltas pair Bullet runs per substep (N3).
// The link passes are ROLLED loops over thread-local arrays: a fully unrolled version (~10k straight-line instructions) is
// instruction-fetch bound on the SM (profiles/: every 128-byte line costs an L2 round trip).  The Cholesky factorisation,
// its inverse and Minv (static triangular indices, ~1.1 k FMAs) ARE unrolled: their arrays then live in registers instead of
// thread-local memory - the setup kernel went from 63 to 48 us per launch on average.
// FLAT = true (latency form, for the short lists of the auto-reset tail: one warp per SM, 255 registers): the link passes are
// fully unrolled as well, so f[], I[], S[] and M[] live in registers and no thread-local load sits in the dependency chains.
// Same arithmetic in the same order as the rolled form.
template <class T, bool FLAT>
XD void arm_dynamics_impl(int a, const ArmState<typename T::MD>& st, bool apply_damping, ArmDyn<typename T::MD>& D);
template <class T>
NOINL void arm_dynamics_call(int a, const ArmState<typename T::MD>& st, bool apply_damping, ArmDyn<typename T::MD>& D) {
  arm_dynamics_impl<T, false>(a, st, apply_damping, D);
}
// throughput form: one out-of-line copy per kernel; latency form: inlined, so that D and st stay in registers too
template <class T, bool FLAT = false>
XD void arm_dynamics(int a, const ArmState<typename T::MD>& st, bool apply_damping, ArmDyn<typename T::MD>& D) {
  if constexpr (FLAT) arm_dynamics_impl<T, true>(a, st, apply_damping, D);
  else arm_dynamics_call<T>(a, st, apply_damping, D);
}
template <class T, bool FLAT>
XD void arm_dynamics_impl(int a, const ArmState<typename T::MD>& st, bool apply_damping, ArmDyn<typename T::MD>& D) {
  using MD = typename T::MD;
  constexpr int N = MD::N, NT = N * (N + 1) / 2;
  constexpr int UNR = FLAT ? N : 1;
  const float h = (float)T::H;
  SV f[N];
  SI I[N];
  M3 Rb; V3 pb;
  arm_base<T>(a, Rb, pb);
  if constexpr (MD::N == 9 && MD::HAS_BOXES) {
    // xArm7 + Panda gripper: a chain (joints 0..6) with the two fingers on link 6, and the URDF parts sorted by owner.
    // ONE pass over the links does FK, velocity, bias acceleration, link force and the per-part gravity / damping: the
    // frame, velocity and acceleration of the parent are carried in registers (and a copy of link 6's for the fingers),
    // so only f[], I[] and D.S[] remain as thread-local arrays (the three separate passes of the generic form below
    // keep R[], p[], v[] as well: 162 more words of thread-local traffic per env and substep).  Same arithmetic, same order.
    M3 Rp = Rb, R6 = Rb; V3 pp = pb, p6 = pb;
    SV vp = sv_zero(), v6 = sv_zero(), ap = sv_zero(), a6 = sv_zero();
    int pt = 0;
#pragma unroll UNR
    for (int i = 0; i < N; i++) {
      const bool root = i == 0, fin = i > 6;
      const M3 Rpar = fin ? R6 : Rp;
      const V3 ppar = fin ? p6 : pp;
      const float* r0 = MD::R0(i);
      M3 R0;
#pragma unroll
      for (int k = 0; k < 9; k++) R0.m[k] = r0[k];
      const M3 Rj = Rpar * R0;
      const V3 tj = ppar + Rpar * MD::t0(i);
      const V3 ax = Rj * MD::axis(i);
      SV Si; M3 Ri; V3 pi_;
      if (!MD::prismatic(i)) {
        Ri = Rj * m3_axis_angle(MD::axis(i), st.q[i]);
        pi_ = tj;
        Si.a = ax; Si.l = cross(tj, ax);
      } else {
        Ri = Rj;
        pi_ = tj + st.q[i] * ax;
        Si.a = v3(0, 0, 0); Si.l = ax;
      }
      D.S[i] = Si;
      const SV vj = st.qd[i] * Si;
      const SV vi = (root ? sv_zero() : (fin ? v6 : vp)) + vj;
      SV ai = motion_cross(vi, vj);
      if (!root) ai += fin ? a6 : ap;
      const SI Ii = si_make(MD::mass(i), pi_ + Ri * MD::com(i), rotate_sym(Ri, MD::inertia(i)));
      I[i] = Ii;
      SV fi = Ii * ai + force_cross(vi, Ii * vi);
      const V3 w = vi.a;
      const V3 Gww = rotate_sym(Ri, MD::central(i)) * w;
      if (!XARM_MB_USE_GYRO) fi.a -= cross(w, Gww);
      fi.a += ((float)XARM_MB_ANGULAR_DAMPING * (1.f + norm(w))) * Gww;
      for (; pt < MD::NPART && MD::part_owner(pt) == i; pt++) {  // gravity + Bullet's linear velocity damping of every URDF link
        const V3 c = pi_ + Ri * MD::part_com(pt);
        const V3 vc = vi.l + cross(w, c);
        const float m = MD::part_mass(pt);
        V3 F = (-m * (float)XARM_MB_LINEAR_DAMPING * (1.f + norm(vc))) * vc;
        F.z -= m * (float)XARM_GRAVITY;
        fi.a -= cross(c, F);
        fi.l -= F;
      }
      f[i] = fi;
      if (i == MD::EEF) { D.Rh = Ri; D.ph = pi_; }
      if (i == MD::F1) D.pf1 = pi_;
      if (i == (MD::F2 < 0 ? 0 : MD::F2)) D.pf2 = pi_;
      if (!fin) { Rp = Ri; pp = pi_; vp = vi; ap = ai; }
      if (i == 6) { R6 = Ri; p6 = pi_; v6 = vi; a6 = ai; }
    }
  } else {
  M3 R[N]; V3 p[N];
  SV v[N];
#pragma unroll 1
  for (int i = 0; i < N; i++) {
    const int pi = MD::parent(i);
    const M3 Rp = pi < 0 ? Rb : R[pi < 0 ? 0 : pi];
    const V3 pp = pi < 0 ? pb : p[pi < 0 ? 0 : pi];
    const float* r0 = MD::R0(i);
    M3 R0;
#pragma unroll
    for (int k = 0; k < 9; k++) R0.m[k] = r0[k];
    const M3 Rj = Rp * R0;
    const V3 tj = pp + Rp * MD::t0(i);
    const V3 ax = Rj * MD::axis(i);
    SV Si;
    if (!MD::prismatic(i)) {
      R[i] = Rj * m3_axis_angle(MD::axis(i), st.q[i]);
      p[i] = tj;
      Si.a = ax; Si.l = cross(tj, ax);
    } else {
      R[i] = Rj;
      p[i] = tj + st.q[i] * ax;
      Si.a = v3(0, 0, 0); Si.l = ax;
    }
    D.S[i] = Si;
    const SV vj = st.qd[i] * Si;
    const SV vi = (pi < 0 ? sv_zero() : v[pi < 0 ? 0 : pi]) + vj;
    v[i] = vi;
    // bias acceleration a_i = a_parent + v_i x vj (f[] holds it until the force pass below)
    SV ai = motion_cross(vi, vj);
    if (pi >= 0) ai += f[pi < 0 ? 0 : pi];
    // link force: f_i = I a + v x* I v - gyro + angular damping; gravity and linear damping per URDF part below
    const M3 Ri = R[i];
    const SI Ii = si_make(MD::mass(i), p[i] + Ri * MD::com(i), rotate_sym(Ri, MD::inertia(i)));
    I[i] = Ii;
    f[i] = ai;
  }
  D.Rh = R[MD::EEF]; D.ph = p[MD::EEF];
  if (MD::HAS_BOXES) { D.pf1 = p[MD::F1]; D.pf2 = p[MD::F2 < 0 ? 0 : MD::F2]; }
#pragma unroll 1
  for (int i = 0; i < N; i++) {
    const SV vi = v[i];
    const SI Ii = I[i];
    SV fi = Ii * f[i] + force_cross(vi, Ii * vi);
    const V3 w = vi.a;
    const V3 Gww = rotate_sym(R[i], MD::central(i)) * w;
    if (!XARM_MB_USE_GYRO) fi.a -= cross(w, Gww);
    fi.a += ((float)XARM_MB_ANGULAR_DAMPING * (1.f + norm(w))) * Gww;  // minus the external damping torque -(G w) ka
    f[i] = fi;
  }
#pragma unroll 1
  for (int pt = 0; pt < MD::NPART; pt++) {  // gravity + Bullet's linear velocity damping of every URDF link
    const int i = MD::part_owner(pt);
    const V3 c = p[i] + R[i] * MD::part_com(pt);
    const V3 w = v[i].a;
    const V3 vc = v[i].l + cross(w, c);
    const float m = MD::part_mass(pt);
    V3 F = (-m * (float)XARM_MB_LINEAR_DAMPING * (1.f + norm(vc))) * vc;
    F.z -= m * (float)XARM_GRAVITY;
    SV fi = f[i];
    fi.a -= cross(c, F);
    fi.l -= F;
    f[i] = fi;
  }
  }
  // backward: bias torques, composite inertias, joint-space inertia matrix
  float M[NT], rhs[N];
#pragma unroll UNR
  for (int i = N - 1; i >= 0; i--) {
    const int pi = MD::parent(i);
    const SV Si = D.S[i], fi = f[i];
    const float tau = apply_damping ? -MD::damping(i) * st.qd[i] : 0.f;
    rhs[i] = tau - dot(Si, fi);
    const SI Ii = I[i];
    const SV F = Ii * Si;
    unsigned anc = 0u;
#pragma unroll UNR
    for (int k = i; k >= 0; k = MD::parent(k)) anc |= 1u << k;
#pragma unroll UNR
    for (int j = 0; j <= i; j++) M[tri(i, j)] = (anc >> j & 1u) ? dot(D.S[j], F) : 0.f;
    if (pi >= 0) { f[pi] += fi; I[pi] = I[pi] + Ii; }
  }
  // Cholesky / inverse / Minv fully unrolled: M, Li and Minv live in registers (static indices), ~1.1 k straight-line FMAs
  float Mr[NT];
#pragma unroll
  for (int k = 0; k < NT; k++) Mr[k] = M[k];
#pragma unroll
  for (int j = 0; j < N; j++) {
    float s = Mr[tri(j, j)];
#pragma unroll
    for (int k = 0; k < j; k++) s -= Mr[tri(j, k)] * Mr[tri(j, k)];
    float di = rsqrtf(s);
    di = di * (1.5f - 0.5f * s * di * di);
    Mr[tri(j, j)] = di;
#pragma unroll
    for (int i = j + 1; i < N; i++) {
      float t = Mr[tri(i, j)];
#pragma unroll
      for (int k = 0; k < j; k++) t -= Mr[tri(i, k)] * Mr[tri(j, k)];
      Mr[tri(i, j)] = t * di;
    }
  }
  float Li[NT];
#pragma unroll
  for (int j = 0; j < N; j++) {
    Li[tri(j, j)] = Mr[tri(j, j)];
#pragma unroll
    for (int i = j + 1; i < N; i++) {
      float s = 0.f;
#pragma unroll
      for (int k = j; k < i; k++) s += Mr[tri(i, k)] * Li[tri(k, j)];
      Li[tri(i, j)] = -s * Mr[tri(i, i)];
    }
  }
  float Mv[NT];
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = 0; j <= i; j++) {
      float s = 0.f;
#pragma unroll
      for (int k = i; k < N; k++) s += Li[tri(k, i)] * Li[tri(k, j)];
      Mv[tri(i, j)] = s;
    }
#pragma unroll
  for (int k = 0; k < NT; k++) D.Minv[k] = Mv[k];
  float rr[N];
#pragma unroll
  for (int i = 0; i < N; i++) rr[i] = rhs[i];
#pragma unroll
  for (int i = 0; i < N; i++) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < N; j++) s += Mv[tri(i, j)] * rr[j];
    D.qdu[i] = st.qd[i] + h * s;
  }
}

// hand (PyBullet link 9) COM position and linear velocity: getLinkState(arm, 9, computeLinkVelocity=1) (N10)
template <class T>
NOINL void hand_state(int a, const ArmState<typename T::MD>& st, V3& pos, V3& vel) {
  using MD = typename T::MD;
  M3 Re; V3 pe, org[7], axs[7];
  arm_fk7<T>(a, st.q, Re, pe, org, axs);
  pos = pe + Re * MD::hand_com();
  V3 vv = v3(0, 0, 0);
#pragma unroll
  for (int j = 0; j < 7; j++) vv += st.qd[j] * cross(axs[j], pos - org[j]);
  vel = vv;
}

// ------------------------------------------------------------------------------------------------ collision (N7)
struct Box {
  V3 c;
  M3 R;  // columns = box axes
  V3 h;
};
struct CPoint {
  V3 pa, pb, n;
  float depth;
};

XD int clip_poly(const float (*in)[2], int n, int axis, float sign, float lim, float (*out)[2]) {
  int m = 0;
  for (int i = 0; i < n; i++) {
    const float* a = in[i];
    const float* b = in[(i + 1 == n) ? 0 : i + 1];
    float da = sign * a[axis] - lim, db = sign * b[axis] - lim;
    if (da <= 0.f) { out[m][0] = a[0]; out[m][1] = a[1]; m++; }
    if ((da < 0.f && db > 0.f) || (da > 0.f && db < 0.f)) {
      float s = da / (da - db);
      out[m][0] = a[0] + s * (b[0] - a[0]); out[m][1] = a[1] + s * (b[1] - a[1]); m++;
    }
    if (m >= 8) break;
  }
  return m;
}

// 15-axis SAT + reference-face clipping / edge-edge closest points; normal points from B to A.
NOINL int box_box(const Box& A, const Box& B, CPoint* out, int max_out) {
  V3 Aa[3], Ba[3];
#pragma unroll
  for (int i = 0; i < 3; i++) { Aa[i] = col(A.R, i); Ba[i] = col(B.R, i); }
  V3 Tv = B.c - A.c;
  float Rm[3][3], Q[3][3], Ta[3];
  float ah[3] = {A.h.x, A.h.y, A.h.z}, bh[3] = {B.h.x, B.h.y, B.h.z};
#pragma unroll
  for (int i = 0; i < 3; i++) {
    Ta[i] = dot(Tv, Aa[i]);
#pragma unroll
    for (int j = 0; j < 3; j++) { Rm[i][j] = dot(Aa[i], Ba[j]); Q[i][j] = fabsf(Rm[i][j]); }
  }
  const float margin = (float)XARM_CONTACT_MARGIN;
  float best = -1e30f, nsign = 1.f;
  int code = -1;
  V3 naxis = v3(0, 0, 0);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float s = fabsf(Ta[i]) - (ah[i] + bh[0] * Q[i][0] + bh[1] * Q[i][1] + bh[2] * Q[i][2]);
    if (s > margin) return 0;
    if (s > best) { best = s; code = i; nsign = Ta[i] < 0.f ? -1.f : 1.f; naxis = Aa[i]; }
  }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    float tb = dot(Tv, Ba[j]);
    float s = fabsf(tb) - (bh[j] + ah[0] * Q[0][j] + ah[1] * Q[1][j] + ah[2] * Q[2][j]);
    if (s > margin) return 0;
    if (s > best) { best = s; code = 3 + j; nsign = tb < 0.f ? -1.f : 1.f; naxis = Ba[j]; }
  }
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      V3 ax = cross(Aa[i], Ba[j]);
      float l = norm(ax);
      if (l < 1e-6f) continue;
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      float tp = dot(Tv, ax);
      float ra = ah[i1] * Q[i2][j] + ah[i2] * Q[i1][j];
      float rb = bh[j1] * Q[i][j2] + bh[j2] * Q[i][j1];
      float s = (fabsf(tp) - (ra + rb)) / l;
      if (s > margin) return 0;
      if (s * 1.05f > best) { best = s; code = 6 + 3 * i + j; nsign = tp < 0.f ? -1.f : 1.f; naxis = (1.f / l) * ax; }
    }
  if (code < 0 || max_out < 1) return 0;
  V3 nAB = nsign * naxis;
  float depth = -best;
  if (code >= 6) {
    int i = (code - 6) / 3, j = (code - 6) % 3;
    V3 pa = A.c, pb = B.c;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      if (k != i) pa += ((dot(nAB, Aa[k]) > 0.f ? 1.f : -1.f) * ah[k]) * Aa[k];
      if (k != j) pb += ((dot(nAB, Ba[k]) > 0.f ? -1.f : 1.f) * bh[k]) * Ba[k];
    }
    V3 ua = i == 0 ? Aa[0] : (i == 1 ? Aa[1] : Aa[2]);
    V3 ub = j == 0 ? Ba[0] : (j == 1 ? Ba[1] : Ba[2]);
    V3 d = pb - pa;
    float uaub = dot(ua, ub), q1 = dot(ua, d), q2 = -dot(ub, d), den = 1.f - uaub * uaub;
    float sa = 0.f, sb = 0.f;
    if (den > 1e-4f) { sa = (q1 + uaub * q2) / den; sb = (uaub * q1 + q2) / den; }
    out[0].pa = pa + sa * ua; out[0].pb = pb + sb * ub; out[0].n = -nAB; out[0].depth = depth;
    return 1;
  }
  const bool refA = code < 3;
  const Box& Rf = refA ? A : B;
  const Box& In = refA ? B : A;
  const int ra = refA ? code : code - 3;
  V3 nref = refA ? nAB : -nAB;
  V3 Ra[3], Ia[3];
#pragma unroll
  for (int i = 0; i < 3; i++) { Ra[i] = col(Rf.R, i); Ia[i] = col(In.R, i); }
  float rh[3] = {Rf.h.x, Rf.h.y, Rf.h.z}, ih[3] = {In.h.x, In.h.y, In.h.z};
  int ia = 0; float bd = -1.f;
#pragma unroll
  for (int i = 0; i < 3; i++) { float d = fabsf(dot(nref, Ia[i])); if (d > bd) { bd = d; ia = i; } }
  V3 Iax = ia == 0 ? Ia[0] : (ia == 1 ? Ia[1] : Ia[2]);
  float isg = dot(nref, Iax) > 0.f ? -1.f : 1.f;
  const int i1 = (ia + 1) % 3, i2 = (ia + 2) % 3, r1 = (ra + 1) % 3, r2 = (ra + 2) % 3;
  V3 I1 = i1 == 0 ? Ia[0] : (i1 == 1 ? Ia[1] : Ia[2]), I2 = i2 == 0 ? Ia[0] : (i2 == 1 ? Ia[1] : Ia[2]);
  V3 R1 = r1 == 0 ? Ra[0] : (r1 == 1 ? Ra[1] : Ra[2]), R2 = r2 == 0 ? Ra[0] : (r2 == 1 ? Ra[1] : Ra[2]);
  float h1 = ih[i1], h2 = ih[i2], hr1 = rh[r1], hr2 = rh[r2], hra = rh[ra];
  V3 fc = In.c + (isg * ih[ia]) * Iax;
  float poly[8][2], tmp[8][2];
  const float sg[4][2] = {{1, 1}, {-1, 1}, {-1, -1}, {1, -1}};
#pragma unroll
  for (int c = 0; c < 4; c++) {
    V3 vtx = fc + (sg[c][0] * h1) * I1 + (sg[c][1] * h2) * I2;
    V3 d = vtx - Rf.c;
    poly[c][0] = dot(d, R1); poly[c][1] = dot(d, R2);
  }
  int n = 4;
  n = clip_poly(poly, n, 0, 1.f, hr1, tmp); if (!n) return 0;
  n = clip_poly(tmp, n, 0, -1.f, hr1, poly); if (!n) return 0;
  n = clip_poly(poly, n, 1, 1.f, hr2, tmp); if (!n) return 0;
  n = clip_poly(tmp, n, 1, -1.f, hr2, poly); if (!n) return 0;
  V3 inorm = isg * Iax;
  float denom = dot(inorm, nref);
  CPoint cand[8]; int nc = 0;
  for (int c = 0; c < n; c++) {
    V3 base = Rf.c + poly[c][0] * R1 + poly[c][1] * R2;
    float z = fabsf(denom) > 1e-9f ? dot(fc - base, inorm) / denom : 0.f;
    V3 pin = base + z * nref;
    float dep = hra - z;
    if (dep < -margin) continue;
    CPoint k;
    k.depth = dep;
    if (refA) { k.pb = pin; k.pa = pin + dep * nref; k.n = -nref; }
    else { k.pa = pin; k.pb = pin + dep * nref; k.n = nref; }
    cand[nc++] = k;
  }
  if (nc > 4) {
    int keep[4]; int i0 = 0;
    for (int c = 1; c < nc; c++) if (cand[c].depth > cand[i0].depth) i0 = c;
    keep[0] = i0;
    int ib = -1; float bdist = -1.f;
    for (int c = 0; c < nc; c++) { V3 d = cand[c].pb - cand[i0].pb; float l = dot(d, d); if (c != i0 && l > bdist) { bdist = l; ib = c; } }
    keep[1] = ib;
    V3 e1 = cand[ib].pb - cand[i0].pb;
    int imax = -1, imin = -1; float amax = 0.f, amin = 0.f;
    for (int c = 0; c < nc; c++) {
      if (c == i0 || c == ib) continue;
      float ar = dot(cross(e1, cand[c].pb - cand[i0].pb), nref);
      if (imax < 0 || ar > amax) { amax = ar; imax = c; }
      if (imin < 0 || ar < amin) { amin = ar; imin = c; }
    }
    keep[2] = imax; keep[3] = imin;
    CPoint red[4]; int nr = 0;
    for (int c = 0; c < 4; c++) {
      bool dup = keep[c] < 0;
      for (int d = 0; d < c && !dup; d++) if (keep[d] == keep[c]) dup = true;
      if (!dup) red[nr++] = cand[keep[c]];
    }
    for (int c = 0; c < nr; c++) cand[c] = red[c];
    nc = nr;
  }
  if (nc > max_out) nc = max_out;
  for (int c = 0; c < nc; c++) out[c] = cand[c];
  return nc;
}

// ------------------------------------------------------------------------------------------------ substep
XD void plane_space(V3 n, V3& p, V3& q) {  // btPlaneSpace1
  if (fabsf(n.z) > 0.7071067811865475244f) {
    float a = n.y * n.y + n.z * n.z, k = rsqrtf(a);
    p = v3(0.f, -n.z * k, n.y * k);
    q = v3(a * k, -n.x * p.z, n.x * p.y);
  } else {
    float a = n.x * n.x + n.y * n.y, k = rsqrtf(a);
    p = v3(-n.y * k, n.x * k, 0.f);
    q = v3(-n.z * p.y, n.z * p.x, a * k);
  }
}

XD V3 door_axis() { const float a[3] = XARM_DOOR_AXIS; return v3(a[0], a[1], a[2]); }

// Contact rows live in thread-local memory (dynamic count and indexing; L1-resident while the PGS sweeps run).
// Everything the arm rows need (inverse inertia, accumulated velocity change, right-hand sides) is kept in
// registers by substep() itself.
template <class T>
struct Contacts {
  static constexpr int N = T::MD::N, MAXC = T::MAXC, MAXAC = XARM_MAXAC;
  int nc, nac, npair;                    // contact points, points on gripper links, pairs that produced points
  uint8_t ba[MAXC], bb[MAXC];            // body codes of side A / side B (normal points from B to A)
  uint8_t pair[MAXC];                    // which collision pair produced the point (the points of one pair share the normal)
  int8_t slot[MAXC];                     // row-pool slot of the arm side (-1: none)
  int8_t o1[MAXC], o2[MAXC];             // object index of the first / second object side (-1: none); o2 only for object-object
  float s1[MAXC];                        // sign of the first object side (+1 side A, -1 side B); the second is always -1
  V3 pa[MAXC], pb[MAXC];                 // world contact points (row setup only)
  float depth[MAXC], mu[MAXC], erp[MAXC], cfm0[MAXC];
  V3 dir[MAXC][3];                       // n, t1, t2
  V3 Jo1[MAXC][3], dVo1[MAXC][3];        // first object side: sign * (r x d), sign * Iinv (r x d)
  V3 Jo2[T::NOBJ > 1 ? MAXC : 1][3], dVo2[T::NOBJ > 1 ? MAXC : 1][3];
  float jdoor[T::HAS_DOOR ? MAXC : 1][3];  // door side: sign * axis . d
  float rhs[MAXC][3], dinv[MAXC][3], app[MAXC][3], cfmr[MAXC];
  float Jarm[MAXAC][3][N], dVarm[MAXAC][3][N];  // arm side: sign * J, Minv (sign * J)
};

template <class T>
XD Box arm_box(const ArmDyn<typename T::MD>& D, int which) {  // 0 hand, 1 finger1, 2 finger2
  Box b;
  b.R = D.Rh;
  if (which == 0) { b.c = D.ph + D.Rh * v3(c_pd_hc[0], c_pd_hc[1], c_pd_hc[2]); b.h = v3(c_pd_hh[0], c_pd_hh[1], c_pd_hh[2]); }
  else if (which == 1) { b.c = D.pf1 + D.Rh * v3(c_pd_f1c[0], c_pd_f1c[1], c_pd_f1c[2]); b.h = v3(c_pd_f1h[0], c_pd_f1h[1], c_pd_f1h[2]); }
  else { b.c = D.pf2 + D.Rh * v3(c_pd_f2c[0], c_pd_f2c[1], c_pd_f2c[2]); b.h = v3(c_pd_f2h[0], c_pd_f2h[1], c_pd_f2h[2]); }
  return b;
}

// collide one pair and append its points (normal from B to A); returns the number of points added
template <class T>
XD int add_pair(Contacts<T>& C, const Box& A, const Box& B, int ca, int cb, float fa, float fb, bool soft, float erp_override) {
  V3 d = A.c - B.c;
  float ra = norm(A.h), rb = norm(B.h), rr = ra + rb + (float)XARM_CONTACT_MARGIN;
  if (dot(d, d) > rr * rr) return 0;
  int room = T::MAXC - C.nc;
  const bool with_arm = bc_is_arm(ca) || bc_is_arm(cb);
  if (with_arm && XARM_MAXAC - C.nac < room) room = XARM_MAXAC - C.nac;
  if (room <= 0) return 0;
  CPoint pts[4];
  int k = box_box(A, B, pts, room < 4 ? room : 4);
  if (k > 0) C.npair++;
  float mu = fminf(fa * fb, (float)XARM_MAX_FRICTION);
  float erp = (float)XARM_ERP2, cfm = 0.f;
  if (erp_override >= 0.f) erp = erp_override;
  if (soft) {
    const float h = (float)T::H, ks = (float)XARM_FINGER_STIFFNESS, kd = (float)XARM_FINGER_DAMPING;
    cfm = 1.f / (h * ks + kd); erp = h * ks / (h * ks + kd); cfm /= h;
  }
  for (int i = 0; i < k; i++) {
    int c = C.nc++;
    C.ba[c] = (uint8_t)ca; C.bb[c] = (uint8_t)cb; C.pair[c] = (uint8_t)C.npair;
    C.pa[c] = pts[i].pa; C.pb[c] = pts[i].pb; C.dir[c][0] = pts[i].n; C.depth[c] = pts[i].depth;
    C.mu[c] = mu; C.erp[c] = erp; C.cfm0[c] = cfm;
    C.slot[c] = with_arm ? (int8_t)(C.nac++) : (int8_t)-1;
  }
  return k;
}

// One internal substep: collide -> unconstrained velocities -> rows -> PGS -> integrate.
// ------------------------------------------------------------------------------------------------ substep pieces
// A substep is split into three pieces so that the CUDA path can run them as separate small kernels (the fused
// per-env kernel is instruction-fetch bound, profiles/): sub_setup -> sub_solve_light / sub_solve_generic -> sub_integrate.
// tests/hostsim and the fused fallback compose them back into substep().
template <class T>
struct ArmRows {  // non-contact rows of the PGS, per arm (btMultiBodyJointMotor / JointLimit / Gear) + the door rows
  static constexpr int N = T::MD::N, NA = T::NARM, NT = N * (N + 1) / 2;
  float Mi[NA][NT], mrhs[NA][N], lrhs[NA][N];
  uint32_t lim_lo[NA], lim_hi[NA];
  float grhs[NA], gdinv[NA], gden[NA];
  float dl_rhs, dl_sign, dm_rhs;
  int door_lim;
};
template <class T>
struct SubBase {  // unconstrained velocities of the substep (what the solved changes are added to)
  float qdu[T::NARM][T::MD::N];
  V3 vu[T::NOBJ > 0 ? T::NOBJ : 1], wu[T::NOBJ > 0 ? T::NOBJ : 1];
  float door_qdu;
};
template <class T>
struct SubSol {  // velocity changes produced by the solver
  float dqd[T::NARM][T::MD::N];
  V3 dv[T::NOBJ > 0 ? T::NOBJ : 1], dw[T::NOBJ > 0 ? T::NOBJ : 1];
  float ddoor;
};
enum { SOLVE_LIGHT = 0, SOLVE_GENERIC_DECOUPLED = 1, SOLVE_GENERIC_JOINT = 2 };
// What the light solver needs of a single manifold (<= 4 points sharing the normal) of object 0 against a static box:
// lever arms, right-hand sides and the box's world inverse inertia; the rows themselves are rebuilt from these
// (manifold_rows) so that only 34 words per env travel between the setup and the solve kernels.
struct ManifoldIn {
  V3 n;
  float mu;
  V3 r[4];
  float rhs[4][3];
  S3 Iinv;
};

// arm / door rows (btMultiBodyJointLimitConstraint, JointMotor, GearConstraint) and the arms' unconstrained
// velocities, from the per-arm dynamics pass
template <class T>
XD void arm_rows(const Env<T>& e, const ArmDyn<typename T::MD>* D, float door_qdu, ArmRows<T>& AR, SubBase<T>& B) {
  using MD = typename T::MD;
  constexpr int N = MD::N, NA = T::NARM, NT = N * (N + 1) / 2;
  const float h = (float)T::H;
  const float gr = (float)XARM_GEAR_RATIO;
#pragma unroll 1
  for (int a = 0; a < NA; a++) {
    const ArmState<MD>& st = e.arm[a];
    for (int i = 0; i < NT; i++) AR.Mi[a][i] = D[a].Minv[i];
    uint32_t lo = 0, hi = 0;
    for (int i = 0; i < N; i++) {
      const float den = D[a].Minv[tri(i, i)], qdu = D[a].qdu[i];
      B.qdu[a][i] = qdu;
      float pen_lo = st.q[i] - MD::lo(i), pen_hi = MD::hi(i) - st.q[i];
      float lr = 0.f;
      if (pen_lo <= 0.f) { lo |= 1u << i; lr = (-pen_lo * (float)XARM_ERP / h - qdu) / den; }
      else if (pen_hi <= 0.f) { hi |= 1u << i; lr = (-pen_hi * (float)XARM_ERP / h + qdu) / den; }
      AR.lrhs[a][i] = lr;
      float target = (float)(XARM_MOTOR_KP * XARM_MOTOR_ERP) * (st.qt[i] - st.q[i]) / h;
      AR.mrhs[a][i] = (target - qdu) / den;
    }
    AR.lim_lo[a] = lo; AR.lim_hi[a] = hi;
    AR.grhs[a] = 0.f; AR.gdinv[a] = 0.f; AR.gden[a] = 0.f;
    if (MD::HAS_GEAR) {
      const int f1 = MD::F1, f2 = MD::F2 < 0 ? 0 : MD::F2;
      float den = D[a].Minv[tri(f1, f1)] + 2.f * gr * D[a].Minv[tri(f1, f2)] + gr * gr * D[a].Minv[tri(f2, f2)];
      float rel = D[a].qdu[f1] + gr * D[a].qdu[f2];
      float pos_err = (float)XARM_GEAR_ERP * (st.q[f1] + gr * st.q[f2]);
      AR.gdinv[a] = 1.f / den; AR.gden[a] = den;
      AR.grhs[a] = (-pos_err * (float)XARM_ERP / h - rel) / den;
    }
  }
  AR.door_lim = 0; AR.dl_rhs = 0.f; AR.dl_sign = 1.f; AR.dm_rhs = 0.f;
  B.door_qdu = door_qdu;
  if (T::HAS_DOOR) {
    const float door_den = 1.f / (float)XARM_DOOR_MASS;
    float pen_lo = e.door_q - (float)XARM_DOOR_LIMIT_LO, pen_hi = (float)XARM_DOOR_LIMIT_HI - e.door_q;
    if (pen_lo <= 0.f) { AR.door_lim = 1; AR.dl_sign = 1.f; AR.dl_rhs = (-pen_lo * (float)XARM_ERP / h - door_qdu) / door_den; }
    else if (pen_hi <= 0.f) { AR.door_lim = 1; AR.dl_sign = -1.f; AR.dl_rhs = (-pen_hi * (float)XARM_ERP / h + door_qdu) / door_den; }
    AR.dm_rhs = (0.f - door_qdu) / door_den;
  }
}

// collide -> unconstrained velocities -> rows.  Returns which solver form applies.
template <class T>
XD int sub_setup(Env<T>& e, bool apply_damping, bool last, ArmRows<T>& AR, Contacts<T>& C, SubBase<T>& B, ManifoldIn& MI,
                 const ArmDyn<typename T::MD>* Dpre = nullptr, S3* Iinv_out = nullptr) {
  using MD = typename T::MD;
  constexpr int N = MD::N, NA = T::NARM, NOBJ = T::NOBJ, NO = NOBJ > 0 ? NOBJ : 1, NT = N * (N + 1) / 2;
  const float h = (float)T::H;
  ArmDyn<MD> D[NA];
  if (Dpre) {  // the dynamics pass of this substep was already made (lean setup of the pipeline)
#pragma unroll
    for (int a = 0; a < NA; a++) D[a] = Dpre[a];
  } else {
#pragma unroll
    for (int a = 0; a < NA; a++) arm_dynamics<T>(a, e.arm[a], apply_damping, D[a]);
  }

  // ---- 1. collision detection on the current poses (fixed pair order, Appendix G)
  C.nc = 0; C.nac = 0; C.npair = 0;
  S3 Iinv[NO];
  if (NOBJ > 0) {
    Box ob[NO];
    for (int o = 0; o < NOBJ; o++) {
      ob[o].c = e.obj[o].pos; ob[o].R = quat_to_m3(e.obj[o].quat); ob[o].h = v3(T::OBJ_HX, T::OBJ_HY, T::OBJ_HZ);
      // world inverse inertia of the box (inertia from its own shape: m/12 (ly^2+lz^2), ...)
      const float lx = 2 * T::OBJ_HX, ly = 2 * T::OBJ_HY, lz = 2 * T::OBJ_HZ, m12 = T::OBJ_MASS / 12.f;
      S3 Il = {1.f / (m12 * (ly * ly + lz * lz)), 0, 0, 1.f / (m12 * (lx * lx + lz * lz)), 0, 1.f / (m12 * (lx * lx + ly * ly))};
      Iinv[o] = rotate_sym(ob[o].R, Il);
      if (Iinv_out) Iinv_out[o] = Iinv[o];
    }
    Box tb[T::NTABLE];
    for (int k = 0; k < T::NTABLE; k++) {
      tb[k].c = v3(T::table_x(k), 0.f, -(float)XARM_TABLE_HALF_Z); tb[k].R = m3_identity();
      tb[k].h = v3((float)XARM_TABLE_HALF_X, (float)XARM_TABLE_HALF_Y, (float)XARM_TABLE_HALF_Z);
    }
    Box ground;
    ground.c = v3(0.f, 0.f, (float)XARM_GROUND_Z - 5.f); ground.R = m3_identity(); ground.h = v3(100.f, 100.f, 5.f);
    const float fo = (float)XARM_DEFAULT_FRICTION;
    for (int o = 0; o < NOBJ; o++) {
      for (int k = 0; k < T::NTABLE; k++) add_pair<T>(C, ob[o], tb[k], BC_OBJ0 + o, BC_STATIC, fo, (float)XARM_TABLE_FRICTION, false, -1.f);
      if (T::HAS_GROUND) add_pair<T>(C, ob[o], ground, BC_OBJ0 + o, BC_STATIC, fo, 1.0f, false, -1.f);
    }
    for (int o = 0; o < NOBJ; o++)
      for (int p2 = o + 1; p2 < NOBJ; p2++) add_pair<T>(C, ob[o], ob[p2], BC_OBJ0 + o, BC_OBJ0 + p2, fo, fo, false, -1.f);
    int gcount[2][2][NO];
    if (MD::HAS_BOXES) {
      for (int a = 0; a < NA; a++) {
        const float ff = (T::FRICTION_SWITCH && e.grasp_cmd[a]) ? (float)XARM_FINGER_FRICTION_GRASP : (float)XARM_FINGER_FRICTION_FREE;
        Box f1 = arm_box<T>(D[a], 1), f2 = arm_box<T>(D[a], 2), hd = arm_box<T>(D[a], 0);
        const int c0 = a == 0 ? BC_ARM0_HAND : BC_ARM1_HAND;
        for (int o = 0; o < NOBJ; o++) {
          gcount[a][0][o] = add_pair<T>(C, f1, ob[o], c0 + 1, BC_OBJ0 + o, ff, fo, true, -1.f);
          gcount[a][1][o] = add_pair<T>(C, f2, ob[o], c0 + 2, BC_OBJ0 + o, ff, fo, true, -1.f);
          add_pair<T>(C, hd, ob[o], c0, BC_OBJ0 + o, fo, fo, false, -1.f);
        }
      }
    }
    if (T::HAS_DOOR) {
      const float b1[3] = XARM_DOOR_FIXED_BAR1, b2[3] = XARM_DOOR_FIXED_BAR2, org[3] = XARM_DOOR_ORIGIN, bh[3] = XARM_DOOR_BAR_HALF;
      Box bar[3];
      for (int b = 0; b < 3; b++) { bar[b].R = m3_identity(); bar[b].h = v3(bh[0], bh[1], bh[2]); }
      bar[0].c = v3(b1[0], b1[1], b1[2]); bar[1].c = v3(b2[0], b2[1], b2[2]);
      bar[2].c = v3(org[0], org[1], org[2]) + e.door_q * door_axis();
      const float fd = (float)XARM_DOOR_FRICTION;
      for (int o = 0; o < NOBJ; o++)
        for (int b = 0; b < 3; b++) add_pair<T>(C, ob[o], bar[b], BC_OBJ0 + o, b == 2 ? BC_DOOR : BC_STATIC, fo, fd, false, 0.f);
      for (int a = 0; a < NA; a++) {
        const float ff = (float)XARM_FINGER_FRICTION_FREE;
        Box f1 = arm_box<T>(D[a], 1), f2 = arm_box<T>(D[a], 2), hd = arm_box<T>(D[a], 0);
        const int c0 = a == 0 ? BC_ARM0_HAND : BC_ARM1_HAND;
        for (int b = 0; b < 3; b++) {
          const int cb = b == 2 ? BC_DOOR : BC_STATIC;
          add_pair<T>(C, f1, bar[b], c0 + 1, cb, ff, fd, true, 0.f);
          add_pair<T>(C, f2, bar[b], c0 + 2, cb, ff, fd, true, 0.f);
          add_pair<T>(C, hd, bar[b], c0, cb, fo, fd, false, 0.f);
        }
      }
    }
    if (T::FINGER_TABLE) {
      for (int a = 0; a < NA; a++) {
        const float ff = (T::FRICTION_SWITCH && e.grasp_cmd[a]) ? (float)XARM_FINGER_FRICTION_GRASP : (float)XARM_FINGER_FRICTION_FREE;
        Box f1 = arm_box<T>(D[a], 1), f2 = arm_box<T>(D[a], 2);
        const int c0 = a == 0 ? BC_ARM0_HAND : BC_ARM1_HAND;
        for (int k = 0; k < T::NTABLE; k++) {
          add_pair<T>(C, f1, tb[k], c0 + 1, BC_STATIC, ff, (float)XARM_TABLE_FRICTION, true, -1.f);
          add_pair<T>(C, f2, tb[k], c0 + 2, BC_STATIC, ff, (float)XARM_TABLE_FRICTION, true, -1.f);
        }
      }
    }
    if (last && MD::HAS_BOXES) {  // grasp flags from this (the last) collision pass
      for (int a = 0; a < NA; a++) {
        bool g = false;
        for (int o = 0; o < NOBJ; o++) {
          if (T::TASK == XARM_TASK_HANDOVER && o > 0) break;
          g = g || (gcount[a][0][o] > 0 && gcount[a][1][o] > 0);
        }
        e.grasp[a] = g;
      }
    }
  }

#ifdef XARM_HOST_SIM
  if (getenv("XARM_TRACE")) for (int c = 0; c < C.nc; c++)
    fprintf(stderr, "K c%d (%d,%d) pa %.6f %.6f %.6f pb %.6f %.6f %.6f n %.4f %.4f %.4f d %.6f\n", c, (int)C.ba[c], (int)C.bb[c], (double)C.pa[c].x, (double)C.pa[c].y, (double)C.pa[c].z,
            (double)C.pb[c].x, (double)C.pb[c].y, (double)C.pb[c].z, (double)C.dir[c][0].x, (double)C.dir[c][0].y, (double)C.dir[c][0].z, (double)C.depth[c]);
#endif
  // ---- 2. unconstrained velocities of the free bodies (the arms' are in D[a].qdu)
  V3 (&vu)[NO] = B.vu; V3 (&wu)[NO] = B.wu;
  float door_qdu = 0.f;
  for (int o = 0; o < NOBJ; o++) {
    const ObjState& b = e.obj[o];
    float kl = (float)XARM_MB_LINEAR_DAMPING * (1.f + norm(b.v)), ka = (float)XARM_MB_ANGULAR_DAMPING * (1.f + norm(b.w));
    vu[o] = b.v + h * ((-kl) * b.v); vu[o].z -= h * (float)XARM_GRAVITY;
    wu[o] = b.w + h * ((-ka) * b.w);
  }
  if (T::HAS_DOOR) {
    float v = e.door_qd;
    float f = (apply_damping ? -(float)XARM_DOOR_DAMPING * v : 0.f) - (float)XARM_DOOR_MASS * v * (float)XARM_MB_LINEAR_DAMPING * (1.f + fabsf(v));
    door_qdu = v + h * f / (float)XARM_DOOR_MASS;
  }

  // ---- 3. contact rows (setupMultiBodyContactConstraint): normal row and two friction rows per point
  for (int c = 0; c < C.nc; c++) {
    V3 t1, t2;
    plane_space(C.dir[c][0], t1, t2);
    C.dir[c][1] = t1; C.dir[c][2] = t2;
    const int ca = C.ba[c], cb = C.bb[c];
    int o1 = -1, o2 = -1; float s1 = 1.f;
    if (bc_is_obj(ca)) { o1 = NOBJ <= 1 ? 0 : ca - BC_OBJ0; s1 = 1.f; if (bc_is_obj(cb)) o2 = NOBJ <= 1 ? 0 : cb - BC_OBJ0; }
    else if (bc_is_obj(cb)) { o1 = NOBJ <= 1 ? 0 : cb - BC_OBJ0; s1 = -1.f; }
    C.o1[c] = (int8_t)o1; C.o2[c] = (int8_t)o2; C.s1[c] = s1;
    for (int k = 0; k < 3; k++) {
      const V3 d = C.dir[c][k];
      float den = 0.f, rel = 0.f;
      if (o1 >= 0) {
        const int o = NOBJ <= 1 ? 0 : o1;
        V3 r = (s1 > 0.f ? C.pa[c] : C.pb[c]) - e.obj[o].pos;
        V3 rxd = cross(r, d), ir = Iinv[o] * rxd;
        C.Jo1[c][k] = s1 * rxd; C.dVo1[c][k] = s1 * ir;
        den += 1.f / T::OBJ_MASS + dot(rxd, ir);
        rel += s1 * (dot(d, vu[o]) + dot(rxd, wu[o]));
      }
      if (NOBJ > 1 && o2 >= 0) {
        V3 r = C.pb[c] - e.obj[o2].pos;
        V3 rxd = cross(r, d), ir = Iinv[o2] * rxd;
        C.Jo2[NOBJ > 1 ? c : 0][k] = -rxd; C.dVo2[NOBJ > 1 ? c : 0][k] = -ir;
        den += 1.f / T::OBJ_MASS + dot(rxd, ir);
        rel -= dot(d, vu[o2]) + dot(rxd, wu[o2]);
      }
      if (C.slot[c] >= 0) {  // gripper link side: J_j = S_j . [p x d ; d] over the link's ancestors, response Minv J
        const int code = bc_is_arm(ca) ? ca : cb;
        const float sign = bc_is_arm(ca) ? 1.f : -1.f;
        const V3 p = bc_is_arm(ca) ? C.pa[c] : C.pb[c];
        const int a = NA == 1 ? 0 : bc_arm(code);
        const int which = (code - BC_ARM0_HAND) % 3;
        const int sl = C.slot[c];
        const ArmDyn<MD>& Da = D[a];
        V3 pxd = cross(p, d);
        float J[N];
#pragma unroll
        for (int j = 0; j < N; j++) {
          bool on = j < 7 || (which == 1 && j == MD::F1) || (which == 2 && j == MD::F2);
          J[j] = on ? sign * (dot(Da.S[j].a, pxd) + dot(Da.S[j].l, d)) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < N; i++) {
          float sacc = 0.f;
#pragma unroll
          for (int j = 0; j < N; j++) sacc += Da.Minv[tri(i, j)] * J[j];
          C.Jarm[sl][k][i] = J[i];
          C.dVarm[sl][k][i] = sacc;
          den += J[i] * sacc;
          rel += J[i] * Da.qdu[i];
        }
      }
      if (T::HAS_DOOR) {
        float jd = 0.f;
        if (ca == BC_DOOR) jd = dot(door_axis(), d); else if (cb == BC_DOOR) jd = -dot(door_axis(), d);
        C.jdoor[T::HAS_DOOR ? c : 0][k] = jd;
        den += jd * jd / (float)XARM_DOOR_MASS;
        rel += jd * door_qdu;
      }
      C.app[c][k] = 0.f;
      if (k == 0) {
        float dinv = 1.f / (den + C.cfm0[c]);
        float pen = -C.depth[c] + (float)XARM_LINEAR_SLOP;
        float pos_err = 0.f, vel_err = -rel;
        if (pen > 0.f) vel_err -= pen / h; else pos_err = -pen * C.erp[c] / h;
        C.rhs[c][0] = (pos_err + vel_err) * dinv;
        C.dinv[c][0] = dinv;
        C.cfmr[c] = C.cfm0[c] * dinv;
      } else {
        float dinv = 1.f / den;
        C.rhs[c][k] = -rel * dinv;
        C.dinv[c][k] = dinv;
      }
    }
  }

  // ---- 4. arm / door rows
  arm_rows<T>(e, D, door_qdu, AR, B);
  // which solver form: islands (one arm, no door, no gripper contact) decouple the arm rows from the object rows
  const bool decoupled = NA == 1 && !T::HAS_DOOR && C.nac == 0;
  const bool manifold = decoupled && NOBJ == 1 && C.nc > 0 && C.nc <= 4 && C.npair == 1 && C.o1[0] == 0 && C.s1[0] > 0.f && C.cfm0[0] == 0.f;
  if (manifold) {
    MI.n = C.dir[0][0]; MI.mu = C.mu[0]; MI.Iinv = Iinv[0];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const bool on = c < C.nc;
      MI.r[c] = on ? C.pa[c] - e.obj[0].pos : v3(0, 0, 0);
#pragma unroll
      for (int k = 0; k < 3; k++) MI.rhs[c][k] = on ? C.rhs[c][k] : 0.f;
    }
  }
  if (decoupled && (C.nc == 0 || manifold)) return SOLVE_LIGHT;
  return decoupled ? SOLVE_GENERIC_DECOUPLED : SOLVE_GENERIC_JOINT;
}

// Lean form of sub_setup for the tasks with one arm, no door and at most one object (Reach, PickAndPlace): the common
// case there is "the gripper touches nothing and the object rests on (or falls towards) one static box", i.e. the
// light solver form.  This version never builds the Contacts record (5.8 KB of thread-local memory per env - in the
// batched setup kernel those writes alone were 720 MB of DRAM traffic per launch): the one object-static manifold goes
// straight into ManifoldIn, the gripper pairs are only tested for ANY contact.  Returns false when the env is not of
// the light form (the caller then takes the generic path: sub_setup with a Contacts record); true with nc = number of
// manifold points otherwise.  Row arithmetic = sub_setup's row loop (setupMultiBodyContactConstraint).
template <class T, bool FLAT = false>
XD bool sub_setup_lean(Env<T>& e, bool apply_damping, bool last, ArmRows<T>& AR, SubBase<T>& B, ManifoldIn& MI, int& nc_out,
                       ArmDyn<typename T::MD>* D) {  // D[1]: the dynamics pass, also valid when the result is false
  using MD = typename T::MD;
  static_assert(T::NARM == 1 && !T::HAS_DOOR && T::NOBJ <= 1, "lean setup: one arm, no door, at most one object");
  const float h = (float)T::H;
  arm_dynamics<T, FLAT>(0, e.arm[0], apply_damping, D[0]);
  nc_out = 0;
  if (T::NOBJ == 1) {
    Box ob;
    ob.c = e.obj[0].pos; ob.R = quat_to_m3(e.obj[0].quat); ob.h = v3(T::OBJ_HX, T::OBJ_HY, T::OBJ_HZ);
    const float r_ob = norm(ob.h);
    // gripper links against the object: any contact point makes the env heavy
    if (MD::HAS_BOXES) {
#pragma unroll 1
      for (int which = 1; which <= 3; which++) {
        const Box g = arm_box<T>(D[0], which == 3 ? 0 : which);
        const V3 d = g.c - ob.c;
        const float rr = norm(g.h) + r_ob + (float)XARM_CONTACT_MARGIN;
        if (dot(d, d) > rr * rr) continue;
        CPoint one[1];
        if (box_box(g, ob, one, 1) > 0) return false;
      }
    }
    // object against the static boxes (tables, ground): at most one of them may produce points
    CPoint pts[4];
    int npts = 0, npair = 0;
#pragma unroll 1
    for (int k = 0; k < T::NTABLE + (T::HAS_GROUND ? 1 : 0); k++) {
      Box tb;
      tb.R = m3_identity();
      if (k < T::NTABLE) {
        tb.c = v3(T::table_x(k), 0.f, -(float)XARM_TABLE_HALF_Z);
        tb.h = v3((float)XARM_TABLE_HALF_X, (float)XARM_TABLE_HALF_Y, (float)XARM_TABLE_HALF_Z);
      } else {
        tb.c = v3(0.f, 0.f, (float)XARM_GROUND_Z - 5.f); tb.h = v3(100.f, 100.f, 5.f);
      }
      const V3 d = ob.c - tb.c;
      const float rr = r_ob + norm(tb.h) + (float)XARM_CONTACT_MARGIN;
      if (dot(d, d) > rr * rr) continue;
      CPoint p4[4];
      const int kk = box_box(ob, tb, p4, 4);
      if (kk > 0) {
        if (++npair > 1) return false;
        npts = kk;
        MI.mu = fminf((float)XARM_DEFAULT_FRICTION * (k < T::NTABLE ? (float)XARM_TABLE_FRICTION : 1.0f), (float)XARM_MAX_FRICTION);
#pragma unroll
        for (int c = 0; c < 4; c++) pts[c] = p4[c];
      }
    }
    // unconstrained velocity of the object
    const ObjState& b = e.obj[0];
    const float kl = (float)XARM_MB_LINEAR_DAMPING * (1.f + norm(b.v)), ka = (float)XARM_MB_ANGULAR_DAMPING * (1.f + norm(b.w));
    V3 vu = b.v + h * ((-kl) * b.v); vu.z -= h * (float)XARM_GRAVITY;
    const V3 wu = b.w + h * ((-ka) * b.w);
    B.vu[0] = vu; B.wu[0] = wu;
    if (npts > 0) {
      const float lx = 2 * T::OBJ_HX, ly = 2 * T::OBJ_HY, lz = 2 * T::OBJ_HZ, m12 = T::OBJ_MASS / 12.f;
      const S3 Il = {1.f / (m12 * (ly * ly + lz * lz)), 0, 0, 1.f / (m12 * (lx * lx + lz * lz)), 0, 1.f / (m12 * (lx * lx + ly * ly))};
      const S3 Iinv = rotate_sym(ob.R, Il);
      MI.Iinv = Iinv;
      MI.n = pts[0].n;
      V3 t1, t2;
      plane_space(MI.n, t1, t2);
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const bool on = c < npts;
        const V3 r = on ? pts[c].pa - b.pos : v3(0, 0, 0);
        MI.r[c] = r;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const V3 d = k == 0 ? MI.n : (k == 1 ? t1 : t2);
          const V3 rxd = cross(r, d), ir = Iinv * rxd;
          float den = 0.f, rel = 0.f;
          den += 1.f / T::OBJ_MASS + dot(rxd, ir);
          rel += 1.f * (dot(d, vu) + dot(rxd, wu));
          float rhs;
          if (k == 0) {
            const float dinv = 1.f / (den + 0.f);
            const float pen = -pts[c].depth + (float)XARM_LINEAR_SLOP;
            float pos_err = 0.f, vel_err = -rel;
            if (pen > 0.f) vel_err -= pen / h; else pos_err = -pen * (float)XARM_ERP2 / h;
            rhs = (pos_err + vel_err) * dinv;
          } else {
            rhs = -rel * (1.f / den);
          }
          MI.rhs[c][k] = on ? rhs : 0.f;
        }
      }
    }
    nc_out = npts;
    if (last && MD::HAS_BOXES) e.grasp[0] = 0;  // no gripper contact in this (the last) collision pass
  } else {
    B.vu[0] = v3(0, 0, 0); B.wu[0] = v3(0, 0, 0);
  }
  arm_rows<T>(e, D, 0.f, AR, B);
  return true;
}

// Lean setup for every other task shape (two arms, several objects, the door): the light form there is "no gripper link
// touches anything, no object touches another object or a door bar, and every object rests on (or falls towards) at most
// one static box".  Then every arm, the door and every object is an island of its own (section 5.4 of DESIGN.md): their rows
// never read each other's velocities.  Same pair set, culls and row arithmetic as sub_setup; any point on a pair outside
// "object x static box" sends the env through the generic path (return false; D[] stays valid).
template <class T, bool FLAT = false>
XD bool sub_setup_lean_multi(Env<T>& e, bool apply_damping, bool last, ArmRows<T>& AR, SubBase<T>& B, ManifoldIn* MI, int* nc_out,
                             ArmDyn<typename T::MD>* D) {  // MI[NO], nc_out[NO], D[NARM]
  using MD = typename T::MD;
  constexpr int NA = T::NARM, NOBJ = T::NOBJ, NO = NOBJ > 0 ? NOBJ : 1;
  const float h = (float)T::H;
#pragma unroll 1
  for (int a = 0; a < NA; a++) arm_dynamics<T, FLAT>(a, e.arm[a], apply_damping, D[a]);
#pragma unroll
  for (int o = 0; o < NO; o++) nc_out[o] = 0;
  float door_qdu = 0.f;
  if (T::HAS_DOOR) {
    const float v = e.door_qd;
    const float f = (apply_damping ? -(float)XARM_DOOR_DAMPING * v : 0.f) - (float)XARM_DOOR_MASS * v * (float)XARM_MB_LINEAR_DAMPING * (1.f + fabsf(v));
    door_qdu = v + h * f / (float)XARM_DOOR_MASS;
  }
  if (NOBJ > 0) {
    Box ob[NO];
#pragma unroll 1
    for (int o = 0; o < NOBJ; o++) { ob[o].c = e.obj[o].pos; ob[o].R = quat_to_m3(e.obj[o].quat); ob[o].h = v3(T::OBJ_HX, T::OBJ_HY, T::OBJ_HZ); }
    const float r_ob = norm(ob[0].h);
    CPoint one[1];
    // ---- pairs whose points make the env heavy: object - object, gripper - object, anything - door bars, finger - table
#pragma unroll 1
    for (int o = 0; o < NOBJ; o++)
#pragma unroll 1
      for (int p2 = o + 1; p2 < NOBJ; p2++) {
        const V3 d = ob[o].c - ob[p2].c;
        const float rr = r_ob + r_ob + (float)XARM_CONTACT_MARGIN;
        if (dot(d, d) > rr * rr) continue;
        if (box_box(ob[o], ob[p2], one, 1) > 0) return false;
      }
    Box bar[3];
    if (T::HAS_DOOR) {
      const float b1[3] = XARM_DOOR_FIXED_BAR1, b2[3] = XARM_DOOR_FIXED_BAR2, org[3] = XARM_DOOR_ORIGIN, bh[3] = XARM_DOOR_BAR_HALF;
      for (int b = 0; b < 3; b++) { bar[b].R = m3_identity(); bar[b].h = v3(bh[0], bh[1], bh[2]); }
      bar[0].c = v3(b1[0], b1[1], b1[2]); bar[1].c = v3(b2[0], b2[1], b2[2]);
      bar[2].c = v3(org[0], org[1], org[2]) + e.door_q * door_axis();
      const float r_bar = norm(bar[0].h);
#pragma unroll 1
      for (int o = 0; o < NOBJ; o++)
#pragma unroll 1
        for (int b = 0; b < 3; b++) {
          const V3 d = ob[o].c - bar[b].c;
          const float rr = r_ob + r_bar + (float)XARM_CONTACT_MARGIN;
          if (dot(d, d) > rr * rr) continue;
          if (box_box(ob[o], bar[b], one, 1) > 0) return false;
        }
    }
    if (MD::HAS_BOXES) {
#pragma unroll 1
      for (int a = 0; a < NA; a++)
#pragma unroll 1
        for (int which = 1; which <= 3; which++) {
          const Box g = arm_box<T>(D[a], which == 3 ? 0 : which);
          const float r_g = norm(g.h);
#pragma unroll 1
          for (int o = 0; o < NOBJ; o++) {
            const V3 d = g.c - ob[o].c;
            const float rr = r_g + r_ob + (float)XARM_CONTACT_MARGIN;
            if (dot(d, d) > rr * rr) continue;
            if (box_box(g, ob[o], one, 1) > 0) return false;
          }
          if (T::HAS_DOOR) {
            const float r_bar = norm(bar[0].h);
#pragma unroll 1
            for (int b = 0; b < 3; b++) {
              const V3 d = g.c - bar[b].c;
              const float rr = r_g + r_bar + (float)XARM_CONTACT_MARGIN;
              if (dot(d, d) > rr * rr) continue;
              if (box_box(g, bar[b], one, 1) > 0) return false;
            }
          }
          if (T::FINGER_TABLE && which != 3) {
#pragma unroll 1
            for (int k = 0; k < T::NTABLE; k++) {
              Box tb;
              tb.R = m3_identity();
              tb.c = v3(T::table_x(k), 0.f, -(float)XARM_TABLE_HALF_Z);
              tb.h = v3((float)XARM_TABLE_HALF_X, (float)XARM_TABLE_HALF_Y, (float)XARM_TABLE_HALF_Z);
              const V3 d = g.c - tb.c;
              const float rr = r_g + norm(tb.h) + (float)XARM_CONTACT_MARGIN;
              if (dot(d, d) > rr * rr) continue;
              if (box_box(g, tb, one, 1) > 0) return false;
            }
          }
        }
    }
    // ---- every object against the static boxes (tables, ground): at most one of them may produce points
#pragma unroll 1
    for (int o = 0; o < NOBJ; o++) {
      CPoint pts[4];
      int npts = 0, npair = 0;
      float mu = 0.f;
#pragma unroll 1
      for (int k = 0; k < T::NTABLE + (T::HAS_GROUND ? 1 : 0); k++) {
        Box tb;
        tb.R = m3_identity();
        if (k < T::NTABLE) {
          tb.c = v3(T::table_x(k), 0.f, -(float)XARM_TABLE_HALF_Z);
          tb.h = v3((float)XARM_TABLE_HALF_X, (float)XARM_TABLE_HALF_Y, (float)XARM_TABLE_HALF_Z);
        } else {
          tb.c = v3(0.f, 0.f, (float)XARM_GROUND_Z - 5.f); tb.h = v3(100.f, 100.f, 5.f);
        }
        const V3 d = ob[o].c - tb.c;
        const float rr = r_ob + norm(tb.h) + (float)XARM_CONTACT_MARGIN;
        if (dot(d, d) > rr * rr) continue;
        CPoint p4[4];
        const int kk = box_box(ob[o], tb, p4, 4);
        if (kk > 0) {
          if (++npair > 1) return false;
          npts = kk;
          mu = fminf((float)XARM_DEFAULT_FRICTION * (k < T::NTABLE ? (float)XARM_TABLE_FRICTION : 1.0f), (float)XARM_MAX_FRICTION);
#pragma unroll
          for (int c = 0; c < 4; c++) pts[c] = p4[c];
        }
      }
      const ObjState& b = e.obj[o];
      const float kl = (float)XARM_MB_LINEAR_DAMPING * (1.f + norm(b.v)), ka = (float)XARM_MB_ANGULAR_DAMPING * (1.f + norm(b.w));
      V3 vu = b.v + h * ((-kl) * b.v); vu.z -= h * (float)XARM_GRAVITY;
      const V3 wu = b.w + h * ((-ka) * b.w);
      B.vu[o] = vu; B.wu[o] = wu;
      if (npts > 0) {
        ManifoldIn& M = MI[o];
        M.mu = mu;
        const float lx = 2 * T::OBJ_HX, ly = 2 * T::OBJ_HY, lz = 2 * T::OBJ_HZ, m12 = T::OBJ_MASS / 12.f;
        const S3 Il = {1.f / (m12 * (ly * ly + lz * lz)), 0, 0, 1.f / (m12 * (lx * lx + lz * lz)), 0, 1.f / (m12 * (lx * lx + ly * ly))};
        const S3 Iinv = rotate_sym(ob[o].R, Il);
        M.Iinv = Iinv;
        M.n = pts[0].n;
        V3 t1, t2;
        plane_space(M.n, t1, t2);
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const bool on = c < npts;
          const V3 r = on ? pts[c].pa - b.pos : v3(0, 0, 0);
          M.r[c] = r;
#pragma unroll
          for (int k = 0; k < 3; k++) {
            const V3 d = k == 0 ? M.n : (k == 1 ? t1 : t2);
            const V3 rxd = cross(r, d), ir = Iinv * rxd;
            float den = 0.f, rel = 0.f;
            den += 1.f / T::OBJ_MASS + dot(rxd, ir);
            rel += 1.f * (dot(d, vu) + dot(rxd, wu));
            float rhs;
            if (k == 0) {
              const float dinv = 1.f / (den + 0.f);
              const float pen = -pts[c].depth + (float)XARM_LINEAR_SLOP;
              float pos_err = 0.f, vel_err = -rel;
              if (pen > 0.f) vel_err -= pen / h; else pos_err = -pen * (float)XARM_ERP2 / h;
              rhs = (pos_err + vel_err) * dinv;
            } else {
              rhs = -rel * (1.f / den);
            }
            M.rhs[c][k] = on ? rhs : 0.f;
          }
        }
      }
      nc_out[o] = npts;
    }
    if (last && MD::HAS_BOXES)
#pragma unroll
      for (int a = 0; a < NA; a++) e.grasp[a] = 0;  // no gripper contact in this (the last) collision pass
  } else {
    B.vu[0] = v3(0, 0, 0); B.wu[0] = v3(0, 0, 0);
  }
  arm_rows<T>(e, D, door_qdu, AR, B);
  return true;
}

// ---- row updates shared by the solver forms (they act on the local register arrays Mi, iden, dqd, ... of the caller)
// one unit row (J = sign * e_i) of arm a: motors and joint limits
#define UNIT_ROW(a, i, sign, rhs_, lo_, hi_, app_)                                   \
  {                                                                                   \
    float delta = (rhs_) - (sign) * dqd[a][i] * iden[a][i];                           \
    const float sum = (app_) + delta;                                                 \
    const float sumc = fminf(fmaxf(sum, (lo_)), (hi_));   /* branch-free clamp; delta is recomputed only when it bites */ \
    delta = (sumc == sum) ? delta : sumc - (app_);                                    \
    (app_) = sumc;                                                                    \
    const float sd = (sign) * delta;                                                  \
    _Pragma("unroll") for (int k_ = 0; k_ < N; k_++) dqd[a][k_] += Mi[a][tri(k_, i)] * sd; \
    resid_bad = resid_bad || fabsf(delta) > sthr_ * iden[a][i];                       \
  }
#define GEAR_ROW(a)                                                                   \
  if (MD::HAS_GEAR) {                                                                 \
    const int f1 = MD::F1, f2 = MD::F2 < 0 ? 0 : MD::F2;                              \
    float delta = grhs[a] - (dqd[a][f1] + gr * dqd[a][f2]) * gdinv[a];                \
    const float sum = gapp[a] + delta;                                                \
    const float sumc = fminf(fmaxf(sum, -hi_gear), hi_gear);                          \
    delta = (sumc == sum) ? delta : sumc - gapp[a];                                   \
    gapp[a] = sumc;                                                                   \
    _Pragma("unroll") for (int k_ = 0; k_ < N; k_++) dqd[a][k_] += (Mi[a][tri(k_, f1)] + gr * Mi[a][tri(k_, f2)]) * delta; \
    resid_bad = resid_bad || fabsf(delta) > sthr_ * gdinv[a];                         \
  }
#define DOOR_LIMIT_ROW()                                                              \
  if (door_lim) {                                                                     \
    float delta = dl_rhs - dl_sign * ddoor / door_den;                                \
    float sum = dl_app + delta;                                                       \
    if (sum < 0.f) { delta = -dl_app; sum = 0.f; } else if (sum > hi_lim) { delta = hi_lim - dl_app; sum = hi_lim; } \
    dl_app = sum; ddoor += dl_sign * delta * door_den;                                \
    resid_bad = resid_bad || fabsf(delta * door_den) > sthr_;                         \
  }
#define DOOR_MOTOR_ROW()                                                              \
  {                                                                                   \
    const float hi_ = (float)XARM_DEFAULT_MOTOR_MAX_IMPULSE;                          \
    float delta = dm_rhs - ddoor / door_den;                                          \
    float sum = dm_app + delta;                                                       \
    if (sum < -hi_) { delta = -hi_ - dm_app; sum = -hi_; } else if (sum > hi_) { delta = hi_ - dm_app; sum = hi_; } \
    dm_app = sum; ddoor += delta * door_den;                                          \
    resid_bad = resid_bad || fabsf(delta * door_den) > sthr_;                         \
  }

// the motor rows of arm a, fully unrolled (no limit row active): forward = dof 0..N-1, backward = N-1..0
#define ARM_MOTOR_ROW_(a, i) { const float hi_ = (i) < 7 ? hi_arm : hi_fin; UNIT_ROW(a, i, 1.f, mrhs[a][i], -hi_, hi_, mapp[a][i]) }
#define ARM_MOTORS_FWD(a) { _Pragma("unroll") for (int i_ = 0; i_ < N; i_++) ARM_MOTOR_ROW_(a, i_) }
#define ARM_MOTORS_BWD(a) { _Pragma("unroll") for (int i_ = N - 1; i_ >= 0; i_--) ARM_MOTOR_ROW_(a, i_) }
// the joint-limit rows of arm a (only the violated limits have a row).  Arm joints (dof < 7) rarely sit on a limit:
// their rows are behind a branch.  Gripper dofs often do (a finger driven to its stop, Reach's knuckles): their rows are
// always executed, with a zero impulse when the limit is not violated - no warp divergence in the batched kernels.
#define ARM_LIMIT_ROW_(a, i)                                                                                     \
  { if (lim_lo[a] >> (i) & 1) UNIT_ROW(a, i, 1.f, lrhs[a][i], 0.f, hi_lim, lapp[a][i])                           \
    else if (lim_hi[a] >> (i) & 1) UNIT_ROW(a, i, -1.f, lrhs[a][i], 0.f, hi_lim, lapp[a][i]) }
#define ARM_LIMIT_ROW_PRED_(a, i)                                                                                \
  {                                                                                                              \
    const bool act_ = ((lim_lo[a] | lim_hi[a]) >> (i) & 1) != 0;                                                 \
    const float sg_ = (lim_hi[a] >> (i) & 1) ? -1.f : 1.f;                                                       \
    float delta = lrhs[a][i] - sg_ * dqd[a][i] * iden[a][i];                                                     \
    const float sum = lapp[a][i] + delta;                                                                        \
    const float sumc = fminf(fmaxf(sum, 0.f), hi_lim);                                                           \
    delta = (sumc == sum) ? delta : sumc - lapp[a][i];                                                           \
    delta = act_ ? delta : 0.f;                                                                                  \
    lapp[a][i] = act_ ? sumc : lapp[a][i];                                                                       \
    const float sd = sg_ * delta;                                                                                \
    _Pragma("unroll") for (int k_ = 0; k_ < N; k_++) dqd[a][k_] += Mi[a][tri(k_, i)] * sd;                       \
    resid_bad = resid_bad || fabsf(delta) > sthr_ * iden[a][i];                                                  \
  }
#define ARM_LIMITS_FWD(a)                                                                                        \
  { if (((lim_lo[a] | lim_hi[a]) & 0x7fu) != 0u) { _Pragma("unroll") for (int i_ = 0; i_ < 7; i_++) ARM_LIMIT_ROW_(a, i_) } \
    _Pragma("unroll") for (int i_ = 7; i_ < N; i_++) ARM_LIMIT_ROW_PRED_(a, i_) }
#define ARM_LIMITS_BWD(a)                                                                                        \
  { _Pragma("unroll") for (int i_ = N - 1; i_ >= 7; i_--) ARM_LIMIT_ROW_PRED_(a, i_)                             \
    if (((lim_lo[a] | lim_hi[a]) & 0x7fu) != 0u) { _Pragma("unroll") for (int i_ = 6; i_ >= 0; i_--) ARM_LIMIT_ROW_(a, i_) } }
// light solver form: an env whose ARM joints (dof < 7) sit on a limit is sent through the heavy path by the setup kernel
// (rare: the IK keeps the arm inside its range), so only the gripper dofs' limit rows remain - 14 registers less
#define ARM_LIMITS_FWD_GRIPPER(a) { _Pragma("unroll") for (int i_ = 7; i_ < N; i_++) ARM_LIMIT_ROW_PRED_(a, i_) }
#define ARM_LIMITS_BWD_GRIPPER(a) { _Pragma("unroll") for (int i_ = N - 1; i_ >= 7; i_--) ARM_LIMIT_ROW_PRED_(a, i_) }
#define SOLVER_LOCALS_FROM(AR)                                                                                   \
  float Mi[NA][NT], iden[NA][N], dqd[NA][N], mrhs[NA][N], mapp[NA][N], lrhs[NA][N], lapp[NA][N];                 \
  uint32_t lim_lo[NA], lim_hi[NA];                                                                               \
  float grhs[NA], gapp[NA], gdinv[NA], gden[NA];                                                                 \
  const float hi_arm = (float)(T::ARM_FORCE * T::TIME_STEP), hi_fin = (float)(T::FINGER_FORCE * T::TIME_STEP);   \
  const float hi_gear = (float)(XARM_GEAR_MAX_FORCE * T::TIME_STEP), hi_lim = (float)XARM_LIMIT_MAX_IMPULSE;     \
  const float gr = (float)XARM_GEAR_RATIO;                                                                       \
  _Pragma("unroll") for (int a = 0; a < NA; a++) {                                                               \
    _Pragma("unroll") for (int i = 0; i < NT; i++) Mi[a][i] = (AR).Mi[a][i];                                     \
    _Pragma("unroll") for (int i = 0; i < N; i++) {                                                              \
      iden[a][i] = 1.f / Mi[a][tri(i, i)]; mrhs[a][i] = (AR).mrhs[a][i]; lrhs[a][i] = (AR).lrhs[a][i];           \
      mapp[a][i] = 0.f; lapp[a][i] = 0.f; dqd[a][i] = 0.f;                                                       \
    }                                                                                                            \
    lim_lo[a] = (AR).lim_lo[a]; lim_hi[a] = (AR).lim_hi[a];                                                      \
    grhs[a] = (AR).grhs[a]; gdinv[a] = (AR).gdinv[a]; gden[a] = (AR).gden[a]; gapp[a] = 0.f;                     \
  }                                                                                                              \
  const bool door_lim = (AR).door_lim != 0;                                                                      \
  float ddoor = 0.f, dl_app = 0.f, dm_app = 0.f;                                                                 \
  const float dl_rhs = (AR).dl_rhs, dl_sign = (AR).dl_sign, dm_rhs = (AR).dm_rhs;                                \
  const float door_den = 1.f / (float)XARM_DOOR_MASS;                                                            \
  const float inv_obj_mass = 1.f / T::OBJ_MASS;                                                                  \
  /* exit test of Bullet: max over rows of (delta / dinv)^2 <= threshold <=> no row has |delta| > sqrt(thr) dinv */ \
  const float sthr_ = sqrtf((float)XARM_RESIDUAL_THRESHOLD);                                                     \
  (void)door_lim; (void)dl_rhs; (void)dl_sign; (void)dm_rhs; (void)door_den; (void)hi_lim; (void)hi_gear; (void)gr; (void)inv_obj_mass;

// Register-resident rows of a single manifold (<= 4 points sharing n, t1, t2) of the one object against a static box,
// rebuilt from the compact ManifoldIn with the arithmetic of sub_setup's row loop.
template <class T>
struct ManifoldRows {
  V3 n, t1, t2;
  float mu;
  V3 Jn[4], Jt1[4], Jt2[4], Vn[4], Vt1[4], Vt2[4];
  float rn[4], r1[4], r2[4], dn[4], d1[4], d2[4];
};
template <class T>
XD void manifold_rows(const ManifoldIn& MI, int nc, ManifoldRows<T>& Mf) {
  V3 t1, t2;
  plane_space(MI.n, t1, t2);
  Mf.n = MI.n; Mf.t1 = t1; Mf.t2 = t2; Mf.mu = MI.mu;
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const bool on = c < nc;
    const V3 r = MI.r[c];
    V3 rxd, ir;
    rxd = cross(r, MI.n); ir = MI.Iinv * rxd;
    Mf.Jn[c] = rxd; Mf.Vn[c] = ir; Mf.dn[c] = on ? 1.f / ((1.f / T::OBJ_MASS + dot(rxd, ir)) + 0.f) : 0.f;
    rxd = cross(r, t1); ir = MI.Iinv * rxd;
    Mf.Jt1[c] = rxd; Mf.Vt1[c] = ir; Mf.d1[c] = on ? 1.f / (1.f / T::OBJ_MASS + dot(rxd, ir)) : 0.f;
    rxd = cross(r, t2); ir = MI.Iinv * rxd;
    Mf.Jt2[c] = rxd; Mf.Vt2[c] = ir; Mf.d2[c] = on ? 1.f / (1.f / T::OBJ_MASS + dot(rxd, ir)) : 0.f;
    Mf.rn[c] = MI.rhs[c][0]; Mf.r1[c] = MI.rhs[c][1]; Mf.r2[c] = MI.rhs[c][2];
  }
}

// Light form (SOLVE_LIGHT): the joint loop of btMultiBodyConstraintSolver::solveSingleIteration over the arm's
// non-contact rows and, if present, the one object manifold - non-contact rows (forward on odd sweeps, backward on even
// ones), normal rows, friction pairs; leave after the first sweep in which no row moved more than the residual
// threshold.  The arm rows live in registers; the manifold rows are read from `mrows` (shared memory in the kernel,
// word w of this env at mrows[w * stride]) so that both row sets fit without spilling and the two independent row
// chains of a sweep can overlap in the pipeline.
#define XARM_MROW_WORDS 96
template <class T>
XD void manifold_rows_store(const ManifoldIn& MI, int nc, float* mrows, int stride) {
  ManifoldRows<T> Mf;
  manifold_rows<T>(MI, nc, Mf);
#pragma unroll
  for (int c = 0; c < 4; c++) {
    float* m = mrows + (size_t)(24 * c) * stride;
    m[0 * stride] = Mf.Jn[c].x; m[1 * stride] = Mf.Jn[c].y; m[2 * stride] = Mf.Jn[c].z;
    m[3 * stride] = Mf.Vn[c].x; m[4 * stride] = Mf.Vn[c].y; m[5 * stride] = Mf.Vn[c].z;
    m[6 * stride] = Mf.Jt1[c].x; m[7 * stride] = Mf.Jt1[c].y; m[8 * stride] = Mf.Jt1[c].z;
    m[9 * stride] = Mf.Vt1[c].x; m[10 * stride] = Mf.Vt1[c].y; m[11 * stride] = Mf.Vt1[c].z;
    m[12 * stride] = Mf.Jt2[c].x; m[13 * stride] = Mf.Jt2[c].y; m[14 * stride] = Mf.Jt2[c].z;
    m[15 * stride] = Mf.Vt2[c].x; m[16 * stride] = Mf.Vt2[c].y; m[17 * stride] = Mf.Vt2[c].z;
    m[18 * stride] = Mf.rn[c]; m[19 * stride] = Mf.dn[c];
    m[20 * stride] = Mf.r1[c]; m[21 * stride] = Mf.d1[c];
    m[22 * stride] = Mf.r2[c]; m[23 * stride] = Mf.d2[c];
  }
}
template <class T>
XD void sub_solve_light(const ArmRows<T>& AR, int nc, const ManifoldIn& MI, float* mrows, int stride, SubSol<T>& S) {
  using MD = typename T::MD;
  constexpr int N = MD::N, NA = T::NARM, NT = N * (N + 1) / 2;
  SOLVER_LOCALS_FROM(AR)
  V3 n = v3(0, 0, 0), t1 = v3(0, 0, 0), t2 = v3(0, 0, 0), nm = n, t1m = n, t2m = n;
  float mu = 0.f;
  if (nc > 0) {
    manifold_rows_store<T>(MI, nc, mrows, stride);
    n = MI.n; mu = MI.mu;
    plane_space(n, t1, t2);
    nm = inv_obj_mass * n; t1m = inv_obj_mass * t1; t2m = inv_obj_mass * t2;
  }
  float an[4], a1[4], a2[4];
#pragma unroll
  for (int c = 0; c < 4; c++) { an[c] = 0.f; a1[c] = 0.f; a2[c] = 0.f; }
  V3 v = v3(0, 0, 0), w = v3(0, 0, 0);
#define MR(c, k) mrows[(size_t)(24 * (c) + (k)) * stride]
#define MRV(c, k) v3(MR(c, k), MR(c, (k) + 1), MR(c, (k) + 2))
  for (int it = 0; it < XARM_SOLVER_ITERATIONS; it++) {
    bool resid_bad = false;
    if (it & 1) {
      ARM_LIMITS_FWD_GRIPPER(0) ARM_MOTORS_FWD(0) GEAR_ROW(0)
    } else {
      GEAR_ROW(0) ARM_MOTORS_BWD(0) ARM_LIMITS_BWD_GRIPPER(0)
    }
    if (nc > 0) {
#pragma unroll
      for (int c = 0; c < 4; c++) {  // normal rows
        const float dn = MR(c, 19);
        float delta = MR(c, 18) - (dot(n, v) + dot(MRV(c, 0), w)) * dn;
        const float sum = an[c] + delta;
        const float sumc = fminf(fmaxf(sum, 0.f), (float)XARM_CONTACT_MAX_IMPULSE);
        delta = (sumc == sum) ? delta : sumc - an[c];
        an[c] = sumc;
        v += delta * nm; w += delta * MRV(c, 3);
        resid_bad = resid_bad || fabsf(delta) > sthr_ * dn;
      }
#pragma unroll
      for (int c = 0; c < 4; c++) {  // friction pairs, implicit cone
        const float lim = mu * an[c];
        const float d1 = MR(c, 21), d2 = MR(c, 23);
        float da = MR(c, 20) - (dot(t1, v) + dot(MRV(c, 6), w)) * d1, db = MR(c, 22) - (dot(t2, v) + dot(MRV(c, 12), w)) * d2;
        float sa = a1[c] + da, sb = a2[c] + db;
        const float l2 = sa * sa + sb * sb;
        if (l2 > lim * lim) {
          const float len = sqrtf(l2);
          if (len > lim) { const float sc = lim / len; sa *= sc; sb *= sc; da = sa - a1[c]; db = sb - a2[c]; }
        }
        a1[c] = sa; a2[c] = sb;
        v += da * t1m; w += da * MRV(c, 9);
        v += db * t2m; w += db * MRV(c, 15);
        resid_bad = resid_bad || fabsf(da) > sthr_ * d1 || fabsf(db) > sthr_ * d2;
      }
    }
    if (!resid_bad) break;
  }
#undef MR
#undef MRV
#pragma unroll
  for (int i = 0; i < N; i++) S.dqd[0][i] = dqd[0][i];
  S.dv[0] = v; S.dw[0] = w; S.ddoor = 0.f;
}

// ---- Light form of the tasks with several islands (two arms, several objects, the door): sub_setup_lean_multi guarantees
// that every arm, the door and every object with a manifold is an island - no row of one reads a velocity of another.  The
// joint loop of Bullet interleaves their rows, but a Gauss-Seidel sweep over independent islands gives every island the
// impulses it would get alone, so the islands are swept ONE AFTER THE OTHER (only one island's rows in registers at a time).
// What the islands share is the loop's exit test (max over ALL rows): every island returns a bit per sweep ("some row of
// mine moved more than the threshold"); the joint loop leaves after the first sweep whose bit is clear in every island, and
// if that happens before the last sweep the islands are swept again up to there (rare: arms and objects at rest).
// Rows of one arm in the scratch-slab order of ar_store: Mi[NT] | mrhs[N] | lrhs[N] | lim_lo lim_hi | grhs gdinv gden.
template <class T>
XHD int ar_arm_words() { return T::MD::N * (T::MD::N + 1) / 2 + 2 * T::MD::N + 5; }
template <class T>
XD uint64_t light_island_arm(const float* __restrict__ s, int64_t n, int64_t i, int sweeps, float* dqd_out) {
  using MD = typename T::MD;
  constexpr int N = MD::N, NT = N * (N + 1) / 2;
  float Mi[1][NT], iden[1][N], dqd[1][N], mrhs[1][N], mapp[1][N], lrhs[1][N], lapp[1][N];
  uint32_t lim_lo[1], lim_hi[1];
  float grhs[1], gapp[1], gdinv[1];
  const float hi_arm = (float)(T::ARM_FORCE * T::TIME_STEP), hi_fin = (float)(T::FINGER_FORCE * T::TIME_STEP);
  const float hi_gear = (float)(XARM_GEAR_MAX_FORCE * T::TIME_STEP), hi_lim = (float)XARM_LIMIT_MAX_IMPULSE;
  const float gr = (float)XARM_GEAR_RATIO;
  const float sthr_ = sqrtf((float)XARM_RESIDUAL_THRESHOLD);
  (void)hi_gear; (void)gr; (void)hi_lim;
  int w = 0;
#pragma unroll
  for (int k = 0; k < NT; k++) Mi[0][k] = s[(w++) * n + i];
#pragma unroll
  for (int k = 0; k < N; k++) { mrhs[0][k] = s[(w++) * n + i]; mapp[0][k] = 0.f; lapp[0][k] = 0.f; dqd[0][k] = 0.f; iden[0][k] = 1.f / Mi[0][tri(k, k)]; }
#pragma unroll
  for (int k = 0; k < N; k++) lrhs[0][k] = s[(w++) * n + i];
  lim_lo[0] = (uint32_t)s[(w++) * n + i]; lim_hi[0] = (uint32_t)s[(w++) * n + i];
  grhs[0] = s[(w++) * n + i]; gdinv[0] = s[(w++) * n + i]; gapp[0] = 0.f;
  uint64_t bad = 0;
  for (int it = 0; it < sweeps; it++) {
    bool resid_bad = false;
    if (it & 1) {
      ARM_LIMITS_FWD_GRIPPER(0) ARM_MOTORS_FWD(0) GEAR_ROW(0)
    } else {
      GEAR_ROW(0) ARM_MOTORS_BWD(0) ARM_LIMITS_BWD_GRIPPER(0)
    }
    bad |= (uint64_t)(resid_bad ? 1u : 0u) << it;
  }
#pragma unroll
  for (int k = 0; k < N; k++) dqd_out[k] = dqd[0][k];
  return bad;
}
// the door's two rows (default velocity motor, violated limit); words dl_rhs dl_sign dm_rhs door_lim at s
XD uint64_t light_island_door(const float* __restrict__ s, int64_t n, int64_t i, int sweeps, float& ddoor_out) {
  const float dl_rhs = s[0 * n + i], dl_sign = s[1 * n + i], dm_rhs = s[2 * n + i];
  const bool door_lim = (int)s[3 * n + i] != 0;
  const float door_den = 1.f / (float)XARM_DOOR_MASS, hi_lim = (float)XARM_LIMIT_MAX_IMPULSE;
  const float sthr_ = sqrtf((float)XARM_RESIDUAL_THRESHOLD);
  float ddoor = 0.f, dl_app = 0.f, dm_app = 0.f;
  uint64_t bad = 0;
  for (int it = 0; it < sweeps; it++) {
    bool resid_bad = false;
    if (it & 1) { DOOR_LIMIT_ROW() DOOR_MOTOR_ROW() } else { DOOR_MOTOR_ROW() DOOR_LIMIT_ROW() }
    bad |= (uint64_t)(resid_bad ? 1u : 0u) << it;
  }
  ddoor_out = ddoor;
  return bad;
}
// one object's manifold against a static box (the manifold part of sub_solve_light)
template <class T>
XD uint64_t light_island_manifold(const ManifoldIn& MI, int nc, float* mrows, int stride, int sweeps, V3& v_out, V3& w_out) {
  const float inv_obj_mass = 1.f / T::OBJ_MASS;
  const float sthr_ = sqrtf((float)XARM_RESIDUAL_THRESHOLD);
  manifold_rows_store<T>(MI, nc, mrows, stride);
  const V3 n = MI.n;
  const float mu = MI.mu;
  V3 t1, t2;
  plane_space(n, t1, t2);
  const V3 nm = inv_obj_mass * n, t1m = inv_obj_mass * t1, t2m = inv_obj_mass * t2;
  float an[4], a1[4], a2[4];
#pragma unroll
  for (int c = 0; c < 4; c++) { an[c] = 0.f; a1[c] = 0.f; a2[c] = 0.f; }
  V3 v = v3(0, 0, 0), w = v3(0, 0, 0);
  uint64_t bad = 0;
#define MR(c, k) mrows[(size_t)(24 * (c) + (k)) * stride]
#define MRV(c, k) v3(MR(c, k), MR(c, (k) + 1), MR(c, (k) + 2))
  for (int it = 0; it < sweeps; it++) {
    bool resid_bad = false;
#pragma unroll
    for (int c = 0; c < 4; c++) {  // normal rows
      const float dn = MR(c, 19);
      float delta = MR(c, 18) - (dot(n, v) + dot(MRV(c, 0), w)) * dn;
      const float sum = an[c] + delta;
      const float sumc = fminf(fmaxf(sum, 0.f), (float)XARM_CONTACT_MAX_IMPULSE);
      delta = (sumc == sum) ? delta : sumc - an[c];
      an[c] = sumc;
      v += delta * nm; w += delta * MRV(c, 3);
      resid_bad = resid_bad || fabsf(delta) > sthr_ * dn;
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {  // friction pairs, implicit cone
      const float lim = mu * an[c];
      const float d1 = MR(c, 21), d2 = MR(c, 23);
      float da = MR(c, 20) - (dot(t1, v) + dot(MRV(c, 6), w)) * d1, db = MR(c, 22) - (dot(t2, v) + dot(MRV(c, 12), w)) * d2;
      float sa = a1[c] + da, sb = a2[c] + db;
      const float l2 = sa * sa + sb * sb;
      if (l2 > lim * lim) {
        const float len = sqrtf(l2);
        if (len > lim) { const float sc = lim / len; sa *= sc; sb *= sc; da = sa - a1[c]; db = sb - a2[c]; }
      }
      a1[c] = sa; a2[c] = sb;
      v += da * t1m; w += da * MRV(c, 9);
      v += db * t2m; w += db * MRV(c, 15);
      resid_bad = resid_bad || fabsf(da) > sthr_ * d1 || fabsf(db) > sthr_ * d2;
    }
    bad |= (uint64_t)(resid_bad ? 1u : 0u) << it;
  }
#undef MR
#undef MRV
  v_out = v; w_out = w;
  return bad;
}

// ---- scratch records of the rows (the pipeline's setup kernel writes them, its light kernel reads them; word w of env i at
// base[w * n + i]).  The fused per-env substep() uses the same records with n = 1.
template <class T>
XD void ar_store(const ArmRows<T>& AR, float* __restrict__ s, int64_t n, int64_t i) {
  constexpr int N = T::MD::N, NT = N * (N + 1) / 2;
  int w = 0;
#pragma unroll
  for (int a = 0; a < T::NARM; a++) {
#pragma unroll
    for (int k = 0; k < NT; k++) s[(w++) * n + i] = AR.Mi[a][k];
#pragma unroll
    for (int k = 0; k < N; k++) s[(w++) * n + i] = AR.mrhs[a][k];
#pragma unroll
    for (int k = 0; k < N; k++) s[(w++) * n + i] = AR.lrhs[a][k];
    s[(w++) * n + i] = (float)AR.lim_lo[a]; s[(w++) * n + i] = (float)AR.lim_hi[a];  // bit masks < 2^13: exact
    s[(w++) * n + i] = AR.grhs[a]; s[(w++) * n + i] = AR.gdinv[a]; s[(w++) * n + i] = AR.gden[a];
  }
  s[(w++) * n + i] = AR.dl_rhs; s[(w++) * n + i] = AR.dl_sign; s[(w++) * n + i] = AR.dm_rhs;
  s[(w++) * n + i] = (float)AR.door_lim;
}
template <class T>
XD void ar_load(ArmRows<T>& AR, const float* __restrict__ s, int64_t n, int64_t i) {
  constexpr int N = T::MD::N, NT = N * (N + 1) / 2;
  int w = 0;
#pragma unroll
  for (int a = 0; a < T::NARM; a++) {
#pragma unroll
    for (int k = 0; k < NT; k++) AR.Mi[a][k] = s[(w++) * n + i];
#pragma unroll
    for (int k = 0; k < N; k++) AR.mrhs[a][k] = s[(w++) * n + i];
#pragma unroll
    for (int k = 0; k < N; k++) AR.lrhs[a][k] = s[(w++) * n + i];
    AR.lim_lo[a] = (uint32_t)s[(w++) * n + i]; AR.lim_hi[a] = (uint32_t)s[(w++) * n + i];
    AR.grhs[a] = s[(w++) * n + i]; AR.gdinv[a] = s[(w++) * n + i]; AR.gden[a] = s[(w++) * n + i];
  }
  AR.dl_rhs = s[(w++) * n + i]; AR.dl_sign = s[(w++) * n + i]; AR.dm_rhs = s[(w++) * n + i];
  AR.door_lim = (int)s[(w++) * n + i];
}
XD void mi_store(const ManifoldIn& M, float* __restrict__ s, int64_t n, int64_t i) {
  int w = 0;
  s[(w++) * n + i] = M.n.x; s[(w++) * n + i] = M.n.y; s[(w++) * n + i] = M.n.z; s[(w++) * n + i] = M.mu;
#pragma unroll
  for (int c = 0; c < 4; c++) { s[(w++) * n + i] = M.r[c].x; s[(w++) * n + i] = M.r[c].y; s[(w++) * n + i] = M.r[c].z; }
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int k = 0; k < 3; k++) s[(w++) * n + i] = M.rhs[c][k];
  s[(w++) * n + i] = M.Iinv.xx; s[(w++) * n + i] = M.Iinv.xy; s[(w++) * n + i] = M.Iinv.xz;
  s[(w++) * n + i] = M.Iinv.yy; s[(w++) * n + i] = M.Iinv.yz; s[(w++) * n + i] = M.Iinv.zz;
}
XD void mi_load(ManifoldIn& M, const float* __restrict__ s, int64_t n, int64_t i) {
  int w = 0;
  M.n.x = s[(w++) * n + i]; M.n.y = s[(w++) * n + i]; M.n.z = s[(w++) * n + i]; M.mu = s[(w++) * n + i];
#pragma unroll
  for (int c = 0; c < 4; c++) { M.r[c].x = s[(w++) * n + i]; M.r[c].y = s[(w++) * n + i]; M.r[c].z = s[(w++) * n + i]; }
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int k = 0; k < 3; k++) M.rhs[c][k] = s[(w++) * n + i];
  M.Iinv.xx = s[(w++) * n + i]; M.Iinv.xy = s[(w++) * n + i]; M.Iinv.xz = s[(w++) * n + i];
  M.Iinv.yy = s[(w++) * n + i]; M.Iinv.yz = s[(w++) * n + i]; M.Iinv.zz = s[(w++) * n + i];
}

XHD int mi_words() { return 3 + 1 + 12 + 12 + 6; }
template <class T>
constexpr bool task_single_island_pair() { return T::NARM == 1 && !T::HAS_DOOR && T::NOBJ <= 1; }   // Reach, PickAndPlace with one lego: sub_setup_lean + sub_solve_light

// the multi-island light solve: rows from the scratch records (ar | door words at the end of ar | mi[o] at s_mi + o * mi_words),
// nc[o] manifold points of object o.  mrows: room for ONE manifold's rows (reused object after object).
template <class T>
XD void sub_solve_light_multi(const float* __restrict__ s_ar, const float* __restrict__ s_mi, int64_t n, int64_t i, const int* nc,
                              float* mrows, int stride, SubSol<T>& S) {
  constexpr int NA = T::NARM, NOBJ = T::NOBJ, NO = NOBJ > 0 ? NOBJ : 1;
  static_assert(XARM_SOLVER_ITERATIONS <= 63, "one bit per sweep");
  int sweeps = XARM_SOLVER_ITERATIONS;
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {
    uint64_t bad = 0;
#pragma unroll 1
    for (int a = 0; a < NA; a++) bad |= light_island_arm<T>(s_ar + (int64_t)(a * ar_arm_words<T>()) * n, n, i, sweeps, S.dqd[a]);
    S.ddoor = 0.f;
    if (T::HAS_DOOR) bad |= light_island_door(s_ar + (int64_t)(NA * ar_arm_words<T>()) * n, n, i, sweeps, S.ddoor);
#pragma unroll 1
    for (int o = 0; o < NO; o++) {
      S.dv[o] = v3(0, 0, 0); S.dw[o] = v3(0, 0, 0);
      if (o < NOBJ && nc[o] > 0) {
        ManifoldIn MI;
        mi_load(MI, s_mi + (int64_t)(o * mi_words()) * n, n, i);
        bad |= light_island_manifold<T>(MI, nc[o], mrows, stride, sweeps, S.dv[o], S.dw[o]);
      }
    }
    // the joint loop leaves after the first sweep in which no row of any island moved more than the threshold
    const uint64_t good = ~bad & ((1ull << sweeps) - 1ull);
    if (good == 0ull) break;
    int first = 0;
    while (!((good >> first) & 1ull)) first++;
    if (first + 1 >= sweeps) break;
    sweeps = first + 1;
  }
}

// Generic form: the joint loop of btMultiBodyConstraintSolver::solveSingleIteration over ALL rows of the env -
// non-contact rows of every arm and the door (forward on odd sweeps, backward on even ones), the normal rows of every
// contact, then its friction pairs (implicit cone) - until no row moved more than the residual threshold.  Contact rows
// are read from C (shared memory in the heavy kernel); the arm rows live in registers.
template <class T>
XD float arm_dot(const float* J, const float* dqd) {  // three partial sums: shorter dependency chain
  constexpr int N = T::MD::N;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < N; i += 3) {
    s0 += J[i] * dqd[i];
    if (i + 1 < N) s1 += J[i + 1] * dqd[i + 1];
    if (i + 2 < N) s2 += J[i + 2] * dqd[i + 2];
  }
  return (s0 + s1) + s2;
}
// An island description restricts the loop to a subset of the env's rows (substep_generic_islands below): the arms in
// `arm_mask`, the door when `door_on`, the contacts idx[0..n_idx) (ascending: the joint loop's order); it then runs exactly
// `sweeps` sweeps and returns a bit per sweep ("a row moved more than the threshold") instead of leaving by itself, and
// writes only the velocity changes of the bodies it swept.
struct GenericIsland {
  uint32_t arm_mask;
  bool door_on;
  const uint8_t* idx;
  int n_idx;
  int sweeps;
  uint32_t obj_mask;
};
template <class T>
XD uint64_t sub_solve_generic(const ArmRows<T>& AR, Contacts<T>& C, SubSol<T>& S, int /*form*/, const GenericIsland* isl = nullptr) {
  using MD = typename T::MD;
  constexpr int N = MD::N, NA = T::NARM, NOBJ = T::NOBJ, NO = NOBJ > 0 ? NOBJ : 1, NT = N * (N + 1) / 2;
  SOLVER_LOCALS_FROM(AR)
  V3 dv[NO], dw[NO];
#pragma unroll
  for (int o = 0; o < NO; o++) { dv[o] = v3(0, 0, 0); dw[o] = v3(0, 0, 0); }
  const int nc = isl ? isl->n_idx : C.nc;
  const uint32_t arm_mask = isl ? isl->arm_mask : 0xffffffffu;
  const bool door_on = isl ? isl->door_on : true;
  const int max_sweeps = isl ? isl->sweeps : XARM_SOLVER_ITERATIONS;
  uint64_t bad_bits = 0;
  for (int it = 0; it < max_sweeps; it++) {
    bool resid_bad = false;
    if (it & 1) {
#pragma unroll
      for (int a = 0; a < NA; a++) {
        if (arm_mask >> a & 1u) { ARM_LIMITS_FWD(a) ARM_MOTORS_FWD(a) GEAR_ROW(a) }
      }
      if (T::HAS_DOOR && door_on) { DOOR_LIMIT_ROW() DOOR_MOTOR_ROW() }
    } else {
      if (T::HAS_DOOR && door_on) { DOOR_MOTOR_ROW() DOOR_LIMIT_ROW() }
#pragma unroll
      for (int a = NA - 1; a >= 0; a--) {
        if (arm_mask >> a & 1u) { GEAR_ROW(a) ARM_MOTORS_BWD(a) ARM_LIMITS_BWD(a) }
      }
    }
    // ---- normal rows
    for (int ci = 0; ci < nc; ci++) {
      const int c = isl ? (int)isl->idx[ci] : ci;
      const int o1 = NOBJ <= 1 ? (C.o1[c] >= 0 ? 0 : -1) : C.o1[c];
      const int o2 = NOBJ > 1 ? C.o2[c] : -1;
      const int sl = C.slot[c];
      const float s1 = C.s1[c];
      const int arm_of = (NA > 1 && sl >= 0) ? bc_arm(bc_is_arm(C.ba[c]) ? C.ba[c] : C.bb[c]) : 0;
      const V3 d = C.dir[c][0];
      float s_ = 0.f;
      if (o1 >= 0) s_ += s1 * dot(d, dv[NOBJ <= 1 ? 0 : o1]) + dot(C.Jo1[c][0], dw[NOBJ <= 1 ? 0 : o1]);
      if (NOBJ > 1 && o2 >= 0) s_ += -dot(d, dv[o2 < 0 ? 0 : o2]) + dot(C.Jo2[NOBJ > 1 ? c : 0][0], dw[o2 < 0 ? 0 : o2]);
      if (sl >= 0) {
        if (NA == 1 || arm_of == 0) s_ += arm_dot<T>(C.Jarm[sl][0], dqd[0]);
        else s_ += arm_dot<T>(C.Jarm[sl][0], dqd[NA - 1]);
      }
      if (T::HAS_DOOR) s_ += C.jdoor[T::HAS_DOOR ? c : 0][0] * ddoor;
      const float app0 = C.app[c][0], dinv0 = C.dinv[c][0];
      float d0 = C.rhs[c][0] - app0 * C.cfmr[c] - s_ * dinv0;
      float sum = app0 + d0;
      if (sum < 0.f) { d0 = -app0; sum = 0.f; }
      else if (sum > (float)XARM_CONTACT_MAX_IMPULSE) { d0 = (float)XARM_CONTACT_MAX_IMPULSE - app0; sum = (float)XARM_CONTACT_MAX_IMPULSE; }
      C.app[c][0] = sum;
      resid_bad = resid_bad || fabsf(d0) > sthr_ * dinv0;
      if (o1 >= 0) { dv[NOBJ <= 1 ? 0 : o1] += (s1 * d0 * inv_obj_mass) * d; dw[NOBJ <= 1 ? 0 : o1] += d0 * C.dVo1[c][0]; }
      if (NOBJ > 1 && o2 >= 0) { dv[o2 < 0 ? 0 : o2] += (-d0 * inv_obj_mass) * d; dw[o2 < 0 ? 0 : o2] += d0 * C.dVo2[NOBJ > 1 ? c : 0][0]; }
      if (sl >= 0) {
        if (NA == 1 || arm_of == 0) { _Pragma("unroll") for (int i = 0; i < N; i++) dqd[0][i] += C.dVarm[sl][0][i] * d0; }
        else { _Pragma("unroll") for (int i = 0; i < N; i++) dqd[NA - 1][i] += C.dVarm[sl][0][i] * d0; }
      }
      if (T::HAS_DOOR) ddoor += C.jdoor[T::HAS_DOOR ? c : 0][0] * d0 / (float)XARM_DOOR_MASS;
    }
    // ---- friction pairs (implicit cone: the pair is scaled back onto mu * normal impulse)
    for (int ci = 0; ci < nc; ci++) {
      const int c = isl ? (int)isl->idx[ci] : ci;
      const int o1 = NOBJ <= 1 ? (C.o1[c] >= 0 ? 0 : -1) : C.o1[c];
      const int o2 = NOBJ > 1 ? C.o2[c] : -1;
      const int sl = C.slot[c];
      const float s1 = C.s1[c];
      const int arm_of = (NA > 1 && sl >= 0) ? bc_arm(bc_is_arm(C.ba[c]) ? C.ba[c] : C.bb[c]) : 0;
      const V3 t1 = C.dir[c][1], t2 = C.dir[c][2];
      float ja = 0.f, jb = 0.f;
      if (o1 >= 0) {
        const V3 v_ = dv[NOBJ <= 1 ? 0 : o1], w_ = dw[NOBJ <= 1 ? 0 : o1];
        ja += s1 * dot(t1, v_) + dot(C.Jo1[c][1], w_);
        jb += s1 * dot(t2, v_) + dot(C.Jo1[c][2], w_);
      }
      if (NOBJ > 1 && o2 >= 0) {
        const V3 v_ = dv[o2 < 0 ? 0 : o2], w_ = dw[o2 < 0 ? 0 : o2];
        ja += -dot(t1, v_) + dot(C.Jo2[NOBJ > 1 ? c : 0][1], w_);
        jb += -dot(t2, v_) + dot(C.Jo2[NOBJ > 1 ? c : 0][2], w_);
      }
      if (sl >= 0) {
        if (NA == 1 || arm_of == 0) { ja += arm_dot<T>(C.Jarm[sl][1], dqd[0]); jb += arm_dot<T>(C.Jarm[sl][2], dqd[0]); }
        else { ja += arm_dot<T>(C.Jarm[sl][1], dqd[NA - 1]); jb += arm_dot<T>(C.Jarm[sl][2], dqd[NA - 1]); }
      }
      if (T::HAS_DOOR) { ja += C.jdoor[T::HAS_DOOR ? c : 0][1] * ddoor; jb += C.jdoor[T::HAS_DOOR ? c : 0][2] * ddoor; }
      const float lim = C.mu[c] * C.app[c][0];
      const float app1 = C.app[c][1], app2 = C.app[c][2], di1 = C.dinv[c][1], di2 = C.dinv[c][2];
      float da = C.rhs[c][1] - ja * di1, db = C.rhs[c][2] - jb * di2;
      float sa = app1 + da, sb = app2 + db;
      const float len = sqrtf(sa * sa + sb * sb);
      if (len > lim) { const float sc = len > 0.f ? lim / len : 0.f; sa *= sc; sb *= sc; da = sa - app1; db = sb - app2; }
      C.app[c][1] = sa; C.app[c][2] = sb;
      resid_bad = resid_bad || fabsf(da) > sthr_ * di1 || fabsf(db) > sthr_ * di2;
      if (o1 >= 0) {
        dv[NOBJ <= 1 ? 0 : o1] += (s1 * da * inv_obj_mass) * t1; dw[NOBJ <= 1 ? 0 : o1] += da * C.dVo1[c][1];
        dv[NOBJ <= 1 ? 0 : o1] += (s1 * db * inv_obj_mass) * t2; dw[NOBJ <= 1 ? 0 : o1] += db * C.dVo1[c][2];
      }
      if (NOBJ > 1 && o2 >= 0) {
        dv[o2 < 0 ? 0 : o2] += (-da * inv_obj_mass) * t1; dw[o2 < 0 ? 0 : o2] += da * C.dVo2[NOBJ > 1 ? c : 0][1];
        dv[o2 < 0 ? 0 : o2] += (-db * inv_obj_mass) * t2; dw[o2 < 0 ? 0 : o2] += db * C.dVo2[NOBJ > 1 ? c : 0][2];
      }
      if (sl >= 0) {
        if (NA == 1 || arm_of == 0) { _Pragma("unroll") for (int i = 0; i < N; i++) dqd[0][i] += C.dVarm[sl][1][i] * da + C.dVarm[sl][2][i] * db; }
        else { _Pragma("unroll") for (int i = 0; i < N; i++) dqd[NA - 1][i] += C.dVarm[sl][1][i] * da + C.dVarm[sl][2][i] * db; }
      }
      if (T::HAS_DOOR) ddoor += (C.jdoor[T::HAS_DOOR ? c : 0][1] * da + C.jdoor[T::HAS_DOOR ? c : 0][2] * db) / (float)XARM_DOOR_MASS;
    }
    if (isl) bad_bits |= (uint64_t)(resid_bad ? 1u : 0u) << it;
    else if (!resid_bad) break;
  }
#pragma unroll
  for (int a = 0; a < NA; a++)
    if (arm_mask >> a & 1u) {
#pragma unroll
      for (int i = 0; i < N; i++) S.dqd[a][i] = dqd[a][i];
    }
#pragma unroll
  for (int o = 0; o < NO; o++)
    if (!isl || (isl->obj_mask >> o & 1u)) { S.dv[o] = dv[o]; S.dw[o] = dw[o]; }
  if (door_on) S.ddoor = ddoor;
  return bad_bits;
}

// ---- Generic substep of the multi-island tasks with the simple islands taken out of the joint loop.  In a heavy env of those
// tasks most rows still belong to islands the light forms can sweep: an arm that touches nothing (registers, light_island_arm),
// an object whose only points come from ONE pair with a static box (light_island_manifold), the door when nothing touches it.
// The generic loop - ~750 cycles per row for a lone thread: dynamic body indices, rows streamed from the record - keeps only
// the rows of the coupled remainder.  Islands are independent, so every row gets the impulses of the joint loop (the manifold
// form rounds differently from the generic contact row: heavy envs are compared statistically); the loop's exit test is
// rebuilt from the islands' per-sweep bits as in sub_solve_light_multi.
template <class T>
XD void solve_generic_islands(const Env<T>& e, const ArmRows<T>& AR, Contacts<T>& C, const S3* Iinv, SubSol<T>& S) {
  using MD = typename T::MD;
  constexpr int NA = T::NARM, NOBJ = T::NOBJ, NO = NOBJ > 0 ? NOBJ : 1;
#ifdef XARM_HOST_SIM
  if (getenv("XARM_NO_ISLANDS")) { sub_solve_generic<T>(AR, C, S, SOLVE_GENERIC_JOINT); return; }   // test hook: the plain joint loop
#endif
  // ---- which bodies the coupled remainder holds
  uint32_t arm_in = 0u, obj_gen = 0u;
  bool door_in = false;
  int first[NO], cnt[NO];
  bool one_pair[NO];
#pragma unroll
  for (int o = 0; o < NO; o++) { first[o] = -1; cnt[o] = 0; one_pair[o] = true; }
  for (int c = 0; c < C.nc; c++) {
    const int ca = C.ba[c], cb = C.bb[c];
    if (C.slot[c] >= 0) arm_in |= 1u << (NA == 1 ? 0 : bc_arm(bc_is_arm(ca) ? ca : cb));
    if (T::HAS_DOOR && (ca == BC_DOOR || cb == BC_DOOR)) door_in = true;
    const int o1 = C.o1[c], o2 = NOBJ > 1 ? C.o2[c] : -1;
    const bool simple = bc_is_obj(ca) && cb == BC_STATIC && C.cfm0[c] == 0.f;   // object (side A) on a static box
    if (!simple) {
      if (o1 >= 0) obj_gen |= 1u << (o1 & 3);
      if (NOBJ > 1 && o2 >= 0) obj_gen |= 1u << (o2 & 3);
    } else {
      const int o = o1 < 0 ? 0 : o1;
      if (first[o] < 0) first[o] = c; else if (C.pair[c] != C.pair[first[o]]) one_pair[o] = false;
      cnt[o]++;
    }
  }
#pragma unroll
  for (int o = 0; o < NO; o++) if (!one_pair[o] || cnt[o] > 4) obj_gen |= 1u << o;
  for (int a = 0; a < NA; a++) if (((AR.lim_lo[a] | AR.lim_hi[a]) & 0x7fu) != 0u) arm_in |= 1u << a;   // arm-joint limit rows: generic form only
  uint8_t idx[T::MAXC];
  int n_idx = 0;
  for (int c = 0; c < C.nc; c++) {
    const int o1 = C.o1[c], o2 = NOBJ > 1 ? C.o2[c] : -1;
    const bool in_gen = C.slot[c] >= 0 || (o1 >= 0 && (obj_gen >> (o1 & 3) & 1u)) || (NOBJ > 1 && o2 >= 0 && (obj_gen >> (o2 & 3) & 1u)) || (o1 < 0 && o2 < 0);
    if (in_gen) idx[n_idx++] = (uint8_t)c;
  }
  float rec_ar[T::NARM * (T::MD::N * (T::MD::N + 1) / 2 + 2 * T::MD::N + 5) + 4];
  ar_store<T>(AR, rec_ar, 1, 0);
  int sweeps = XARM_SOLVER_ITERATIONS;
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {
    uint64_t bad = 0;
#pragma unroll 1
    for (int a = 0; a < NA; a++)
      if (!(arm_in >> a & 1u)) bad |= light_island_arm<T>(rec_ar + a * ar_arm_words<T>(), 1, 0, sweeps, S.dqd[a]);
    S.ddoor = 0.f;
    if (T::HAS_DOOR && !door_in) bad |= light_island_door(rec_ar + NA * ar_arm_words<T>(), 1, 0, sweeps, S.ddoor);
#pragma unroll 1
    for (int o = 0; o < NO; o++) {
      if (obj_gen >> o & 1u) continue;
      S.dv[o] = v3(0, 0, 0); S.dw[o] = v3(0, 0, 0);
      if (o < NOBJ && cnt[o] > 0) {
        ManifoldIn MI;
        const int c0 = first[o];
        MI.n = C.dir[c0][0]; MI.mu = C.mu[c0]; MI.Iinv = Iinv[o];
        int k = 0;
        for (int c = c0; c < C.nc && k < 4; c++) {
          if (C.o1[c] != (NOBJ <= 1 ? C.o1[c0] : o) || C.pair[c] != C.pair[c0]) continue;
          MI.r[k] = C.pa[c] - e.obj[o].pos;
          MI.rhs[k][0] = C.rhs[c][0]; MI.rhs[k][1] = C.rhs[c][1]; MI.rhs[k][2] = C.rhs[c][2];
          k++;
        }
        for (; k < 4; k++) { MI.r[k] = v3(0, 0, 0); MI.rhs[k][0] = 0.f; MI.rhs[k][1] = 0.f; MI.rhs[k][2] = 0.f; }
        float mrows[XARM_MROW_WORDS];
        bad |= light_island_manifold<T>(MI, cnt[o], mrows, 1, sweeps, S.dv[o], S.dw[o]);
      }
    }
    if (n_idx > 0 || arm_in != 0u || (T::HAS_DOOR && door_in)) {
      for (int k = 0; k < n_idx; k++) { const int c = idx[k]; C.app[c][0] = 0.f; C.app[c][1] = 0.f; C.app[c][2] = 0.f; }
      GenericIsland isl = {arm_in, T::HAS_DOOR && door_in, idx, n_idx, sweeps, obj_gen};
      bad |= sub_solve_generic<T>(AR, C, S, SOLVE_GENERIC_JOINT, &isl);
    }
    const uint64_t good = ~bad & ((1ull << sweeps) - 1ull);
    if (good == 0ull) break;
    int firstg = 0;
    while (!((good >> firstg) & 1ull)) firstg++;
    if (firstg + 1 >= sweeps) break;
    sweeps = firstg + 1;
  }
}

// stepPositionsMultiDof: add the solved velocity changes, integrate joints and boxes
template <class T>
XD void sub_integrate(Env<T>& e, const SubBase<T>& B, const SubSol<T>& S) {
  using MD = typename T::MD;
  constexpr int N = MD::N, NA = T::NARM, NOBJ = T::NOBJ;
  const float h = (float)T::H;
#pragma unroll
  for (int a = 0; a < NA; a++)
#pragma unroll
    for (int i = 0; i < N; i++) {
      float qd = B.qdu[a][i] + S.dqd[a][i];
      e.arm[a].qd[i] = qd;
      e.arm[a].q[i] += qd * h;
    }
  for (int o = 0; o < NOBJ; o++) {
    ObjState& b = e.obj[o];
    b.v = B.vu[o] + S.dv[o]; b.w = B.wu[o] + S.dw[o];
    b.pos += h * b.v;
    float ang = norm(b.w);
    if (ang * h > (float)XARM_ANGULAR_MOTION_THRESHOLD) ang = (float)XARM_ANGULAR_MOTION_THRESHOLD / h;
    V3 ax;
    if (ang < 0.001f) ax = (0.5f * h - h * h * h * 0.020833333333f * ang * ang) * b.w;
    else ax = (sinf(0.5f * ang * h) / ang) * b.w;
    Q4 dq = {ax.x, ax.y, ax.z, cosf(ang * h * 0.5f)};
    Q4 qn = quat_mul(dq, b.quat);
    float inv = rsqrtf(qn.x * qn.x + qn.y * qn.y + qn.z * qn.z + qn.w * qn.w);
    b.quat.x = qn.x * inv; b.quat.y = qn.y * inv; b.quat.z = qn.z * inv; b.quat.w = qn.w * inv;
  }
  if (T::HAS_DOOR) { e.door_qd = B.door_qdu + S.ddoor; e.door_q += e.door_qd * h; }
#ifdef XARM_HOST_SIM
  if (getenv("XARM_TRACE")) {
    fprintf(stderr, "KS qd");
    for (int a = 0; a < NA; a++) for (int i = 0; i < N; i++) fprintf(stderr, " %.10f", (double)e.arm[a].qd[i]);
    for (int o = 0; o < NOBJ; o++) fprintf(stderr, " %.10f %.10f %.10f %.10f %.10f %.10f", (double)e.obj[o].v.x, (double)e.obj[o].v.y, (double)e.obj[o].v.z, (double)e.obj[o].w.x, (double)e.obj[o].w.y, (double)e.obj[o].w.z);
    fprintf(stderr, "\n");
  }
#endif
}

// One internal substep, fused: collide -> unconstrained velocities -> rows -> PGS -> integrate.
template <class T>
constexpr bool task_has_light() { return true; }   // every task has a light solver form (single island pair or multi-island)
template <class T>
XHD int ar_words() { return T::NARM * ar_arm_words<T>() + 4; }   // ar_store record: per arm, then dl_rhs dl_sign dm_rhs door_lim
template <class T>
XD bool arm_joint_on_limit(const ArmRows<T>& AR) {   // the light forms keep limit rows for the gripper dofs only
  uint32_t m = 0u;
#pragma unroll
  for (int a = 0; a < T::NARM; a++) m |= AR.lim_lo[a] | AR.lim_hi[a];
  return (m & 0x7fu) != 0u;
}
template <class T>
NOINL void substep(Env<T>& e, bool apply_damping, bool last) {
  ArmRows<T> AR;
  SubBase<T> B;
  SubSol<T> S;
  ManifoldIn MI;
  if constexpr (task_single_island_pair<T>()) {
    int nc = 0;
    ArmDyn<typename T::MD> D[1];
    if (sub_setup_lean<T>(e, apply_damping, last, AR, B, MI, nc, D) && !arm_joint_on_limit<T>(AR)) {  // (arm-joint limits: generic form)
      float mrows[XARM_MROW_WORDS];
      sub_solve_light<T>(AR, nc, MI, mrows, 1, S);
      sub_integrate<T>(e, B, S);
      return;
    }
  } else {
    constexpr int NO = T::NOBJ > 0 ? T::NOBJ : 1;
    ManifoldIn MIs[NO];
    int nc[NO];
    ArmDyn<typename T::MD> D[T::NARM];
    int g0[T::NARM];
    for (int a = 0; a < T::NARM; a++) g0[a] = e.grasp[a];
#ifdef XARM_HOST_SIM
    const bool no_lean_ = getenv("XARM_NO_LEAN") != nullptr;   // test hook: every env through the generic setup + solve
#else
    const bool no_lean_ = false;
#endif
    if (!no_lean_ && sub_setup_lean_multi<T>(e, apply_damping, last, AR, B, MIs, nc, D) && !arm_joint_on_limit<T>(AR)) {
      float rec_ar[T::NARM * (T::MD::N * (T::MD::N + 1) / 2 + 2 * T::MD::N + 5) + 4], rec_mi[NO * 34];
      static_assert(sizeof(rec_mi) / sizeof(float) / NO == 34, "ManifoldIn record");
      ar_store<T>(AR, rec_ar, 1, 0);
      for (int o = 0; o < T::NOBJ; o++) if (nc[o] > 0) mi_store(MIs[o], rec_mi + o * mi_words(), 1, 0);
      float mrows[XARM_MROW_WORDS];
      sub_solve_light_multi<T>(rec_ar, rec_mi, 1, 0, nc, mrows, 1, S);
      sub_integrate<T>(e, B, S);
      return;
    }
    for (int a = 0; a < T::NARM; a++) e.grasp[a] = g0[a];   // (the lean setup may have cleared the flags before it found a contact)
  }
  Contacts<T> C;
  if constexpr (task_single_island_pair<T>()) {
    const int form = sub_setup<T>(e, apply_damping, last, AR, C, B, MI);
    sub_solve_generic<T>(AR, C, S, form == SOLVE_GENERIC_JOINT ? SOLVE_GENERIC_JOINT : SOLVE_GENERIC_DECOUPLED);
  } else {
    S3 Iinv[T::NOBJ > 0 ? T::NOBJ : 1];
    sub_setup<T>(e, apply_damping, last, AR, C, B, MI, nullptr, Iinv);
    solve_generic_islands<T>(e, AR, C, Iinv, S);
  }
  sub_integrate<T>(e, B, S);
}
// the same substep without the light solver (pipeline: envs the setup kernel classified as not light).  The contact
// rows live where the caller puts them: shared memory in the heavy kernel, thread-local memory elsewhere.
template <class T>
XD void substep_generic(Env<T>& e, bool apply_damping, bool last, Contacts<T>& C) {
  ArmRows<T> AR;
  SubBase<T> B;
  SubSol<T> S;
  ManifoldIn MI;
#if defined(XARM_REC_PROF) && defined(__CUDA_ARCH__)
  const long long t0_ = clock64();
#endif
  S3 Iinv_[T::NOBJ > 0 ? T::NOBJ : 1];
  const int form = sub_setup<T>(e, apply_damping, last, AR, C, B, MI, nullptr, Iinv_);
#if defined(XARM_REC_PROF) && defined(__CUDA_ARCH__)
  const long long t1_ = clock64();
#endif
  if constexpr (task_single_island_pair<T>()) sub_solve_generic<T>(AR, C, S, form == SOLVE_GENERIC_JOINT ? SOLVE_GENERIC_JOINT : SOLVE_GENERIC_DECOUPLED);
  else solve_generic_islands<T>(e, AR, C, Iinv_, S);
#if defined(XARM_REC_PROF) && defined(__CUDA_ARCH__)
  const long long t2_ = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) printf("[rec prof] nc %d nac %d | setup %lld | solve %lld cycles\n", C.nc, C.nac, t1_ - t0_, t2_ - t1_);
#endif
  sub_integrate<T>(e, B, S);
}

// p.stepSimulation() x calls per env step
template <class T>
XD void simulate(Env<T>& e) {
  for (int s = 0; s < T::NSUB; s++) substep<T>(e, T::DAMP_EACH || s == 0, s == T::NSUB - 1);
}
