// xarm_heavy.cuh - the coupled PGS of a gripper-contact env (one arm, one object, no door: PickAndPlace) with every
// contact row in a solver-friendly record in SHARED memory.
//
// The generic solver (sub_solve_generic) reads its rows from the Contacts record field by field: ~115 instructions and
// 60 scalar shared-memory loads per arm-coupled row, a single warp per SM (the record is 5.8 KB per env) at 0.23 IPC -
// 0.7 ms per substep for one warp, and every step waits for 90 of those in the auto-reset tail (profiles/).  Here a
// row is 8 float4: [J(9 joint, 3 linear, 3 angular) rhs | M^-1 J^T (15) 1/denominator], float4 u of lane l at
// rows4[u * 32 + l] (conflict-free LDS.128), so one row is 8 loads, 15 + 15 FMAs and the clamp.  Row order, clamps and
// exit test are those of btMultiBodyConstraintSolver::solveSingleIteration, as in sub_solve_generic.
#pragma once
#include "xarm_pipeline.cuh"

// Record of one heavy env, written by k_heavy_rows (one thread per env: generic setup) and consumed by k_heavy_solve
// (16 lanes per env).  Global memory, HeavyRec<T>::WORDS floats per heavy-list slot.
//   header: arm rows (ArmRows: Mi 45, mrhs 9, lrhs 9, limit masks 2, gear 3) | unconstrained velocities (qdu 9, vu 3, wu 3) | nc
//   rows  : (c * 3 + k) * ROW: J[0..14] pad | V[0..14] pad | rhs dinv cfmr mu        J = [9 joint | 3 linear | 3 angular]
template <class T>
struct HeavyRec {
  static constexpr int N = T::MD::N, NT = N * (N + 1) / 2, MAXC = T::MAXC;
  static constexpr int MI = 0, MRHS = MI + NT, LRHS = MRHS + N, LIM = LRHS + N, GEAR = LIM + 2, QDU = GEAR + 3, VU = QDU + N, WU = VU + 3, NC = WU + 3;
  static constexpr int HDR = (NC + 1 + 3) / 4 * 4;
  static constexpr int ROW = 36;
  static constexpr int WORDS = HDR + MAXC * 3 * ROW;
  static_assert(N == 9 && T::NARM == 1 && T::NOBJ == 1 && !T::HAS_DOOR, "heavy rows: one 9-dof arm, one object, no door");
};

template <class T>
struct HeavyLayout {
  static constexpr int N = T::MD::N, MAXC = T::MAXC;
  static constexpr int ROW4 = 8;                   // float4 per row
  static constexpr int VEC4 = MAXC * 3 * ROW4;     // rows of all contacts: (c * 3 + k) * ROW4
  static constexpr int APP = VEC4 * 4;             // scalar words after the rows: accumulated impulses [MAXC * 3]
  static constexpr int CFM = APP + MAXC * 3;       // cfm * dinv of the normal row [MAXC]
  static constexpr int MU = CFM + MAXC;            // friction coefficient [MAXC]
  static constexpr int WORDS = MU + MAXC;
  static constexpr size_t BYTES = (size_t)WORDS * 32 * sizeof(float);
  static_assert(N == 9 && T::NARM == 1 && T::NOBJ == 1 && !T::HAS_DOOR, "heavy rows: one 9-dof arm, one object, no door");
};

#if defined(__CUDACC__) && !defined(XARM_HOST_SIM)
// Contacts record (thread-local, filled by sub_setup) -> solver rows in shared memory (lane = threadIdx.x & 31)
template <class T>
__device__ __forceinline__ void heavy_rows_fill(const Contacts<T>& C, float* smem, int lane) {
  using L = HeavyLayout<T>;
  float4* rows4 = reinterpret_cast<float4*>(smem);
  const float inv_m = 1.f / T::OBJ_MASS;
  for (int c = 0; c < C.nc; c++) {
    const int sl = C.slot[c];
    const float s1 = C.s1[c];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      float J[16], V[16];
#pragma unroll
      for (int i = 0; i < 9; i++) { J[i] = sl >= 0 ? C.Jarm[sl < 0 ? 0 : sl][k][i] : 0.f; V[i] = sl >= 0 ? C.dVarm[sl < 0 ? 0 : sl][k][i] : 0.f; }
      const V3 d = C.dir[c][k], jo = C.Jo1[c][k], vo = C.dVo1[c][k];
      J[9] = s1 * d.x; J[10] = s1 * d.y; J[11] = s1 * d.z; J[12] = jo.x; J[13] = jo.y; J[14] = jo.z; J[15] = C.rhs[c][k];
      V[9] = s1 * inv_m * d.x; V[10] = s1 * inv_m * d.y; V[11] = s1 * inv_m * d.z; V[12] = vo.x; V[13] = vo.y; V[14] = vo.z; V[15] = C.dinv[c][k];
      float4* r = rows4 + (size_t)((c * 3 + k) * L::ROW4) * 32 + lane;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        r[(size_t)u * 32] = make_float4(J[4 * u], J[4 * u + 1], J[4 * u + 2], J[4 * u + 3]);
        r[(size_t)(4 + u) * 32] = make_float4(V[4 * u], V[4 * u + 1], V[4 * u + 2], V[4 * u + 3]);
      }
      smem[(size_t)(L::APP + c * 3 + k) * 32 + lane] = 0.f;
    }
    smem[(size_t)(L::CFM + c) * 32 + lane] = C.cfmr[c];
    smem[(size_t)(L::MU + c) * 32 + lane] = C.mu[c];
  }
}

// generic setup of a heavy env (thread-local Contacts record; Dpre: the dynamics pass the lean setup already made) ->
// record `rec` for the cooperative solver.  128-bit stores: every thread writes its own 7 KB record.
template <class T>
__device__ __forceinline__ void heavy_rows_record(Env<T>& e, bool apply_damping, bool last, float* __restrict__ rec,
                                                  const ArmDyn<typename T::MD>* Dpre) {
  using R = HeavyRec<T>;
  static_assert(R::HDR % 4 == 0 && R::ROW % 4 == 0 && R::WORDS % 4 == 0, "records are written as float4");
  ArmRows<T> AR;
  SubBase<T> B;
  ManifoldIn MI;
  Contacts<T> C;
  sub_setup<T>(e, apply_damping, last, AR, C, B, MI, Dpre);
  float4* rec4 = reinterpret_cast<float4*>(rec);
  {
    float hd[R::HDR];
#pragma unroll
    for (int i = 0; i < R::HDR; i++) hd[i] = 0.f;
#pragma unroll
    for (int i = 0; i < R::NT; i++) hd[R::MI + i] = AR.Mi[0][i];
#pragma unroll
    for (int i = 0; i < R::N; i++) { hd[R::MRHS + i] = AR.mrhs[0][i]; hd[R::LRHS + i] = AR.lrhs[0][i]; hd[R::QDU + i] = B.qdu[0][i]; }
    hd[R::LIM] = (float)AR.lim_lo[0]; hd[R::LIM + 1] = (float)AR.lim_hi[0];
    hd[R::GEAR] = AR.grhs[0]; hd[R::GEAR + 1] = AR.gdinv[0]; hd[R::GEAR + 2] = AR.gden[0];
    hd[R::VU] = B.vu[0].x; hd[R::VU + 1] = B.vu[0].y; hd[R::VU + 2] = B.vu[0].z;
    hd[R::WU] = B.wu[0].x; hd[R::WU + 1] = B.wu[0].y; hd[R::WU + 2] = B.wu[0].z;
    hd[R::NC] = (float)C.nc;
#pragma unroll
    for (int u = 0; u < R::HDR / 4; u++) rec4[u] = make_float4(hd[4 * u], hd[4 * u + 1], hd[4 * u + 2], hd[4 * u + 3]);
  }
  const float inv_m = 1.f / T::OBJ_MASS;
  for (int c = 0; c < C.nc; c++) {
    const int sl = C.slot[c];
    const float s1 = C.s1[c];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      float r[R::ROW];
#pragma unroll
      for (int i = 0; i < 9; i++) { r[i] = sl >= 0 ? C.Jarm[sl < 0 ? 0 : sl][k][i] : 0.f; r[16 + i] = sl >= 0 ? C.dVarm[sl < 0 ? 0 : sl][k][i] : 0.f; }
      const V3 d = C.dir[c][k], jo = C.Jo1[c][k], vo = C.dVo1[c][k];
      r[9] = s1 * d.x; r[10] = s1 * d.y; r[11] = s1 * d.z; r[12] = jo.x; r[13] = jo.y; r[14] = jo.z; r[15] = 0.f;
      r[25] = s1 * inv_m * d.x; r[26] = s1 * inv_m * d.y; r[27] = s1 * inv_m * d.z; r[28] = vo.x; r[29] = vo.y; r[30] = vo.z; r[31] = 0.f;
      r[32] = C.rhs[c][k]; r[33] = C.dinv[c][k]; r[34] = C.cfmr[c]; r[35] = C.mu[c];
      float4* o = rec4 + (R::HDR + (c * 3 + k) * R::ROW) / 4;
#pragma unroll
      for (int u = 0; u < R::ROW / 4; u++) o[u] = make_float4(r[4 * u], r[4 * u + 1], r[4 * u + 2], r[4 * u + 3]);
    }
  }
}

// sum over the 16 lanes of a half warp (result in every lane of the half)
__device__ __forceinline__ float half_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 8, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 4, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 2, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 1, 16);
  return v;
}

// Cooperative joint loop: 16 lanes per env.  Lane l holds component l of the velocity change u = [dqd 0..8 | dv | dw]
// (lane 15 idle) and row l of the inverse joint-space inertia; a contact row costs one J load, one product, a 4-step
// shuffle sum, the clamp (computed redundantly by all 16 lanes) and one V load + FMA.  Same row order, clamps and exit
// test as sub_solve_generic.  srec: this env's record in shared memory; sapp: its accumulated impulses [MAXC * 3].
// Returns u of this lane.  Both halves of the warp run in lock-step (full-mask shuffles): a half that has finished keeps
// executing with its updates switched off.
template <class T>
__device__ __forceinline__ float heavy_solve_coop(const float* srec, float* sapp, int l) {
  using R = HeavyRec<T>;
  using MD = typename T::MD;
  constexpr int N = R::N;
  float Mrow[N], mrhs[N], iden[N], mapp[N], lrhs[N], lapp[N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    Mrow[i] = l < N ? srec[R::MI + tri(l < N ? l : 0, i)] : 0.f;
    iden[i] = 1.f / srec[R::MI + tri(i, i)];
    mrhs[i] = srec[R::MRHS + i]; lrhs[i] = srec[R::LRHS + i];
    mapp[i] = 0.f; lapp[i] = 0.f;
  }
  const uint32_t lim_lo = (uint32_t)srec[R::LIM], lim_hi = (uint32_t)srec[R::LIM + 1];
  const float grhs = srec[R::GEAR], gdinv = srec[R::GEAR + 1];
  float gapp = 0.f;
  const int nc = (int)srec[R::NC];
  const float hi_arm = (float)(T::ARM_FORCE * T::TIME_STEP), hi_fin = (float)(T::FINGER_FORCE * T::TIME_STEP);
  const float hi_gear = (float)(XARM_GEAR_MAX_FORCE * T::TIME_STEP), hi_lim = (float)XARM_LIMIT_MAX_IMPULSE;
  const float gr = (float)XARM_GEAR_RATIO;
  const float sthr = sqrtf((float)XARM_RESIDUAL_THRESHOLD);
  const float Mgear = Mrow[MD::F1] + gr * Mrow[MD::F2];
  for (int i = 0; i < nc * 3; i++) if (l == 0) sapp[i] = 0.f;
  __syncwarp();
  // the other half may have more contacts: loop bounds are warp-uniform
  const int nc_max = max(nc, __shfl_xor_sync(0xffffffffu, nc, 16));
  float u = 0.f;
  bool done = false;
#define CO_UNIT(s_, i, sign, rhs_, lo_, hi_, app_)                                                               \
  {                                                                                                              \
    float delta = (rhs_) - (sign) * (s_) * iden[i];                                                              \
    const float sum = (app_) + delta;                                                                            \
    const float sumc = fminf(fmaxf(sum, (lo_)), (hi_));                                                          \
    delta = (sumc == sum) ? delta : sumc - (app_);                                                               \
    if (!done) { (app_) = sumc; u += Mrow[i] * ((sign) * delta); bad = bad || fabsf(delta) > sthr * iden[i]; }   \
  }
#define CO_MOTOR(i)                                                                                              \
  { const float hi_ = (i) < 7 ? hi_arm : hi_fin; const float sm_ = __shfl_sync(0xffffffffu, u, (i), 16);         \
    CO_UNIT(sm_, i, 1.f, mrhs[i], -hi_, hi_, mapp[i]) }
#define CO_LIMIT(i)                                                                                              \
  { const float sl_ = __shfl_sync(0xffffffffu, u, (i), 16);  /* the shuffle stays outside the divergent part */  \
    if (lim_lo >> (i) & 1) CO_UNIT(sl_, i, 1.f, lrhs[i], 0.f, hi_lim, lapp[i])                                   \
    else if (lim_hi >> (i) & 1) CO_UNIT(sl_, i, -1.f, lrhs[i], 0.f, hi_lim, lapp[i]) }
#define CO_GEAR()                                                                                                \
  {                                                                                                              \
    const float s_ = __shfl_sync(0xffffffffu, u, MD::F1, 16) + gr * __shfl_sync(0xffffffffu, u, MD::F2, 16);     \
    float delta = grhs - s_ * gdinv;                                                                             \
    const float sum = gapp + delta;                                                                              \
    const float sumc = fminf(fmaxf(sum, -hi_gear), hi_gear);                                                     \
    delta = (sumc == sum) ? delta : sumc - gapp;                                                                 \
    if (!done) { gapp = sumc; u += Mgear * delta; bad = bad || fabsf(delta) > sthr * gdinv; }                    \
  }
  // limit rows exist in few envs: the branch has to be warp-uniform because of the shuffles inside
  const bool any_lim = __any_sync(0xffffffffu, (lim_lo | lim_hi) != 0u);
  for (int it = 0; it < XARM_SOLVER_ITERATIONS; it++) {
    bool bad = false;
    if (it & 1) {
      if (any_lim) { _Pragma("unroll") for (int i = 0; i < N; i++) CO_LIMIT(i) }
      _Pragma("unroll") for (int i = 0; i < N; i++) CO_MOTOR(i)
      CO_GEAR()
    } else {
      CO_GEAR()
      _Pragma("unroll") for (int i = N - 1; i >= 0; i--) CO_MOTOR(i)
      if (any_lim) { _Pragma("unroll") for (int i = N - 1; i >= 0; i--) CO_LIMIT(i) }
    }
    // ---- normal rows (all loads of a row are issued before the shuffle sum)
    for (int c = 0; c < nc_max; c++) {
      const bool on = c < nc && !done;
      const float* r = srec + R::HDR + (c < nc ? c * 3 : 0) * R::ROW;
      const float j = c < nc ? r[l] : 0.f, vv = r[16 + l], rhs0 = r[32], dinv0 = r[33], cfmr0 = r[34];
      const float app0 = sapp[c < nc ? c * 3 : 0];
      const float s_ = half_sum(j * u);
      float d0 = rhs0 - app0 * cfmr0 - s_ * dinv0;
      const float sum = app0 + d0;
      const float sumc = fminf(fmaxf(sum, 0.f), (float)XARM_CONTACT_MAX_IMPULSE);
      d0 = (sumc == sum) ? d0 : sumc - app0;
      if (on) {
        if (l == 0) sapp[c * 3] = sumc;
        bad = bad || fabsf(d0) > sthr * dinv0;
        u += vv * d0;
      }
    }
    __syncwarp();
    // ---- friction pairs (implicit cone)
    for (int c = 0; c < nc_max; c++) {
      const bool on = c < nc && !done;
      const int cc = c < nc ? c : 0;
      const float* r1 = srec + R::HDR + (cc * 3 + 1) * R::ROW;
      const float* r2 = r1 + R::ROW;
      const float ja_ = c < nc ? r1[l] : 0.f, jb_ = c < nc ? r2[l] : 0.f, va_ = r1[16 + l], vb_ = r2[16 + l];
      const float rhs1 = r1[32], di1 = r1[33], rhs2 = r2[32], di2 = r2[33], mu_ = r1[35];
      const float appn = sapp[cc * 3], app1 = sapp[cc * 3 + 1], app2 = sapp[cc * 3 + 2];
      float pa_ = ja_ * u, pb_ = jb_ * u;
      pa_ += __shfl_xor_sync(0xffffffffu, pa_, 8, 16); pb_ += __shfl_xor_sync(0xffffffffu, pb_, 8, 16);
      pa_ += __shfl_xor_sync(0xffffffffu, pa_, 4, 16); pb_ += __shfl_xor_sync(0xffffffffu, pb_, 4, 16);
      pa_ += __shfl_xor_sync(0xffffffffu, pa_, 2, 16); pb_ += __shfl_xor_sync(0xffffffffu, pb_, 2, 16);
      pa_ += __shfl_xor_sync(0xffffffffu, pa_, 1, 16); pb_ += __shfl_xor_sync(0xffffffffu, pb_, 1, 16);
      const float lim = mu_ * appn;
      float da = rhs1 - pa_ * di1, db = rhs2 - pb_ * di2;
      float sa = app1 + da, sb = app2 + db;
      const float l2 = sa * sa + sb * sb;
      if (l2 > lim * lim) {
        const float len = sqrtf(l2);
        if (len > lim) { const float sc = len > 0.f ? lim / len : 0.f; sa *= sc; sb *= sc; da = sa - app1; db = sb - app2; }
      }
      if (on) {
        if (l == 0) { sapp[c * 3 + 1] = sa; sapp[c * 3 + 2] = sb; }
        bad = bad || fabsf(da) > sthr * di1 || fabsf(db) > sthr * di2;
        u += va_ * da + vb_ * db;
      }
    }
    __syncwarp();
    if (!bad) done = true;
    if (__all_sync(0xffffffffu, done)) break;
  }
#undef CO_UNIT
#undef CO_MOTOR
#undef CO_LIMIT
#undef CO_GEAR
  return u;
}

#define HV_DOT(j0, j1, j2, j3)                                                                                   \
  (((j0.x * dqd[0][0] + j0.y * dqd[0][1]) + (j0.z * dqd[0][2] + j0.w * dqd[0][3])) +                             \
   ((j1.x * dqd[0][4] + j1.y * dqd[0][5]) + (j1.z * dqd[0][6] + j1.w * dqd[0][7])) +                             \
   ((j2.x * dqd[0][8] + j2.y * dv.x) + (j2.z * dv.y + j2.w * dv.z)) + ((j3.x * dw.x + j3.y * dw.y) + j3.z * dw.z))
#define HV_AXPY(v0, v1, v2, v3, d_)                                                                              \
  {                                                                                                              \
    dqd[0][0] += v0.x * (d_); dqd[0][1] += v0.y * (d_); dqd[0][2] += v0.z * (d_); dqd[0][3] += v0.w * (d_);      \
    dqd[0][4] += v1.x * (d_); dqd[0][5] += v1.y * (d_); dqd[0][6] += v1.z * (d_); dqd[0][7] += v1.w * (d_);      \
    dqd[0][8] += v2.x * (d_); dv.x += v2.y * (d_); dv.y += v2.z * (d_); dv.z += v2.w * (d_);                     \
    dw.x += v3.x * (d_); dw.y += v3.y * (d_); dw.z += v3.z * (d_);                                               \
  }

// the joint loop over the arm's non-contact rows (registers) and the contact rows in shared memory
template <class T>
__device__ __forceinline__ void heavy_solve(const ArmRows<T>& AR, int nc, float* smem, int lane, SubSol<T>& S) {
  using MD = typename T::MD;
  using L = HeavyLayout<T>;
  constexpr int N = MD::N, NA = T::NARM, NT = N * (N + 1) / 2;
  SOLVER_LOCALS_FROM(AR)
  const float4* rows4 = reinterpret_cast<const float4*>(smem);
  V3 dv = v3(0, 0, 0), dw = v3(0, 0, 0);
  for (int it = 0; it < XARM_SOLVER_ITERATIONS; it++) {
    bool resid_bad = false;
    if (it & 1) {
      ARM_LIMITS_FWD(0) ARM_MOTORS_FWD(0) GEAR_ROW(0)
    } else {
      GEAR_ROW(0) ARM_MOTORS_BWD(0) ARM_LIMITS_BWD(0)
    }
    // ---- normal rows
    for (int c = 0; c < nc; c++) {
      const float4* r = rows4 + (size_t)(c * 3 * L::ROW4) * 32 + lane;
      const float4 j0 = r[0], j1 = r[32], j2 = r[64], j3 = r[96], v0 = r[128], v1 = r[160], v2 = r[192], v3_ = r[224];
      float* pa = smem + (size_t)(L::APP + c * 3) * 32 + lane;
      const float app0 = *pa, cfmr = smem[(size_t)(L::CFM + c) * 32 + lane];
      const float s_ = HV_DOT(j0, j1, j2, j3);
      const float dinv0 = v3_.w;
      float d0 = j3.w - app0 * cfmr - s_ * dinv0;
      float sum = app0 + d0;
      const float sumc = fminf(fmaxf(sum, 0.f), (float)XARM_CONTACT_MAX_IMPULSE);
      d0 = (sumc == sum) ? d0 : sumc - app0;
      *pa = sumc;
      resid_bad = resid_bad || fabsf(d0) > sthr_ * dinv0;
      HV_AXPY(v0, v1, v2, v3_, d0)
    }
    // ---- friction pairs (implicit cone)
    for (int c = 0; c < nc; c++) {
      const float4* r = rows4 + (size_t)((c * 3 + 1) * L::ROW4) * 32 + lane;
      const float4 a0 = r[0], a1 = r[32], a2 = r[64], a3 = r[96], p0 = r[128], p1 = r[160], p2 = r[192], p3 = r[224];
      const float4 b0 = r[256], b1 = r[288], b2 = r[320], b3 = r[352], q0 = r[384], q1 = r[416], q2 = r[448], q3 = r[480];
      float* pa = smem + (size_t)(L::APP + c * 3) * 32 + lane;
      const float lim = smem[(size_t)(L::MU + c) * 32 + lane] * pa[0];
      const float app1 = pa[32], app2 = pa[64];
      const float ja = HV_DOT(a0, a1, a2, a3), jb = HV_DOT(b0, b1, b2, b3);
      const float di1 = p3.w, di2 = q3.w;
      float da = a3.w - ja * di1, db = b3.w - jb * di2;
      float sa = app1 + da, sb = app2 + db;
      const float l2 = sa * sa + sb * sb;
      if (l2 > lim * lim) {
        const float len = sqrtf(l2);
        if (len > lim) { const float sc = len > 0.f ? lim / len : 0.f; sa *= sc; sb *= sc; da = sa - app1; db = sb - app2; }
      }
      pa[32] = sa; pa[64] = sb;
      resid_bad = resid_bad || fabsf(da) > sthr_ * di1 || fabsf(db) > sthr_ * di2;
      HV_AXPY(p0, p1, p2, p3, da)
      HV_AXPY(q0, q1, q2, q3, db)
    }
    if (!resid_bad) break;
  }
#pragma unroll
  for (int i = 0; i < N; i++) S.dqd[0][i] = dqd[0][i];
  S.dv[0] = dv; S.dw[0] = dw; S.ddoor = 0.f;
}
#undef HV_DOT
#undef HV_AXPY

// one substep of a heavy env: generic setup (Contacts record in thread-local memory) -> shared-memory rows -> joint loop
template <class T>
__device__ __forceinline__ void heavy_substep(Env<T>& e, bool apply_damping, bool last, float* smem, int lane) {
  ArmRows<T> AR;
  SubBase<T> B;
  SubSol<T> S;
  ManifoldIn MI;
  int nc;
  {
    Contacts<T> C;
    sub_setup<T>(e, apply_damping, last, AR, C, B, MI);
    heavy_rows_fill<T>(C, smem, lane);
    nc = C.nc;
  }
  heavy_solve<T>(AR, nc, smem, lane, S);
  sub_integrate<T>(e, B, S);
}
// ---- kernels' bodies of the cooperative heavy path
// thread t of the rows kernel: generic setup of heavy env heavy_list[t] -> its record; grasp flags of the last pass
template <class T>
__device__ __forceinline__ void heavy_rows_body(const KArgs& a, int t, int sub, float* __restrict__ hrec) {
  const int64_t i = a.heavy_list[a.heavy_dir * t];
  Env<T> e;
  env_load<T>(e, a.state, a.n, i);
  const bool last = sub == T::NSUB - 1;
  ArmDyn<typename T::MD> D[1];
  dyn_load<T>(D[0], a.scratch, a.n, i);
  heavy_rows_record<T>(e, T::DAMP_EACH || sub == 0, last, hrec + (size_t)(a.heavy_dir > 0 ? t : a.n - 1 - t) * HeavyRec<T>::WORDS, D);
  if (last) {
    const int w = state_words<T>() - 2;
    a.state[(int64_t)w * a.n + i] = e.grasp[0] ? 1.f : 0.f;
    a.state[(int64_t)(w + 1) * a.n + i] = e.grasp[1] ? 1.f : 0.f;
  }
}

// 16 lanes of the solve kernel: record -> shared memory, joint loop, stepPositionsMultiDof, state
template <class T>
__device__ __forceinline__ void heavy_solve_body(const KArgs& a, int t, bool valid, const float* __restrict__ hrec, float* srec, int l) {
  using R = HeavyRec<T>;
  float* sapp = srec + R::WORDS;
  if (valid) {
    const float* rec = hrec + (size_t)(a.heavy_dir > 0 ? t : a.n - 1 - t) * R::WORDS;
    const int words = R::HDR + (int)rec[R::NC] * 3 * R::ROW;
    for (int w = l; w < words; w += 16) srec[w] = rec[w];
  } else {
    for (int w = l; w < R::HDR; w += 16) srec[w] = w < R::NT ? 1.f : 0.f;  // a harmless empty env for the idle half
  }
  __syncwarp();
  const float u = heavy_solve_coop<T>(srec, sapp, l);
  const float h = (float)T::H;
  const float dvx = __shfl_sync(0xffffffffu, u, 9, 16), dvy = __shfl_sync(0xffffffffu, u, 10, 16), dvz = __shfl_sync(0xffffffffu, u, 11, 16);
  const float dwx = __shfl_sync(0xffffffffu, u, 12, 16), dwy = __shfl_sync(0xffffffffu, u, 13, 16), dwz = __shfl_sync(0xffffffffu, u, 14, 16);
  if (valid) {
    const int64_t i = a.heavy_list[a.heavy_dir * t], n = a.n;
    float* st = a.state;
    if (l < R::N) {  // joints: word l = q, word N + l = qd
      const float qd = srec[R::QDU + l] + u;
      st[(int64_t)(R::N + l) * n + i] = qd;
      st[(int64_t)l * n + i] += qd * h;
    } else if (l == R::N) {  // the object (words 3N .. 3N + 12: pos, quat, v, w), as in sub_integrate
      const int w0 = 3 * R::N;
      ObjState b;
      b.pos = v3(st[(int64_t)w0 * n + i], st[(int64_t)(w0 + 1) * n + i], st[(int64_t)(w0 + 2) * n + i]);
      b.quat.x = st[(int64_t)(w0 + 3) * n + i]; b.quat.y = st[(int64_t)(w0 + 4) * n + i]; b.quat.z = st[(int64_t)(w0 + 5) * n + i]; b.quat.w = st[(int64_t)(w0 + 6) * n + i];
      b.v = v3(srec[R::VU] + dvx, srec[R::VU + 1] + dvy, srec[R::VU + 2] + dvz);
      b.w = v3(srec[R::WU] + dwx, srec[R::WU + 1] + dwy, srec[R::WU + 2] + dwz);
      b.pos += h * b.v;
      float ang = norm(b.w);
      if (ang * h > (float)XARM_ANGULAR_MOTION_THRESHOLD) ang = (float)XARM_ANGULAR_MOTION_THRESHOLD / h;
      V3 ax;
      if (ang < 0.001f) ax = (0.5f * h - h * h * h * 0.020833333333f * ang * ang) * b.w;
      else ax = (sinf(0.5f * ang * h) / ang) * b.w;
      Q4 dq = {ax.x, ax.y, ax.z, cosf(ang * h * 0.5f)};
      Q4 qn = quat_mul(dq, b.quat);
      const float inv = rsqrtf(qn.x * qn.x + qn.y * qn.y + qn.z * qn.z + qn.w * qn.w);
      st[(int64_t)w0 * n + i] = b.pos.x; st[(int64_t)(w0 + 1) * n + i] = b.pos.y; st[(int64_t)(w0 + 2) * n + i] = b.pos.z;
      st[(int64_t)(w0 + 3) * n + i] = qn.x * inv; st[(int64_t)(w0 + 4) * n + i] = qn.y * inv; st[(int64_t)(w0 + 5) * n + i] = qn.z * inv; st[(int64_t)(w0 + 6) * n + i] = qn.w * inv;
      st[(int64_t)(w0 + 7) * n + i] = b.v.x; st[(int64_t)(w0 + 8) * n + i] = b.v.y; st[(int64_t)(w0 + 9) * n + i] = b.v.z;
      st[(int64_t)(w0 + 10) * n + i] = b.w.x; st[(int64_t)(w0 + 11) * n + i] = b.w.y; st[(int64_t)(w0 + 12) * n + i] = b.w.z;
    }
  }
  __syncwarp();
}
#endif
