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

// ---- Low-latency cooperative joint loop in IMPULSE space (Delassus form), 16 lanes per env.
// The velocity form above pays a 4-level shuffle sum (J_r . u) on the critical path of EVERY row: ~250 cycles per row,
// ~10 k cycles per sweep, 250 us per substep for a warp that runs alone - and the auto-reset tail of a step is a chain
// of 105 such substeps (profiles/).  Here every lane OWNS rows and keeps s_r = J_r . u of its rows up to date instead:
//   lane c < nc : the three rows of contact c (normal, friction 1, friction 2)
//   lane i < 9  : also motor i and the limit row of dof i (J = +-e_i, so s_limit = sign * s_motor: no extra storage)
//   lane 9      : also the gear row
// Processing row k = the owner computes delta_k from its own s_k (registers only), ONE shuffle broadcasts it, and every
// lane adds A[r][k] * delta_k to its <= 4 sums (A = J M^-1 J^T, built once per substep in shared memory, layout
// A[(k * 4 + slot) * 16 + lane]: conflict-free, and the loads do not depend on the solve so they are issued ahead).
// Critical path per row: delta (~7 dependent FP ops) + shuffle + FMA ~ 60 cycles.  Same row order, clamps, cone and exit
// test as sub_solve_generic / btMultiBodyConstraintSolver::solveSingleIteration; rounding differs from the
// velocity form (sums are accumulated per row, u is formed once at the end from the accumulated impulses).
// Row numbering k: 0..8 motors, 9 gear, 10 + 3 c + j contact rows.  nc <= XARM_DELA_MAXC = T::MAXC (one lane per contact).
// Validation: tools/_emul_pgs.py replays dumped records in both forms, float32 against float64 - both forms sit at
// 5e-5..2e-4 of the float64 sweep on ordinary grasps, and both lose all digits on the same (diverging) records.
#define XARM_DELA_MAXC 16
template <class T>
struct DelaLayout {
  static constexpr int NU = 10;                                  // unit columns: 9 motors + gear
  static constexpr int RMAX = NU + 3 * XARM_DELA_MAXC;           // 58
  static constexpr int A_WORDS = RMAX * 4 * 16;                  // 3712
  static constexpr int VU = A_WORDS;                             // M^-1 J^T of the unit rows [NU][16]
  static constexpr int LAM = VU + NU * 16;                       // accumulated impulses [RMAX] (+ pad)
  static constexpr int WORDS_ = LAM + 64;
  static constexpr int W0 = WORDS_;
  static constexpr int SLOT = ((W0 + 31) / 32) * 32 + 16;        // = 16 mod 32: the two envs of a warp use different banks
};

// reads of the record (written by k_heavy_rows in the previous launch): XARM_DELA_LDCG routes them past L1
#ifdef XARM_DELA_LDCG
#define RL(p) __ldcg(p)
#define RL4(p) __ldcg(p)
#else
#define RL(p) (*(p))
#define RL4(p) (*(p))
#endif
// COMPACT (the fused kernel, record in shared memory): a row is [M^-1 J^T (15) pad | rhs dinv cfmr mu] = 20 words and the lane's
// own J rows arrive in registers (Jr[3][16]: the lane that built them keeps them) - 4.2 KB instead of 7.2 KB of rows per env.
template <class T, bool COMPACT = false>
__device__ __forceinline__ float heavy_solve_dela(const float* __restrict__ rec, float* sm, int l, bool valid, const float (*Jr)[16] = nullptr) {
  using R = HeavyRec<T>;
  using D = DelaLayout<T>;
  constexpr int RS = COMPACT ? 20 : R::ROW, VO = COMPACT ? 0 : 16, CO = COMPACT ? 16 : 32;
  using MD = typename T::MD;
  constexpr int N = R::N;
  constexpr unsigned FULL = 0xffffffffu;
#ifdef XARM_FUSED_PROF
  const long long dp0_ = clock64();
#endif
  float* A = sm;
  float* Vu = sm + D::VU;
  float* lam = sm + D::LAM;
  const float gr = (float)XARM_GEAR_RATIO;
  const int nc = valid ? (int)RL(rec + R::NC) : 0;
  const int Rn = D::NU + 3 * nc;
  const bool own_c = l < nc;
  // ---- unit columns: M^-1 e_k (motors), M^-1 (e_F1 + gr e_F2) (gear); the object part is zero
#pragma unroll
  for (int k = 0; k < D::NU; k++) {
    float v = 0.f;
    if (valid && l < N) v = k < N ? RL(rec + R::MI + tri(l, k)) : RL(rec + R::MI + tri(l, MD::F1)) + gr * RL(rec + R::MI + tri(l, MD::F2));
    Vu[k * 16 + l] = v;
  }
  // ---- this lane's contact rows (J) and row constants
  float Jn[16], Ja[16], Jb[16];
  float rhs0 = 0.f, dinv0 = 1.f, cfmr0 = 0.f, rhs1 = 0.f, di1 = 1.f, rhs2 = 0.f, di2 = 1.f, mu = 0.f;
  {
    const float4* r4 = reinterpret_cast<const float4*>(rec + R::HDR + (own_c ? l * 3 : 0) * RS);
    if constexpr (COMPACT) {
#pragma unroll
      for (int q = 0; q < 16; q++) { Jn[q] = own_c ? Jr[0][q] : 0.f; Ja[q] = own_c ? Jr[1][q] : 0.f; Jb[q] = own_c ? Jr[2][q] : 0.f; }
    } else {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const float4 a = own_c ? RL4(r4 + u) : make_float4(0, 0, 0, 0);
        const float4 b = own_c ? RL4(r4 + RS / 4 + u) : make_float4(0, 0, 0, 0);
        const float4 c = own_c ? RL4(r4 + 2 * (RS / 4) + u) : make_float4(0, 0, 0, 0);
        Jn[4 * u] = a.x; Jn[4 * u + 1] = a.y; Jn[4 * u + 2] = a.z; Jn[4 * u + 3] = a.w;
        Ja[4 * u] = b.x; Ja[4 * u + 1] = b.y; Ja[4 * u + 2] = b.z; Ja[4 * u + 3] = b.w;
        Jb[4 * u] = c.x; Jb[4 * u + 1] = c.y; Jb[4 * u + 2] = c.z; Jb[4 * u + 3] = c.w;
      }
    }
    if (own_c) {
      const float4 k0 = RL4(r4 + CO / 4), k1 = RL4(r4 + RS / 4 + CO / 4), k2 = RL4(r4 + 2 * (RS / 4) + CO / 4);
      rhs0 = k0.x; dinv0 = k0.y; cfmr0 = k0.z; mu = k0.w;
      rhs1 = k1.x; di1 = k1.y; rhs2 = k2.x; di2 = k2.y;
    }
  }
  // unit row of this lane: motor l (l < 9) or the gear (l == 9); limit row of dof l
  float urhs = 0.f, uden = 1.f, uhi = 0.f, lrhs = 0.f, lsign = 0.f;
  if (valid && l < N) {
    urhs = RL(rec + R::MRHS + l); uden = 1.f / RL(rec + R::MI + tri(l, l));
    uhi = l < 7 ? (float)(T::ARM_FORCE * T::TIME_STEP) : (float)(T::FINGER_FORCE * T::TIME_STEP);
    lrhs = RL(rec + R::LRHS + l);
    const uint32_t lo = (uint32_t)RL(rec + R::LIM), hi = (uint32_t)RL(rec + R::LIM + 1);
    lsign = (lo >> l & 1u) ? 1.f : ((hi >> l & 1u) ? -1.f : 0.f);
  } else if (valid && l == N) {
    urhs = RL(rec + R::GEAR); uden = RL(rec + R::GEAR + 1); uhi = (float)(XARM_GEAR_MAX_FORCE * T::TIME_STEP);
  }
  const float hi_lim = (float)XARM_LIMIT_MAX_IMPULSE;
  const float sthr = sqrtf((float)XARM_RESIDUAL_THRESHOLD);
  __syncwarp();
  // ---- A = J M^-1 J^T: column k against this lane's rows
#define DELA_COLUMN(k_, V0, V1, V2, V3, vl_)                                                                      \
  {                                                                                                               \
    const float vv_[16] = {V0.x, V0.y, V0.z, V0.w, V1.x, V1.y, V1.z, V1.w, V2.x, V2.y, V2.z, V2.w, V3.x, V3.y, V3.z, V3.w}; \
    float an_ = 0.f, aa_ = 0.f, ab_ = 0.f;                                                                        \
    _Pragma("unroll") for (int q_ = 0; q_ < 15; q_++) { an_ += Jn[q_] * vv_[q_]; aa_ += Ja[q_] * vv_[q_]; ab_ += Jb[q_] * vv_[q_]; } \
    float* o_ = A + (k_) * 64 + l;                                                                                \
    o_[0] = an_; o_[16] = aa_; o_[32] = ab_; o_[48] = (vl_);                                                      \
  }
  for (int k = 0; k < D::NU; k++) {
    const float4* v4 = reinterpret_cast<const float4*>(Vu + k * 16);
    const float4 V0 = v4[0], V1 = v4[1], V2 = v4[2], V3 = v4[3];
    // unit slot: e_l . V_k (motor), (e_F1 + gr e_F2) . V_k (gear)
    const float vl = l < N ? Vu[k * 16 + l] : (l == N ? Vu[k * 16 + MD::F1] + gr * Vu[k * 16 + MD::F2] : 0.f);
    DELA_COLUMN(k, V0, V1, V2, V3, vl)
  }
#pragma unroll 4
  for (int k = D::NU; k < Rn; k++) {
    const float* row = rec + R::HDR + (k - D::NU) * RS + VO;
    const float4* v4 = reinterpret_cast<const float4*>(row);
    const float4 V0 = RL4(v4), V1 = RL4(v4 + 1), V2 = RL4(v4 + 2), V3 = RL4(v4 + 3);
    const float vl = l < N ? RL(row + l) : (l == N ? RL(row + MD::F1) + gr * RL(row + MD::F2) : 0.f);
    DELA_COLUMN(k, V0, V1, V2, V3, vl)
  }
#undef DELA_COLUMN
  __syncwarp();
  // ---- sweeps.  Row update in its short-chain form: with c = app + rhs - app * cfm kept ready (it only changes in the
  // row's own step), x = c - s * dinv, app' = clamp(x), delta = app' - app: on the critical path of a row remain
  // FFMA (s) -> FFMA (x) -> FMNMX -> FMNMX -> FADD -> SHFL.  (Same fixed point and clamps as the textbook form
  // delta = rhs - app cfm - s dinv, app' = clamp(app + delta); tools/_emul_pgs.py: same float32 error against float64.)
  // Limit rows carry the SIGNED impulse q = sign * app in sign * [0, hi]: x = q + sign * rhs - s3 * den, and the
  // broadcast delta is already the joint-space one.  A finished env (exit test) has its rows switched off: dinv = 0,
  // c = app, so every delta is exactly zero.
#ifdef XARM_FUSED_PROF
  const long long dp1_ = clock64();
  int its_ = 0;
#endif
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  float appn = 0.f, app1 = 0.f, app2 = 0.f, mapp = 0.f, q = 0.f;
#ifdef XARM_DELA_NCMAX_SHFL
  const int nc_max = max(nc, __shfl_xor_sync(FULL, nc, 16));
#else
  // REDUX leaves the bound in a uniform register: the contact loops are then provably convergent and their shuffles are plain
  // SHFL.IDX (with a per-lane bound every one of them sits inside a WARPSYNC.COLLECTIVE ... ENDCOLLECTIVE bracket)
  const int nc_max = __reduce_max_sync(FULL, nc);
#endif
  const uint32_t lim_bits = __ballot_sync(FULL, lsign != 0.f);
  const uint32_t lim_any = (lim_bits | (lim_bits >> 16)) & 0x1ffu;  // dofs with a limit row in either env of the warp
  bool done = !valid;
  const float llo = lsign < 0.f ? -hi_lim : 0.f, lhi = lsign > 0.f ? hi_lim : 0.f;
  float lrs = lsign * lrhs;
  if (!own_c) { dinv0 = 0.f; di1 = 0.f; di2 = 0.f; }
  const float dinv_t = dinv0, di1_t = di1, di2_t = di2;
  if (done) { uden = 0.f; urhs = 0.f; }
  float uden_t = uden;              // residual test scale (kept when the rows are switched off)
  float cu = urhs, cl = lrs, cn = rhs0, ca = rhs1, cb = rhs2;
#define DELA_APPLY(k_, d_)                                                                                        \
  { const float* c_ = A + (k_) * 64 + l; s0 += c_[0] * (d_); s1 += c_[16] * (d_); s2 += c_[32] * (d_); s3 += c_[48] * (d_); }
#define DELA_UNIT(o_)                                                                                             \
  {                                                                                                               \
    const float x_ = fmaf(-s3, uden, cu);                                                                         \
    const float xn_ = fminf(fmaxf(x_, -uhi), uhi);                                                                \
    const float delta = xn_ - mapp;                                                                               \
    const float d_ = __shfl_sync(FULL, delta, (o_), 16);                                                          \
    const bool own_ = l == (o_);                                                                                  \
    mapp = own_ ? xn_ : mapp; cu = mapp + urhs;                                                                   \
    bad = bad || (own_ && fabsf(delta) > sthr * uden_t);                                                          \
    DELA_APPLY((o_), d_)                                                                                          \
  }
#define DELA_LIMIT(o_)                                                                                            \
  if (lim_any >> (o_) & 1u) {                                                                                     \
    const float x_ = fmaf(-s3, uden, cl);                                                                         \
    const float xn_ = fminf(fmaxf(x_, llo), lhi);                                                                 \
    const float delta = xn_ - q;                                                                                  \
    const float d_ = __shfl_sync(FULL, delta, (o_), 16);                                                          \
    const bool own_ = l == (o_);                                                                                  \
    q = own_ ? xn_ : q; cl = q + lrs;                                                                             \
    bad = bad || (own_ && fabsf(delta) > sthr * uden_t);                                                          \
    DELA_APPLY((o_), d_)                                                                                          \
  }
#ifndef XARM_DELA_UNROLL_STR
#define XARM_DELA_UNROLL_STR "unroll 2"   /* contact loops of the sweep: a lone warp pays the loop branch on its critical path (measured: 1 -> 2: -1.4 ms per step, 4: same) */
#endif
  for (int it = 0; it < XARM_SOLVER_ITERATIONS; it++) {
    bool bad = false;
    if (it & 1) {
      if (lim_any) { _Pragma("unroll") for (int i = 0; i < N; i++) DELA_LIMIT(i) }
      _Pragma("unroll") for (int i = 0; i < N; i++) DELA_UNIT(i)
      DELA_UNIT(N)
    } else {
      DELA_UNIT(N)
      _Pragma("unroll") for (int i = N - 1; i >= 0; i--) DELA_UNIT(i)
      if (lim_any) { _Pragma("unroll") for (int i = N - 1; i >= 0; i--) DELA_LIMIT(i) }
    }
    // ---- normal rows.  The Delassus column of row c is loaded BEFORE the row's update chain (it does not depend on the
    // solve): a lane past this env's contacts (the other env of the warp has more) broadcasts an exact zero, so it may read
    // any finite column - column 0, always built - and the loads need no branch (shared-memory latency off the critical path).
    _Pragma(XARM_DELA_UNROLL_STR)
    for (int c = 0; c < nc_max; c++) {
      const float* c_ = A + (c < nc ? D::NU + 3 * c : 0) * 64 + l;
      const float a0 = c_[0], a1 = c_[16], a2 = c_[32], a3 = c_[48];
      const float x_ = fmaf(-s0, dinv0, cn);
      const float xn_ = fminf(fmaxf(x_, 0.f), (float)XARM_CONTACT_MAX_IMPULSE);
      const float d0 = xn_ - appn;
      const float d_ = __shfl_sync(FULL, d0, c, 16);
      const bool own_ = l == c;
      appn = own_ ? xn_ : appn; cn = fmaf(-appn, cfmr0, appn + rhs0);
      bad = bad || (own_ && fabsf(d0) > sthr * dinv_t);
      s0 = fmaf(a0, d_, s0); s1 = fmaf(a1, d_, s1); s2 = fmaf(a2, d_, s2); s3 = fmaf(a3, d_, s3);
    }
    // ---- friction pairs (implicit cone)
    const float lim = mu * appn, lim2 = lim * lim;
    _Pragma(XARM_DELA_UNROLL_STR)
    for (int c = 0; c < nc_max; c++) {
      const float* c_ = A + (c < nc ? D::NU + 3 * c + 1 : 0) * 64 + l;
      const float p0 = c_[0], p1 = c_[16], p2 = c_[32], p3 = c_[48], q0 = c_[64], q1 = c_[80], q2 = c_[96], q3 = c_[112];
      float xa = fmaf(-s1, di1, ca), xb = fmaf(-s2, di2, cb);
      const float l2 = xa * xa + xb * xb;
      // rsqrtf() wraps MUFU.RSQ in a denormal fix-up (two compares, two multiplies) that sits on the row's dependency chain;
      // rsqrt.approx.ftz on max(l2, 1e-30) is the same MUFU result for every l2 >= 1e-30 and stays finite below (a pair of
      // impulses < 1e-15 against a cone radius that is smaller still): friction step 166 -> 131 cycles (tools/micro/unit_micro.cu)
      float rs_;
      asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs_) : "f"(fmaxf(l2, 1e-30f)));
      const float sc = l2 > lim2 ? lim * rs_ : 1.f;
      xa *= sc; xb *= sc;
      const float da = xa - app1, db = xb - app2;
      const float da_ = __shfl_sync(FULL, da, c, 16), db_ = __shfl_sync(FULL, db, c, 16);
      const bool own_ = l == c;
      app1 = own_ ? xa : app1; app2 = own_ ? xb : app2; ca = app1 + rhs1; cb = app2 + rhs2;
      bad = bad || (own_ && (fabsf(da) > sthr * di1_t || fabsf(db) > sthr * di2_t));
      s0 = fmaf(q0, db_, fmaf(p0, da_, s0)); s1 = fmaf(q1, db_, fmaf(p1, da_, s1));
      s2 = fmaf(q2, db_, fmaf(p2, da_, s2)); s3 = fmaf(q3, db_, fmaf(p3, da_, s3));
    }
    const uint32_t bb = __ballot_sync(FULL, bad);
    if (!done && ((bb >> (threadIdx.x & 16)) & 0xffffu) == 0u) {  // this env is finished: switch its rows off
      done = true;
      uden = 0.f; urhs = 0.f; cu = mapp; lrs = 0.f; cl = q;
      dinv0 = 0.f; cfmr0 = 0.f; rhs0 = 0.f; cn = appn;
      di1 = 0.f; di2 = 0.f; rhs1 = 0.f; rhs2 = 0.f; ca = app1; cb = app2; mu = 1e30f;   // (cone never bites: impulses stay)
    }
#ifdef XARM_FUSED_PROF
    its_++;
#endif
    if (__all_sync(FULL, done)) break;
  }
#ifdef XARM_FUSED_PROF
  if (blockIdx.x == 0 && threadIdx.x == 0) printf("[dela prof] nc %d nc_max %d | build %lld | %d sweeps %lld cycles\n", nc, nc_max, dp1_ - dp0_, its_, clock64() - dp1_);
#endif
#undef DELA_APPLY
#undef DELA_UNIT
#undef DELA_LIMIT
  // ---- u = M^-1 J^T lambda, component l (lane 15 idle)
  if (l < N) lam[l] = mapp + q; else if (l == N) lam[N] = mapp;
  if (own_c) { lam[D::NU + 3 * l] = appn; lam[D::NU + 3 * l + 1] = app1; lam[D::NU + 3 * l + 2] = app2; }
  __syncwarp();
  float u0 = 0.f, u1 = 0.f;
#pragma unroll
  for (int k = 0; k < D::NU; k++) u0 += Vu[k * 16 + l] * lam[k];
  for (int k = D::NU; k < Rn; k += 2) {
    u0 += RL(rec + R::HDR + (k - D::NU) * RS + VO + l) * lam[k];
    if (k + 1 < Rn) u1 += RL(rec + R::HDR + (k + 1 - D::NU) * RS + VO + l) * lam[k + 1];
  }
  __syncwarp();
  return u0 + u1;
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
    a.state[(int64_t)w * a.n + i] = grasp_word(e.grasp[0], e.grasp_cmd[0]);
    a.state[(int64_t)(w + 1) * a.n + i] = grasp_word(e.grasp[1], e.grasp_cmd[1]);
  }
}

// stepPositionsMultiDof of a heavy env by its 16 lanes: u = velocity change of component l; hdr = the record's header
template <class T>
__device__ __forceinline__ void heavy_integrate_coop(const KArgs& a, int t, bool valid, const float* hdr, float u, int l) {
  using R = HeavyRec<T>;
  const float h = (float)T::H;
  const float dvx = __shfl_sync(0xffffffffu, u, 9, 16), dvy = __shfl_sync(0xffffffffu, u, 10, 16), dvz = __shfl_sync(0xffffffffu, u, 11, 16);
  const float dwx = __shfl_sync(0xffffffffu, u, 12, 16), dwy = __shfl_sync(0xffffffffu, u, 13, 16), dwz = __shfl_sync(0xffffffffu, u, 14, 16);
  if (valid) {
    const int64_t i = a.heavy_list[a.heavy_dir * t], n = a.n;
    float* st = a.state;
    if (l < R::N) {  // joints: word l = q, word N + l = qd
      const float qd = hdr[R::QDU + l] + u;
      st[(int64_t)(R::N + l) * n + i] = qd;
      st[(int64_t)l * n + i] += qd * h;
    } else if (l == R::N) {  // the object (words 3N .. 3N + 12: pos, quat, v, w), as in sub_integrate
      const int w0 = 3 * R::N;
      ObjState b;
      b.pos = v3(st[(int64_t)w0 * n + i], st[(int64_t)(w0 + 1) * n + i], st[(int64_t)(w0 + 2) * n + i]);
      b.quat.x = st[(int64_t)(w0 + 3) * n + i]; b.quat.y = st[(int64_t)(w0 + 4) * n + i]; b.quat.z = st[(int64_t)(w0 + 5) * n + i]; b.quat.w = st[(int64_t)(w0 + 6) * n + i];
      b.v = v3(hdr[R::VU] + dvx, hdr[R::VU + 1] + dvy, hdr[R::VU + 2] + dvz);
      b.w = v3(hdr[R::WU] + dwx, hdr[R::WU + 1] + dwy, hdr[R::WU + 2] + dwz);
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
}

// ---- Fused heavy substep: collision, rows and the impulse-space joint loop in ONE launch, 16 lanes per env.
// Round 1 ran k_heavy_rows (one THREAD per env: generic collision + rows through a 5.8 KB thread-local Contacts record into
// a 7.2 KB global record) and then k_heavy_solve2.  In the auto-reset tail of a step that is two dependent launches per
// substep, each at the latency floor of a lone warp (~60 + ~115 us), 105 substeps in sequence.  Here the 16 lanes of the env
// share that work: lane p < 4 collides pair p (lego-table, finger1-lego, finger2-lego, hand-lego: the fixed pair order of
// sub_setup; with MAXC = 16 = 4 pairs x 4 points and XARM_MAXAC = 12 = 3 gripper pairs x 4 no cap can bite, so the pairs are
// independent), lane c < nc builds the three rows of contact c with the arithmetic of sub_setup's row loop
// (setupMultiBodyContactConstraint), lanes 0..9 build the arm's motor / limit / gear rows (arm_rows), and the
// record stays in shared memory, in the HeavyRec layout heavy_solve_dela reads.  Same results as k_heavy_rows + k_heavy_solve2
// bit for bit (tests/test_gpu_parity.py::test_fused_heavy_kernel_equals_rows_then_solve).
template <class T>
struct FusedLayout {
  using R = HeavyRec<T>;
  static constexpr int CROW = 20;                     // compact row: M^-1 J^T (15) pad | rhs dinv cfmr mu (J stays in the owner lane's registers)
  static constexpr int REC = 0;                       // the record: header + compact rows
  static constexpr int DELA = (R::HDR + T::MAXC * 3 * CROW + 31) / 32 * 32;   // DelaLayout slot (bank-aligned: a multiple of 32 words)
  // scratch that lives where the Delassus matrix is built later: dynamics pass of the setup kernel, env state, contact points
  static constexpr int DYN = DELA;                    // 126 words (dyn_store order)
  static constexpr int ST = DYN + 128;                // state words 0..39 (q, qd, qt, object) | 40, 41: the two grasp words
  static constexpr int PTS = ST + 48;                 // [4 pairs][4 points][pa 3 | pb 3 | n 3 | depth]
  static constexpr int CNT = PTS + 160;               // points per pair [4]
  static constexpr int SLOT = DELA + DelaLayout<T>::SLOT;   // = 16 mod 32: the two envs of a warp use different banks
  static_assert(DELA % 32 == 0 && SLOT % 32 == 16 && CNT + 4 <= DELA + DelaLayout<T>::A_WORDS, "fused heavy kernel: shared-memory layout");
  static_assert(T::NARM == 1 && T::NOBJ == 1 && T::NTABLE == 1 && !T::HAS_GROUND && !T::FINGER_TABLE && T::MAXC >= 16 && XARM_MAXAC >= 12,
                "fused heavy kernel: one arm, one object, one table; the contact caps never bite");
};

template <class T>
__device__ __forceinline__ void heavy_fused_body(const KArgs& a, int t, bool valid, int sub, float* sm, int l) {
  using R = HeavyRec<T>;
  using F = FusedLayout<T>;
  using MD = typename T::MD;
  constexpr int N = R::N;
  constexpr unsigned FULL = 0xffffffffu;
  const float h = (float)T::H;
  const int64_t n = a.n;
  const int64_t i = valid ? (int64_t)a.heavy_list[a.heavy_dir * t] : 0;
#ifdef XARM_FUSED_PROF
  long long pc_[8]; int pn_ = 0;
#define FPROF() pc_[pn_++] = clock64();
#else
#define FPROF()
#endif
  FPROF()
  float* rec = sm + F::REC;
  float* dyn = sm + F::DYN;
  float* st = sm + F::ST;
  float* pts = sm + F::PTS;
  float* cnt = sm + F::CNT;
  // ---- the dynamics pass the setup kernel left in the scratch slab, and the env state
  if (valid) {
    for (int w = l; w < pipe_dyn_words<T>(); w += 16) dyn[w] = a.scratch[(int64_t)w * n + i];
    for (int w = l; w < 40; w += 16) st[w] = a.state[(int64_t)w * n + i];
    if (l < 2) st[40 + l] = a.state[(int64_t)(state_words<T>() - 2 + l) * n + i];
  }
  __syncwarp();
  const float* Minv = dyn + 6 * N;
  const float* qdu = Minv + R::NT;
  const float* Rh = qdu + N;
  const V3 opos = v3(st[27], st[28], st[29]);
  Q4 oq; oq.x = st[30]; oq.y = st[31]; oq.z = st[32]; oq.w = st[33];
  const M3 oR = quat_to_m3(oq);
  const int grasp_cmd0 = ((int)st[40] >> 1) & 1;
  FPROF()
  // ---- 1. collision: lane p collides pair p
  if (valid && l < 4) {
    Box ob; ob.c = opos; ob.R = oR; ob.h = v3(T::OBJ_HX, T::OBJ_HY, T::OBJ_HZ);
    Box A, B;
    if (l == 0) {
      A = ob;
      B.c = v3(T::table_x(0), 0.f, -(float)XARM_TABLE_HALF_Z); B.R = m3_identity();
      B.h = v3((float)XARM_TABLE_HALF_X, (float)XARM_TABLE_HALF_Y, (float)XARM_TABLE_HALF_Z);
    } else {
      ArmDyn<MD> D;
#pragma unroll
      for (int k = 0; k < 9; k++) D.Rh.m[k] = Rh[k];
      D.ph = v3(Rh[9], Rh[10], Rh[11]); D.pf1 = v3(Rh[12], Rh[13], Rh[14]); D.pf2 = v3(Rh[15], Rh[16], Rh[17]);
      A = arm_box<T>(D, l == 3 ? 0 : l);   // pair 1: finger 1, pair 2: finger 2, pair 3: hand
      B = ob;
    }
    int k = 0;
    CPoint cp[4];
    const V3 d = A.c - B.c;
    const float rr = norm(A.h) + norm(B.h) + (float)XARM_CONTACT_MARGIN;
    if (!(dot(d, d) > rr * rr)) k = box_box(A, B, cp, 4);
    cnt[l] = (float)k;
    for (int j = 0; j < k; j++) {
      float* o = pts + (l * 4 + j) * 10;
      o[0] = cp[j].pa.x; o[1] = cp[j].pa.y; o[2] = cp[j].pa.z; o[3] = cp[j].pb.x; o[4] = cp[j].pb.y; o[5] = cp[j].pb.z;
      o[6] = cp[j].n.x; o[7] = cp[j].n.y; o[8] = cp[j].n.z; o[9] = cp[j].depth;
    }
  }
  __syncwarp();
  FPROF()
  const int c0 = valid ? (int)cnt[0] : 0, c1 = valid ? (int)cnt[1] : 0, c2 = valid ? (int)cnt[2] : 0, c3 = valid ? (int)cnt[3] : 0;
  const int nc = c0 + c1 + c2 + c3;
  // ---- 2. unconstrained velocity of the object, its world inverse inertia (every lane: a few dozen flops)
  const V3 bv = v3(st[34], st[35], st[36]), bw = v3(st[37], st[38], st[39]);
  const float kl = (float)XARM_MB_LINEAR_DAMPING * (1.f + norm(bv)), ka = (float)XARM_MB_ANGULAR_DAMPING * (1.f + norm(bw));
  V3 vu = bv + h * ((-kl) * bv); vu.z -= h * (float)XARM_GRAVITY;
  const V3 wu = bw + h * ((-ka) * bw);
  const float lx = 2 * T::OBJ_HX, ly = 2 * T::OBJ_HY, lz = 2 * T::OBJ_HZ, m12 = T::OBJ_MASS / 12.f;
  const S3 Il = {1.f / (m12 * (ly * ly + lz * lz)), 0, 0, 1.f / (m12 * (lx * lx + lz * lz)), 0, 1.f / (m12 * (lx * lx + ly * ly))};
  const S3 Iinv = rotate_sym(oR, Il);
  // ---- 3. the header: inverse joint inertia, motor / limit / gear rows (arm_rows), unconstrained velocities, nc
  {
    for (int w = l; w < R::HDR; w += 16) rec[w] = w < R::NT ? Minv[w] : 0.f;
    __syncwarp();
    uint32_t lo = 0u, hi = 0u;
    if (l < N) {
      const float den = Minv[tri(l, l)], qu = qdu[l], q = st[l], qt = st[2 * N + l];
      const float pen_lo = q - MD::lo(l), pen_hi = MD::hi(l) - q;
      float lr = 0.f;
      if (pen_lo <= 0.f) { lo = 1u; lr = (-pen_lo * (float)XARM_ERP / h - qu) / den; }
      else if (pen_hi <= 0.f) { hi = 1u; lr = (-pen_hi * (float)XARM_ERP / h + qu) / den; }
      const float target = (float)(XARM_MOTOR_KP * XARM_MOTOR_ERP) * (qt - q) / h;
      rec[R::LRHS + l] = lr;
      rec[R::MRHS + l] = (target - qu) / den;
      rec[R::QDU + l] = qu;
    }
    const uint32_t blo = (__ballot_sync(FULL, lo != 0u) >> (threadIdx.x & 16)) & 0x1ffu;
    const uint32_t bhi = (__ballot_sync(FULL, hi != 0u) >> (threadIdx.x & 16)) & 0x1ffu;
    if (l == N) {
      const float gr = (float)XARM_GEAR_RATIO;
      const int f1 = MD::F1, f2 = MD::F2 < 0 ? 0 : MD::F2;
      const float den = Minv[tri(f1, f1)] + 2.f * gr * Minv[tri(f1, f2)] + gr * gr * Minv[tri(f2, f2)];
      const float rel = qdu[f1] + gr * qdu[f2];
      const float pos_err = (float)XARM_GEAR_ERP * (st[f1] + gr * st[f2]);
      rec[R::GEAR] = (-pos_err * (float)XARM_ERP / h - rel) / den; rec[R::GEAR + 1] = 1.f / den; rec[R::GEAR + 2] = den;
      rec[R::LIM] = (float)blo; rec[R::LIM + 1] = (float)bhi;
      rec[R::VU] = vu.x; rec[R::VU + 1] = vu.y; rec[R::VU + 2] = vu.z;
      rec[R::WU] = wu.x; rec[R::WU + 1] = wu.y; rec[R::WU + 2] = wu.z;
      rec[R::NC] = (float)nc;
    }
  }
  FPROF()
  // ---- 4. rows of contact l (normal + two friction directions), as sub_setup's row loop
  float Jr[3][16];
#pragma unroll
  for (int k = 0; k < 3; k++)
#pragma unroll
    for (int q = 0; q < 16; q++) Jr[k][q] = 0.f;
  if (l < nc) {
    const int p = l < c0 ? 0 : (l < c0 + c1 ? 1 : (l < c0 + c1 + c2 ? 2 : 3));
    const int j = l - (p == 0 ? 0 : (p == 1 ? c0 : (p == 2 ? c0 + c1 : c0 + c1 + c2)));
    const float* o = pts + (p * 4 + j) * 10;
    const V3 pa = v3(o[0], o[1], o[2]), pb = v3(o[3], o[4], o[5]), nrm = v3(o[6], o[7], o[8]);
    const float depth = o[9];
    const bool arm = p > 0;            // pairs 1..3: side A is a gripper link, side B the object; pair 0: side A the object
    const bool soft = p == 1 || p == 2;
    const float fo = (float)XARM_DEFAULT_FRICTION;
    const float ff = (T::FRICTION_SWITCH && grasp_cmd0) ? (float)XARM_FINGER_FRICTION_GRASP : (float)XARM_FINGER_FRICTION_FREE;
    const float mu = fminf(p == 0 ? fo * (float)XARM_TABLE_FRICTION : (soft ? ff * fo : fo * fo), (float)XARM_MAX_FRICTION);
    float erp = (float)XARM_ERP2, cfm0 = 0.f;
    if (soft) {
      const float ks = (float)XARM_FINGER_STIFFNESS, kd = (float)XARM_FINGER_DAMPING;
      cfm0 = 1.f / (h * ks + kd); erp = h * ks / (h * ks + kd); cfm0 /= h;
    }
    V3 dir[3];
    dir[0] = nrm;
    plane_space(nrm, dir[1], dir[2]);
    const float s1 = arm ? -1.f : 1.f;
    const float inv_m = 1.f / T::OBJ_MASS;
    const int which = p == 3 ? 0 : p;
    float cfmr = 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const V3 d = dir[k];
      float den = 0.f, rel = 0.f;
      const V3 r = (s1 > 0.f ? pa : pb) - opos;
      const V3 rxd = cross(r, d), ir = Iinv * rxd;
      const V3 Jo = s1 * rxd, Vo = s1 * ir;
      den += 1.f / T::OBJ_MASS + dot(rxd, ir);
      rel += s1 * (dot(d, vu) + dot(rxd, wu));
      float J[N], V[N];
#pragma unroll
      for (int q = 0; q < N; q++) { J[q] = 0.f; V[q] = 0.f; }
      if (arm) {
        const V3 pxd = cross(pa, d);
#pragma unroll
        for (int q = 0; q < N; q++) {
          const bool on = q < 7 || (which == 1 && q == MD::F1) || (which == 2 && q == MD::F2);
          const float* S = dyn + 6 * q;
          J[q] = on ? 1.f * (dot(v3(S[0], S[1], S[2]), pxd) + dot(v3(S[3], S[4], S[5]), d)) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < N; q++) {
          float sacc = 0.f;
#pragma unroll
          for (int q2 = 0; q2 < N; q2++) sacc += Minv[tri(q, q2)] * J[q2];
          V[q] = sacc;
          den += J[q] * sacc;
          rel += J[q] * qdu[q];
        }
      }
      float rhs, dinv;
      if (k == 0) {
        dinv = 1.f / (den + cfm0);
        const float pen = -depth + (float)XARM_LINEAR_SLOP;
        float pos_err = 0.f, vel_err = -rel;
        if (pen > 0.f) vel_err -= pen / h; else pos_err = -pen * erp / h;
        rhs = (pos_err + vel_err) * dinv;
        cfmr = cfm0 * dinv;
      } else {
        dinv = 1.f / den;
        rhs = -rel * dinv;
      }
      float* ro = rec + R::HDR + (l * 3 + k) * F::CROW;
#pragma unroll
      for (int q = 0; q < N; q++) { Jr[k][q] = J[q]; ro[q] = V[q]; }
      Jr[k][9] = s1 * d.x; Jr[k][10] = s1 * d.y; Jr[k][11] = s1 * d.z; Jr[k][12] = Jo.x; Jr[k][13] = Jo.y; Jr[k][14] = Jo.z; Jr[k][15] = 0.f;
      ro[9] = s1 * inv_m * d.x; ro[10] = s1 * inv_m * d.y; ro[11] = s1 * inv_m * d.z; ro[12] = Vo.x; ro[13] = Vo.y; ro[14] = Vo.z; ro[15] = 0.f;
      ro[16] = rhs; ro[17] = dinv; ro[18] = cfmr; ro[19] = mu;
    }
  }
  const int grasp_now = (c1 > 0 && c2 > 0) ? 1 : 0;   // both fingers hold manifold points with the lego in this collision pass
  __syncwarp();
  FPROF()
  // ---- 5. the joint loop (it builds the Delassus matrix over the scratch above), stepPositionsMultiDof
  const float u = heavy_solve_dela<T, true>(rec, sm + F::DELA, l, valid, Jr);
  FPROF()
  heavy_integrate_coop<T>(a, t, valid, rec, u, l);
  FPROF()
#ifdef XARM_FUSED_PROF
  if (blockIdx.x == 0 && threadIdx.x == 0 && sub == 3)
    printf("[fused prof] nc %d | load %lld | collide %lld | header %lld | rows %lld | solve %lld | integrate %lld cycles\n", nc,
           pc_[1] - pc_[0], pc_[2] - pc_[1], pc_[3] - pc_[2], pc_[4] - pc_[3], pc_[5] - pc_[4], pc_[6] - pc_[5]);
#endif
  if (valid && sub == T::NSUB - 1 && l == 0 && MD::HAS_BOXES)
    a.state[(int64_t)(state_words<T>() - 2) * n + i] = grasp_word(grasp_now, grasp_cmd0);
  __syncwarp();
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
  heavy_integrate_coop<T>(a, t, valid, srec, u, l);
  __syncwarp();
}
// 16 lanes of the low-latency solve kernel: impulse-space joint loop straight from the record in global memory (L2)
template <class T>
__device__ __forceinline__ void heavy_solve2_body(const KArgs& a, int t, bool valid, const float* __restrict__ hrec, float* sm, int l) {
  using R = HeavyRec<T>;
  static_assert(T::MAXC <= XARM_DELA_MAXC && XARM_DELA_MAXC <= 16, "one lane per contact");
  const int tt = valid ? t : 0;
  const float* rec = hrec + (size_t)(a.heavy_dir > 0 ? tt : a.n - 1 - tt) * R::WORDS;
  const float u = heavy_solve_dela<T>(rec, sm, l, valid);
  heavy_integrate_coop<T>(a, t, valid, rec, u, l);
  __syncwarp();
}
#endif
