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
#include "xarm_sim.cuh"

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
  const bool any_lim = (lim_lo[0] | lim_hi[0]) != 0u;
  for (int it = 0; it < XARM_SOLVER_ITERATIONS; it++) {
    bool resid_bad = false;
    if (it & 1) {
      if (any_lim) ARM_LIMITS_FWD(0)
      ARM_MOTORS_FWD(0) GEAR_ROW(0)
    } else {
      GEAR_ROW(0) ARM_MOTORS_BWD(0)
      if (any_lim) ARM_LIMITS_BWD(0)
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
#endif
