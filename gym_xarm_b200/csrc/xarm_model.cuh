// xarm_model.cuh - robot model tables in constant memory + compile-time topology of the two arm models.
// Numbers come from include/xarm_model_tables.h (baked from the reference URDFs by tools/bake_model.py).
#pragma once
#include "../../include/xarm_constants.h"
#include "../../include/xarm_model_tables.h"
#include "xarm_math.cuh"

// ---- xArm7 + Panda hand [REF gym_xarm/envs/urdf/xarm7_pd.urdf] ----
__constant__ float c_pd_R0[XARM_PD_NDOF][9] = XARM_PD_R0;
__constant__ float c_pd_t0[XARM_PD_NDOF][3] = XARM_PD_T0;
__constant__ float c_pd_axis[XARM_PD_NDOF][3] = XARM_PD_AXIS;
__constant__ float c_pd_lo[XARM_PD_NDOF] = XARM_PD_LIMIT_LO;
__constant__ float c_pd_hi[XARM_PD_NDOF] = XARM_PD_LIMIT_HI;
__constant__ float c_pd_damping[XARM_PD_NDOF] = XARM_PD_DAMPING;
__constant__ float c_pd_mass[XARM_PD_NDOF] = XARM_PD_MASS;
__constant__ float c_pd_com[XARM_PD_NDOF][3] = XARM_PD_COM;
__constant__ float c_pd_inertia[XARM_PD_NDOF][6] = XARM_PD_INERTIA;
__constant__ float c_pd_central[XARM_PD_NDOF][6] = XARM_PD_CENTRAL_INERTIA;
__constant__ int c_pd_part_owner[XARM_PD_NPART] = XARM_PD_PART_OWNER;
__constant__ float c_pd_part_mass[XARM_PD_NPART] = XARM_PD_PART_MASS;
__constant__ float c_pd_part_com[XARM_PD_NPART][3] = XARM_PD_PART_COM;
__constant__ float c_pd_hand_com[3] = XARM_PD_HAND_COM;
__constant__ float c_pd_f1c[3] = XARM_PD_FINGER1_BOX_C;
__constant__ float c_pd_f1h[3] = XARM_PD_FINGER1_BOX_H;
__constant__ float c_pd_f2c[3] = XARM_PD_FINGER2_BOX_C;
__constant__ float c_pd_f2h[3] = XARM_PD_FINGER2_BOX_H;
__constant__ float c_pd_hc[3] = XARM_PD_HAND_BOX_C;
__constant__ float c_pd_hh[3] = XARM_PD_HAND_BOX_H;

// ---- xArm7 + xArm gripper [REF gym_xarm/envs/urdf/xarm7.urdf] ----
__constant__ float c_xg_R0[XARM_XG_NDOF][9] = XARM_XG_R0;
__constant__ float c_xg_t0[XARM_XG_NDOF][3] = XARM_XG_T0;
__constant__ float c_xg_axis[XARM_XG_NDOF][3] = XARM_XG_AXIS;
__constant__ float c_xg_lo[XARM_XG_NDOF] = XARM_XG_LIMIT_LO;
__constant__ float c_xg_hi[XARM_XG_NDOF] = XARM_XG_LIMIT_HI;
__constant__ float c_xg_damping[XARM_XG_NDOF] = XARM_XG_DAMPING;
__constant__ float c_xg_mass[XARM_XG_NDOF] = XARM_XG_MASS;
__constant__ float c_xg_com[XARM_XG_NDOF][3] = XARM_XG_COM;
__constant__ float c_xg_inertia[XARM_XG_NDOF][6] = XARM_XG_INERTIA;
__constant__ float c_xg_central[XARM_XG_NDOF][6] = XARM_XG_CENTRAL_INERTIA;
__constant__ int c_xg_part_owner[XARM_XG_NPART] = XARM_XG_PART_OWNER;
__constant__ float c_xg_part_mass[XARM_XG_NPART] = XARM_XG_PART_MASS;
__constant__ float c_xg_part_com[XARM_XG_NPART][3] = XARM_XG_PART_COM;
__constant__ float c_xg_hand_com[3] = XARM_XG_HAND_COM;

__constant__ float c_joint_init[7] = {-0.009068751632859924, -0.08153217279952825, 0.09299669711139864, 1.067692645248743,
                                      0.0004018824370178429, 1.1524205092196147, -0.0004991403332530034};  // [REF xarm_reach.py:33]

#define XHD __host__ __device__ constexpr

struct ModelPD {
  static constexpr int N = XARM_PD_NDOF, NPART = XARM_PD_NPART, EEF = XARM_PD_EEF_DOF;
  static constexpr int F1 = XARM_PD_FINGER1_DOF, F2 = XARM_PD_FINGER2_DOF;
  static constexpr bool HAS_BOXES = true, HAS_GEAR = true;
  static XHD int parent(int i) { return i == 0 ? -1 : (i <= 6 ? i - 1 : 6); }
  static XHD bool prismatic(int i) { return i >= 7; }
  static XD const float* R0(int i) { return c_pd_R0[i]; }
  static XD V3 t0(int i) { return v3(c_pd_t0[i][0], c_pd_t0[i][1], c_pd_t0[i][2]); }
  static XD V3 axis(int i) { return v3(c_pd_axis[i][0], c_pd_axis[i][1], c_pd_axis[i][2]); }
  static XD float lo(int i) { return c_pd_lo[i]; }
  static XD float hi(int i) { return c_pd_hi[i]; }
  static XD float damping(int i) { return c_pd_damping[i]; }
  static XD float mass(int i) { return c_pd_mass[i]; }
  static XD V3 com(int i) { return v3(c_pd_com[i][0], c_pd_com[i][1], c_pd_com[i][2]); }
  static XD S3 inertia(int i) { S3 s = {c_pd_inertia[i][0], c_pd_inertia[i][1], c_pd_inertia[i][2], c_pd_inertia[i][3], c_pd_inertia[i][4], c_pd_inertia[i][5]}; return s; }
  static XD S3 central(int i) { S3 s = {c_pd_central[i][0], c_pd_central[i][1], c_pd_central[i][2], c_pd_central[i][3], c_pd_central[i][4], c_pd_central[i][5]}; return s; }
  static XHD int part_owner(int p) { return p <= 6 ? p : (p <= 8 ? 6 : p - 2); }  // == XARM_PD_PART_OWNER (checked on the host at load)
  static XD float part_mass(int p) { return c_pd_part_mass[p]; }
  static XD V3 part_com(int p) { return v3(c_pd_part_com[p][0], c_pd_part_com[p][1], c_pd_part_com[p][2]); }
  static XD V3 hand_com() { return v3(c_pd_hand_com[0], c_pd_hand_com[1], c_pd_hand_com[2]); }
};

struct ModelXG {
  static constexpr int N = XARM_XG_NDOF, NPART = XARM_XG_NPART, EEF = XARM_XG_EEF_DOF;
  static constexpr int F1 = XARM_XG_DRIVE_DOF, F2 = -1;
  static constexpr bool HAS_BOXES = false, HAS_GEAR = false;
  // XARM_XG_PARENT {-1,0,1,2,3,4,5,6,7,6,6,10,6}
  static XHD int parent(int i) { return i == 0 ? -1 : (i <= 7 ? i - 1 : (i == 8 ? 7 : (i == 11 ? 10 : 6))); }
  static XHD bool prismatic(int) { return false; }
  static XD const float* R0(int i) { return c_xg_R0[i]; }
  static XD V3 t0(int i) { return v3(c_xg_t0[i][0], c_xg_t0[i][1], c_xg_t0[i][2]); }
  static XD V3 axis(int i) { return v3(c_xg_axis[i][0], c_xg_axis[i][1], c_xg_axis[i][2]); }
  static XD float lo(int i) { return c_xg_lo[i]; }
  static XD float hi(int i) { return c_xg_hi[i]; }
  static XD float damping(int i) { return c_xg_damping[i]; }
  static XD float mass(int i) { return c_xg_mass[i]; }
  static XD V3 com(int i) { return v3(c_xg_com[i][0], c_xg_com[i][1], c_xg_com[i][2]); }
  static XD S3 inertia(int i) { S3 s = {c_xg_inertia[i][0], c_xg_inertia[i][1], c_xg_inertia[i][2], c_xg_inertia[i][3], c_xg_inertia[i][4], c_xg_inertia[i][5]}; return s; }
  static XD S3 central(int i) { S3 s = {c_xg_central[i][0], c_xg_central[i][1], c_xg_central[i][2], c_xg_central[i][3], c_xg_central[i][4], c_xg_central[i][5]}; return s; }
  // parts: link1..7, link_eef, gripper base, 6 gripper links, link_tcp -> owners {0..6,6,6,7,8,9,10,11,12,6}
  static XHD int part_owner(int p) { return p <= 6 ? p : (p <= 8 ? 6 : (p <= 14 ? p - 2 : 6)); }
  static XD float part_mass(int p) { return c_xg_part_mass[p]; }
  static XD V3 part_com(int p) { return v3(c_xg_part_com[p][0], c_xg_part_com[p][1], c_xg_part_com[p][2]); }
  static XD V3 hand_com() { return v3(c_xg_hand_com[0], c_xg_hand_com[1], c_xg_hand_com[2]); }
};

// j is i or an ancestor of i
template <class MD>
XHD bool is_anc(int j, int i) {
  while (i >= 0) {
    if (i == j) return true;
    i = MD::parent(i);
  }
  return false;
}
XHD int tri(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }
