/* xarm_oracle.c - CPU ORACLE for the gym-xarm environment step.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load or call
 * this file.  The product (gym_xarm_b200/, libxarm_b200.so) never does and has no CPU fallback.
 *
 * What it is: a plain-C, double-precision, one-env-at-a-time restatement of what the reference env classes do per
 * step, INCLUDING the PyBullet subsystems they call (the arithmetic lives in the third-party, unpinned `pybullet`
 * wheel [REF setup.py:17], absent from /root/reference and not installable here).  The gym-level logic follows the
 * reference files line by line (cited per function).  The Bullet-level algorithms (ABA, joint motors/limits/gear as
 * velocity rows, box-box contacts, sequential-impulse PGS, DLS IK) are restated from Bullet3's published
 * algorithms as recalled in SURVEY.md Appendix B/I; every recalled default sits in include/xarm_constants.h.
 *
 * PARITY STATUS: "parity unpinned" for dynamics - the reference ships no golden vectors and PyBullet cannot run
 * here.  What IS pinned (tests/test_oracle_golden.py): G1 free-fall terminal velocity -15.1604595 m/s recovered from
 * the reference's saved VecNormalize pickle; FK known answers of the URDF; obs/goal dims (29 = pickle obs dim);
 * reward/success formulas against NumPy float32 evaluation of the reference's own expressions.
 *
 * Deliberately written differently from the CUDA path: dynamics here is Featherstone's articulated-body algorithm
 * on dense 6x6 spatial matrices in world coordinates and every constraint row is a dense vector over the whole
 * world's generalized velocity; the CUDA path uses CRBA/RNEA with an explicit inverse mass matrix and sparse rows.
 * Agreement between the two is therefore a real check, not an identity. */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/xarm_abi.h"
#include "../include/xarm_constants.h"
#include "../include/xarm_model_tables.h"

#define MAXDOF 13
#define MAXPART 16
#define MAXARM 2
#define MAXOBJ 3
#define MAXNV (MAXARM * MAXDOF + MAXOBJ * 6 + 1)
#define MAXCONTACT XARM_MAX_CONTACTS
#define MAXROW (MAXARM * (2 * MAXDOF + 1) + 4 + 3 * MAXCONTACT)
#define MAXCOLLIDER 24
#define MAXPAIR 64

/* ------------------------------------------------------------------------------------------------ small math */
typedef double v3[3];
typedef double m3[9]; /* row-major */

static void v3set(v3 a, double x, double y, double z) { a[0] = x; a[1] = y; a[2] = z; }
static void v3cpy(v3 a, const v3 b) { a[0] = b[0]; a[1] = b[1]; a[2] = b[2]; }
static double v3dot(const v3 a, const v3 b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void v3cross(v3 o, const v3 a, const v3 b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static void v3add(v3 o, const v3 a, const v3 b) { o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2]; }
static void v3sub(v3 o, const v3 a, const v3 b) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static void v3axpy(v3 o, double s, const v3 a) { o[0] += s * a[0]; o[1] += s * a[1]; o[2] += s * a[2]; }
static void v3scale(v3 o, double s, const v3 a) { o[0] = s * a[0]; o[1] = s * a[1]; o[2] = s * a[2]; }
static double v3norm(const v3 a) { return sqrt(v3dot(a, a)); }
static void m3mul(m3 o, const m3 a, const m3 b) {
  m3 t;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  memcpy(o, t, sizeof(m3));
}
static void m3vec(v3 o, const m3 a, const v3 x) {
  v3 t;
  for (int i = 0; i < 3; i++) t[i] = a[3 * i] * x[0] + a[3 * i + 1] * x[1] + a[3 * i + 2] * x[2];
  v3cpy(o, t);
}
static void m3tvec(v3 o, const m3 a, const v3 x) {
  v3 t;
  for (int i = 0; i < 3; i++) t[i] = a[i] * x[0] + a[3 + i] * x[1] + a[6 + i] * x[2];
  v3cpy(o, t);
}
static void m3ident(m3 a) { memset(a, 0, sizeof(m3)); a[0] = a[4] = a[8] = 1; }
/* rotation about a unit axis by angle (Rodrigues) */
static void m3axis_angle(m3 R, const v3 a, double th) {
  double c = cos(th), s = sin(th), t = 1 - c;
  R[0] = t * a[0] * a[0] + c;        R[1] = t * a[0] * a[1] - s * a[2]; R[2] = t * a[0] * a[2] + s * a[1];
  R[3] = t * a[0] * a[1] + s * a[2]; R[4] = t * a[1] * a[1] + c;        R[5] = t * a[1] * a[2] - s * a[0];
  R[6] = t * a[0] * a[2] - s * a[1]; R[7] = t * a[1] * a[2] + s * a[0]; R[8] = t * a[2] * a[2] + c;
}
/* quaternion xyzw (PyBullet order) */
static void quat_to_m3(m3 R, const double q[4]) {
  double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}
static void m3_to_quat(double q[4], const m3 R) { /* Shepperd, as btMatrix3x3::getRotation */
  double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    double s = sqrt(tr + 1.0);
    q[3] = 0.5 * s; s = 0.5 / s;
    q[0] = (R[7] - R[5]) * s; q[1] = (R[2] - R[6]) * s; q[2] = (R[3] - R[1]) * s;
  } else {
    int i = R[0] < R[4] ? (R[4] < R[8] ? 2 : 1) : (R[0] < R[8] ? 2 : 0);
    int j = (i + 1) % 3, k = (i + 2) % 3;
    double s = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[i] = 0.5 * s; s = 0.5 / s;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * s;
    q[j] = (R[3 * j + i] + R[3 * i + j]) * s;
    q[k] = (R[3 * k + i] + R[3 * i + k]) * s;
  }
}
static void quat_mul(double o[4], const double a[4], const double b[4]) {
  double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  double y = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
  double z = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
  double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}

/* ------------------------------------------------------------------------------------------------ robot model */
typedef struct {
  int ndof, npart;
  int parent[MAXDOF], jtype[MAXDOF];
  double R0[MAXDOF][9], t0[MAXDOF][3], axis[MAXDOF][3], lo[MAXDOF], hi[MAXDOF], damping[MAXDOF];
  int part_owner[MAXPART];
  double part_mass[MAXPART], part_com[MAXPART][3], part_inertia[MAXPART][6];
  int eef_dof;
  double hand_com[3];
  int finger1, finger2; /* PD: finger DoFs; XG: finger1 = drive joint, finger2 = -1 */
  int has_gripper_boxes;
  double f1c[3], f1h[3], f2c[3], f2h[3], hc[3], hh[3];
} Model;

static Model g_pd, g_xg;
static int g_models_ready = 0;

static void model_init(void) {
  if (g_models_ready) return;
  {
    static const int parent[] = XARM_PD_PARENT, jt[] = XARM_PD_JTYPE, po[] = XARM_PD_PART_OWNER;
    static const double R0[][9] = XARM_PD_R0, t0[][3] = XARM_PD_T0, ax[][3] = XARM_PD_AXIS, lo[] = XARM_PD_LIMIT_LO,
                        hi[] = XARM_PD_LIMIT_HI, dm[] = XARM_PD_DAMPING, pm[] = XARM_PD_PART_MASS,
                        pc[][3] = XARM_PD_PART_COM, pi[][6] = XARM_PD_PART_INERTIA, hc[] = XARM_PD_HAND_COM;
    static const double f1c[] = XARM_PD_FINGER1_BOX_C, f1h[] = XARM_PD_FINGER1_BOX_H, f2c[] = XARM_PD_FINGER2_BOX_C,
                        f2h[] = XARM_PD_FINGER2_BOX_H, bc[] = XARM_PD_HAND_BOX_C, bh[] = XARM_PD_HAND_BOX_H;
    Model* m = &g_pd;
    m->ndof = XARM_PD_NDOF; m->npart = XARM_PD_NPART;
    for (int i = 0; i < m->ndof; i++) {
      m->parent[i] = parent[i]; m->jtype[i] = jt[i];
      memcpy(m->R0[i], R0[i], sizeof(m3)); memcpy(m->t0[i], t0[i], sizeof(v3)); memcpy(m->axis[i], ax[i], sizeof(v3));
      m->lo[i] = lo[i]; m->hi[i] = hi[i]; m->damping[i] = dm[i];
    }
    for (int i = 0; i < m->npart; i++) {
      m->part_owner[i] = po[i]; m->part_mass[i] = pm[i];
      memcpy(m->part_com[i], pc[i], sizeof(v3)); memcpy(m->part_inertia[i], pi[i], 6 * sizeof(double));
    }
    m->eef_dof = XARM_PD_EEF_DOF; memcpy(m->hand_com, hc, sizeof(v3));
    m->finger1 = XARM_PD_FINGER1_DOF; m->finger2 = XARM_PD_FINGER2_DOF; m->has_gripper_boxes = 1;
    memcpy(m->f1c, f1c, sizeof(v3)); memcpy(m->f1h, f1h, sizeof(v3)); memcpy(m->f2c, f2c, sizeof(v3));
    memcpy(m->f2h, f2h, sizeof(v3)); memcpy(m->hc, bc, sizeof(v3)); memcpy(m->hh, bh, sizeof(v3));
  }
  {
    static const int parent[] = XARM_XG_PARENT, jt[] = XARM_XG_JTYPE, po[] = XARM_XG_PART_OWNER;
    static const double R0[][9] = XARM_XG_R0, t0[][3] = XARM_XG_T0, ax[][3] = XARM_XG_AXIS, lo[] = XARM_XG_LIMIT_LO,
                        hi[] = XARM_XG_LIMIT_HI, dm[] = XARM_XG_DAMPING, pm[] = XARM_XG_PART_MASS,
                        pc[][3] = XARM_XG_PART_COM, pi[][6] = XARM_XG_PART_INERTIA, hc[] = XARM_XG_HAND_COM;
    Model* m = &g_xg;
    m->ndof = XARM_XG_NDOF; m->npart = XARM_XG_NPART;
    for (int i = 0; i < m->ndof; i++) {
      m->parent[i] = parent[i]; m->jtype[i] = jt[i];
      memcpy(m->R0[i], R0[i], sizeof(m3)); memcpy(m->t0[i], t0[i], sizeof(v3)); memcpy(m->axis[i], ax[i], sizeof(v3));
      m->lo[i] = lo[i]; m->hi[i] = hi[i]; m->damping[i] = dm[i];
    }
    for (int i = 0; i < m->npart; i++) {
      m->part_owner[i] = po[i]; m->part_mass[i] = pm[i];
      memcpy(m->part_com[i], pc[i], sizeof(v3)); memcpy(m->part_inertia[i], pi[i], 6 * sizeof(double));
    }
    m->eef_dof = XARM_XG_EEF_DOF; memcpy(m->hand_com, hc, sizeof(v3));
    m->finger1 = XARM_XG_DRIVE_DOF; m->finger2 = -1; m->has_gripper_boxes = 0;
  }
  g_models_ready = 1;
}

/* ------------------------------------------------------------------------------------------------ task table
 * Per-task constants: SURVEY.md Appendix A, each from the reference constructor cited there. */
typedef struct {
  int task, n_arms, n_obj, has_door, n_tables, has_ground;
  const Model* model;
  double base_pos[MAXARM][3], base_yaw[MAXARM];
  double time_step;     /* p.setTimeStep */
  double h;             /* internal substep */
  int n_sub;            /* substeps per env step */
  int damping_each_sub; /* Handover calls stepSimulation 15x => joint damping torque applied every substep */
  double dt_cmd, max_vel, max_grip_vel;
  int n_ik;
  float pos_lo[MAXARM][3], pos_hi[MAXARM][3];
  int grip_clip, grip_cmd;
  float grip_lo, grip_hi;
  double arm_force, finger_force;
  int gear, friction_switch, lego_clamp;
  double obj_half[3], obj_mass;
  double table_c[2][3];
  float threshold;
  int max_steps;
  double hand_offset[3]; /* subtracted from the hand position in obs (Handover eef2grip_offset) */
  int act_dim, obs_dim, goal_dim;
} Task;

static const double JOINT_INIT[8] = {0, -0.009068751632859924, -0.08153217279952825, 0.09299669711139864,
                                     1.067692645248743, 0.0004018824370178429, 1.1524205092196147,
                                     -0.0004991403332530034}; /* [REF xarm_reach.py:33] index = PyBullet joint 0..7 */

static int task_fill(Task* t, int task, int num_obj) {
  model_init();
  memset(t, 0, sizeof(*t));
  t->task = task;
  t->model = &g_pd;
  t->n_arms = 1;
  t->n_tables = 1;
  t->grip_cmd = 1;
  t->arm_force = XARM_MOTOR_DEFAULT_FORCE;
  t->finger_force = XARM_MOTOR_DEFAULT_FORCE;
  t->gear = 1;
  switch (task) {
    case XARM_TASK_REACH: /* [REF xarm_reach.py:15-35] */
      t->model = &g_xg; t->gear = 0;
      t->time_step = 1. / 240; t->n_sub = 20; t->h = t->time_step / 20; t->dt_cmd = t->time_step * 20;
      t->max_vel = 1; t->max_grip_vel = 20; t->n_ik = 20;
      t->pos_lo[0][0] = 0.2f; t->pos_lo[0][1] = -0.4f; t->pos_lo[0][2] = 0.2f;
      t->pos_hi[0][0] = 0.8f; t->pos_hi[0][1] = 0.4f; t->pos_hi[0][2] = 0.6f;
      t->grip_clip = 0; t->arm_force = 5 * 240.; t->finger_force = 5 * 240.;
      t->threshold = 0.05f; t->max_steps = 25; t->n_obj = 0;
      t->act_dim = 4; t->obs_dim = 8; t->goal_dim = 3;
      break;
    case XARM_TASK_PICK_AND_PLACE: /* [REF xarm_pick_and_place.py:17-50] */
      t->time_step = 1. / 60; t->n_sub = 15; t->h = t->time_step / 15; t->dt_cmd = t->time_step * 15;
      t->max_vel = 0.25; t->max_grip_vel = 0.08; t->n_ik = 15;
      t->pos_lo[0][0] = 0.3f; t->pos_lo[0][1] = -0.3f; t->pos_lo[0][2] = 0.15f;
      t->pos_hi[0][0] = 0.5f; t->pos_hi[0][1] = 0.3f; t->pos_hi[0][2] = 0.4f;
      t->grip_clip = 1; t->grip_lo = 0.01f; t->grip_hi = 0.04f; t->finger_force = 1000;
      t->friction_switch = 1;
      t->n_obj = num_obj; t->obj_half[0] = 0.05 / 2; t->obj_half[1] = 0.025; t->obj_half[2] = 0.04; t->obj_mass = 0.5;
      t->threshold = 0.05f; t->max_steps = 50;
      t->act_dim = 4; t->obs_dim = 8 + 16 * num_obj; t->goal_dim = 3 * num_obj;
      break;
    case XARM_TASK_STACK_TOWER:   /* [REF xarm_stack_tower.py:14-43] */
    case XARM_TASK_PUSH_WITH_DOOR: /* [REF xarm_push_with_door.py:14-42] */
      t->n_arms = 2;
      t->base_pos[0][0] = -0.6; t->base_pos[1][0] = 0.6; t->base_yaw[1] = M_PI;
      t->time_step = 1. / 60; t->n_sub = 15; t->h = t->time_step / 15; t->dt_cmd = t->time_step * 15;
      t->max_vel = 0.25; t->max_grip_vel = 1; t->n_ik = 15;
      t->pos_lo[0][0] = -0.4f; t->pos_lo[0][1] = -0.3f; t->pos_lo[0][2] = 0.125f;
      t->pos_hi[0][0] = 0.3f; t->pos_hi[0][1] = 0.3f; t->pos_hi[0][2] = 0.4f;
      t->pos_lo[1][0] = -0.3f; t->pos_lo[1][1] = -0.3f; t->pos_lo[1][2] = 0.125f;
      t->pos_hi[1][0] = 0.4f; t->pos_hi[1][1] = 0.3f; t->pos_hi[1][2] = 0.4f;
      t->grip_clip = 1; t->grip_lo = 0.021f; t->grip_hi = 0.04f;
      t->obj_half[0] = t->obj_half[1] = t->obj_half[2] = 0.025; t->obj_mass = 0.1;
      t->max_steps = 50;
      if (task == XARM_TASK_STACK_TOWER) {
        t->n_obj = 3; t->threshold = (float)(0.03 * 3); /* [REF xarm_stack_tower.py:20-21] python double 0.09 -> compare in f32 */
        t->act_dim = 8; t->obs_dim = 13 * 3 + 16; t->goal_dim = 9;
      } else {
        t->n_obj = 1; t->threshold = (float)(0.03 * 1); t->has_door = 1; t->grip_cmd = 0; /* D1: no finger command */
        t->act_dim = 6; t->obs_dim = 25; t->goal_dim = 3;
      }
      break;
    case XARM_TASK_HANDOVER: /* [REF xarm_handover.py:25-57,79-83] */
      t->n_arms = 2;
      t->base_pos[0][0] = -0.6; t->base_pos[1][0] = 0.6; t->base_yaw[1] = M_PI;
      t->time_step = 1. / 240; t->n_sub = 15; t->h = t->time_step; t->dt_cmd = t->time_step * 15; t->damping_each_sub = 1;
      t->max_vel = 1.8; t->max_grip_vel = 1; t->n_ik = 15;
      t->pos_lo[0][0] = -0.3f; t->pos_lo[0][1] = -0.2f; t->pos_lo[0][2] = 0.1f;
      t->pos_hi[0][0] = 0.0f; t->pos_hi[0][1] = 0.2f; t->pos_hi[0][2] = 0.22f;
      t->pos_lo[1][0] = 0.0f; t->pos_lo[1][1] = -0.2f; t->pos_lo[1][2] = 0.1f;
      t->pos_hi[1][0] = 0.3f; t->pos_hi[1][1] = 0.2f; t->pos_hi[1][2] = 0.22f;
      t->grip_clip = 1; t->grip_lo = 0.020f; t->grip_hi = 0.04f;
      t->friction_switch = 1; t->lego_clamp = 1;
      t->n_obj = num_obj; t->obj_half[0] = 0.15 / 2; t->obj_half[1] = 0.025; t->obj_half[2] = 0.025; t->obj_mass = 0.5;
      t->n_tables = 2; t->table_c[0][0] = -0.85; t->table_c[1][0] = 0.85; t->has_ground = 1;
      t->threshold = 0.05f; t->max_steps = 100;
      t->hand_offset[2] = 0.088 - 0.021; /* eef2grip_offset [REF xarm_handover.py:49] */
      t->act_dim = 8; t->obs_dim = 13 * num_obj + 16; t->goal_dim = 3 * num_obj;
      break;
    default: return -1;
  }
  if (t->n_obj > MAXOBJ || t->n_obj < 0) return -1;
  for (int k = 0; k < t->n_tables; k++) t->table_c[k][2] = -XARM_TABLE_HALF_Z;
  return 0;
}

static int state_words(const Task* t) {
  return t->n_arms * 3 * t->model->ndof + t->n_obj * 13 + (t->has_door ? 2 : 0) + t->goal_dim + 5;
}

/* ------------------------------------------------------------------------------------------------ env state */
typedef struct {
  double q[MAXDOF], qd[MAXDOF], qt[MAXDOF]; /* position, velocity, motor position target */
} ArmState;
typedef struct {
  double pos[3], quat[4], v[3], w[3];
} ObjState;

typedef struct OrEnv {
  Task t;
  XarmConfig cfg;
  int64_t env_index;
  ArmState arm[MAXARM];
  ObjState obj[MAXOBJ];
  double door_q, door_qd;
  float goal[3 * MAXOBJ];
  int step_count;
  uint32_t episode;
  float d_old;
  int grasp[MAXARM];     /* both fingers of arm a hold manifold points with lego 0 after the LAST collision pass */
  int grasp_cmd[MAXARM]; /* those flags as _set_action read them: the friction switch and Handover's if_xarm*_grasp
                            [REF xarm_pick_and_place.py:212-218, xarm_handover.py:262-280]; they persist through reset() */
  uint32_t rng_draw; /* draws consumed in the current episode */
  double flops;      /* instrumented flop counter (FMA=2) over PGS/ABA inner loops: see or_flops() */
  long st_substeps, st_iters, st_rows, st_contacts; /* solver statistics (or_solver_stats) */
  long st_hist_nc[XARM_MAX_CONTACTS + 1], st_hist_nac[XARM_MAX_ARM_CONTACTS + 1];
  long arm_contacts; /* contact points that involved a gripper link since the last or_arm_contacts() (test aid) */
} OrEnv;

/* ------------------------------------------------------------------------------------------------ RNG (Appendix E) */
static void philox4x32_10(uint32_t ctr[4], const uint32_t key_in[2]) {
  uint32_t k0 = key_in[0], k1 = key_in[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)XARM_PHILOX_M0 * ctr[0], p1 = (uint64_t)XARM_PHILOX_M1 * ctr[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ ctr[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ ctr[3] ^ k1,
             n3 = (uint32_t)p0;
    ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
    k0 += XARM_PHILOX_W0; k1 += XARM_PHILOX_W1;
  }
}
/* uniform in [0,1): draw index k of (seed, global env, episode) */
static double rng_uniform(OrEnv* e) {
  uint32_t k = e->rng_draw++;
  uint64_t gi = (uint64_t)e->env_index;
  uint32_t ctr[4] = {(uint32_t)gi, (uint32_t)(gi >> 32), e->episode, k >> 2};
  uint32_t key[2] = {(uint32_t)e->cfg.seed, (uint32_t)(e->cfg.seed >> 32)};
  philox4x32_10(ctr, key);
  return (double)ctr[k & 3] * (1.0 / 4294967296.0);
}
/* gym Box.sample() for a bounded float32 box: float64 arithmetic on the float32-rounded bounds, cast to float32 */
static float box_sample(OrEnv* e, float lo, float hi) {
  double u = rng_uniform(e);
  return (float)((double)lo + ((double)hi - (double)lo) * u);
}

/* ------------------------------------------------------------------------------------------------ kinematics */
typedef struct {
  m3 R[MAXDOF];  /* world rotation of each moving link frame */
  v3 p[MAXDOF];  /* world origin of each moving link frame */
  v3 a[MAXDOF];  /* world joint axis */
  double S[MAXDOF][6]; /* spatial motion subspace about the world origin: [angular; linear] */
} Kin;

static void base_transform(const Task* t, int arm, m3 R, v3 p) {
  v3 z = {0, 0, 1};
  m3axis_angle(R, z, t->base_yaw[arm]);
  v3cpy(p, t->base_pos[arm]);
}

/* forward kinematics of one arm: N10 getLinkState / URDF joint origins (Appendix C) */
static void arm_fk(const Task* t, int arm, const double* q, Kin* k) {
  const Model* m = t->model;
  m3 Rb; v3 pb;
  base_transform(t, arm, Rb, pb);
  for (int i = 0; i < m->ndof; i++) {
    const double* Rp = m->parent[i] < 0 ? Rb : k->R[m->parent[i]];
    const double* pp = m->parent[i] < 0 ? pb : k->p[m->parent[i]];
    m3 Rj, Rq; v3 tj;
    m3mul(Rj, Rp, m->R0[i]);
    m3vec(tj, Rp, m->t0[i]);
    v3add(tj, tj, pp);
    m3vec(k->a[i], Rj, m->axis[i]);
    if (m->jtype[i] == 0) {
      m3axis_angle(Rq, m->axis[i], q[i]);
      m3mul(k->R[i], Rj, Rq);
      v3cpy(k->p[i], tj);
      v3cpy(&k->S[i][0], k->a[i]);
      v3cross(&k->S[i][3], k->p[i], k->a[i]);
    } else {
      memcpy(k->R[i], Rj, sizeof(m3));
      v3cpy(k->p[i], tj);
      v3axpy(k->p[i], q[i], k->a[i]);
      v3set(&k->S[i][0], 0, 0, 0);
      v3cpy(&k->S[i][3], k->a[i]);
    }
  }
}
static void link_point(const Kin* k, int link, const v3 local, v3 out) {
  m3vec(out, k->R[link], local);
  v3add(out, out, k->p[link]);
}
/* world velocity of a world point rigidly attached to `link` */
static void link_point_vel(const Model* m, const Kin* k, int link, const double* qd, const v3 pw, v3 out) {
  v3set(out, 0, 0, 0);
  for (int j = link; j >= 0; j = m->parent[j]) {
    v3 t;
    v3cross(t, &k->S[j][0], pw);
    v3add(t, t, &k->S[j][3]);
    v3axpy(out, qd[j], t);
  }
}

/* calculateInverseKinematics(body, link 8, pos, [1,0,0,0], maxNumIterations=n) - SURVEY B.3 (N9):
 * iterate dq = (J^T J + 0.5 I)^-1 J^T e with e = [p*-p ; angle*axis(q* q^-1)], clamp max |dq| to 45 deg.
 * Finger/gripper columns of J are zero for link 8, so the solve is 7x7 and those joints are returned unchanged. */
static void arm_ik(const Task* t, int arm, const double* q_in, const v3 target, double* q_out) {
  const Model* m = t->model;
  const int n = 7;
  double q[MAXDOF];
  memcpy(q, q_in, sizeof(double) * m->ndof);
  const double qT[4] = {1, 0, 0, 0}; /* target orientation xyzw [REF xarm_pick_and_place.py:207] */
  for (int it = 0; it < t->n_ik; it++) {
    Kin k;
    arm_fk(t, arm, q, &k);
    const int L = m->eef_dof;
    const double* pe = k.p[L];
    double e[6];
    for (int c = 0; c < 3; c++) e[c] = target[c] - pe[c];
    double qc[4], qci[4], dq[4];
    m3_to_quat(qc, k.R[L]);
    qci[0] = -qc[0]; qci[1] = -qc[1]; qci[2] = -qc[2]; qci[3] = qc[3];
    quat_mul(dq, qT, qci);
    double w = dq[3] > 1 ? 1 : (dq[3] < -1 ? -1 : dq[3]);
    double angle = 2 * acos(w);
    double s2 = 1 - w * w;
    v3 axis = {1, 0, 0};
    if (s2 >= 10 * 2.220446049250313e-16) { double s = 1 / sqrt(s2); v3set(axis, dq[0] * s, dq[1] * s, dq[2] * s); }
    if (angle > M_PI) angle -= 2 * M_PI;
    double an = v3norm(axis);
    for (int c = 0; c < 3; c++) e[3 + c] = angle * axis[c] / an;
    double J[6][7];
    for (int j = 0; j < n; j++) {
      v3 r, lin;
      v3sub(r, pe, k.p[j]);
      v3cross(lin, k.a[j], r);
      for (int c = 0; c < 3; c++) { J[c][j] = lin[c]; J[3 + c][j] = k.a[j][c]; }
    }
    double A[7][8];
    for (int i = 0; i < n; i++) {
      for (int j = 0; j < n; j++) {
        double s = 0;
        for (int r = 0; r < 6; r++) s += J[r][i] * J[r][j];
        A[i][j] = s + (i == j ? XARM_IK_DAMPING : 0.0);
      }
      double s = 0;
      for (int r = 0; r < 6; r++) s += J[r][i] * e[r];
      A[i][7] = s;
    }
    for (int c = 0; c < n; c++) { /* Gaussian elimination with partial pivoting (MatrixRmn::Solve) */
      int piv = c;
      for (int r = c + 1; r < n; r++) if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
      if (piv != c) for (int j = 0; j <= n; j++) { double tmp = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = tmp; }
      for (int r = c + 1; r < n; r++) {
        double f = A[r][c] / A[c][c];
        for (int j = c; j <= n; j++) A[r][j] -= f * A[c][j];
      }
    }
    double d[7];
    for (int i = n - 1; i >= 0; i--) {
      double s = A[i][7];
      for (int j = i + 1; j < n; j++) s -= A[i][j] * d[j];
      d[i] = s / A[i][i];
    }
    double mx = 0;
    for (int i = 0; i < n; i++) if (fabs(d[i]) > mx) mx = fabs(d[i]);
    if (mx > XARM_IK_MAX_STEP) for (int i = 0; i < n; i++) d[i] *= XARM_IK_MAX_STEP / mx;
    for (int i = 0; i < n; i++) q[i] += d[i];
    /* residual check on the position error after the update */
    arm_fk(t, arm, q, &k);
    v3 r;
    v3sub(r, target, k.p[L]);
    if (v3norm(r) < XARM_IK_RESIDUAL) break;
  }
  memcpy(q_out, q, sizeof(double) * m->ndof);
}

/* ------------------------------------------------------------------------------------------------ dynamics (ABA)
 * Featherstone articulated-body algorithm (N3) with every spatial quantity expressed in world
 * coordinates about the world origin, so no link-to-link transforms are needed.  Spatial motion = [w; v_O],
 * spatial force = [n_O; f]. */
typedef double sv[6];
typedef double sm[36];

static void part_spatial_inertia(sm I, double mass, const v3 c, const m3 Ic) {
  /* [[Ic + m cx cx^T, m cx],[m cx^T, m 1]] */
  double cx[9] = {0, -c[2], c[1], c[2], 0, -c[0], -c[1], c[0], 0};
  memset(I, 0, sizeof(sm));
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += cx[3 * i + k] * cx[3 * j + k];
      I[6 * i + j] = Ic[3 * i + j] + mass * s;
      I[6 * i + 3 + j] = mass * cx[3 * i + j];
      I[6 * (3 + i) + j] = mass * cx[3 * j + i];
    }
  for (int i = 0; i < 3; i++) I[6 * (3 + i) + 3 + i] = mass;
}
static void sm_vec(sv o, const sm A, const sv x) {
  sv t;
  for (int i = 0; i < 6; i++) { double s = 0; for (int j = 0; j < 6; j++) s += A[6 * i + j] * x[j]; t[i] = s; }
  memcpy(o, t, sizeof(sv));
}
static double sv_dot(const sv a, const sv b) { double s = 0; for (int i = 0; i < 6; i++) s += a[i] * b[i]; return s; }
static void motion_cross(sv o, const sv v, const sv m) { /* v x m */
  v3 a, b, c;
  v3cross(a, &v[0], &m[0]);
  v3cross(b, &v[0], &m[3]);
  v3cross(c, &v[3], &m[0]);
  v3cpy(&o[0], a); v3add(&o[3], b, c);
}
static void force_cross(sv o, const sv v, const sv f) { /* v x* f */
  v3 a, b, c;
  v3cross(a, &v[0], &f[0]);
  v3cross(b, &v[3], &f[3]);
  v3cross(c, &v[0], &f[3]);
  v3add(&o[0], a, b); v3cpy(&o[3], c);
}
static void sym6_to_m3(m3 I, const double s[6]) {
  I[0] = s[0]; I[1] = s[1]; I[2] = s[2]; I[3] = s[1]; I[4] = s[3]; I[5] = s[4]; I[6] = s[2]; I[7] = s[4]; I[8] = s[5];
}

/* One ABA evaluation.  with_bias=0 gives the pure unit-torque response (v=0, no gravity, no damping) used to build
 * the columns of M^-1 (Bullet: calcAccelerationDeltasMultiDof). */
static void arm_aba(OrEnv* e, const Task* t, const Kin* k, const double* qd, const double* tau, int with_bias, double* qdd) {
  const Model* m = t->model;
  const int n = m->ndof;
  sv v[MAXDOF], c[MAXDOF], pA[MAXDOF], U[MAXDOF], a[MAXDOF];
  sm IA[MAXDOF];
  double D[MAXDOF], u[MAXDOF];
  for (int i = 0; i < n; i++) {
    memset(IA[i], 0, sizeof(sm));
    memset(pA[i], 0, sizeof(sv));
    memset(v[i], 0, sizeof(sv));
    memset(c[i], 0, sizeof(sv));
    if (with_bias) {
      sv vj;
      for (int r = 0; r < 6; r++) vj[r] = k->S[i][r] * qd[i];
      if (m->parent[i] >= 0) memcpy(v[i], v[m->parent[i]], sizeof(sv));
      for (int r = 0; r < 6; r++) v[i][r] += vj[r];
      motion_cross(c[i], v[i], vj);
    }
  }
  for (int p = 0; p < m->npart; p++) {
    int i = m->part_owner[p];
    v3 cw; m3 Il, Iw, tmp, Rt;
    link_point(k, i, m->part_com[p], cw);
    sym6_to_m3(Il, m->part_inertia[p]);
    m3mul(tmp, k->R[i], Il);
    for (int r = 0; r < 3; r++) for (int s = 0; s < 3; s++) Rt[3 * r + s] = k->R[i][3 * s + r];
    m3mul(Iw, tmp, Rt);
    sm Ip;
    part_spatial_inertia(Ip, m->part_mass[p], cw, Iw);
    for (int r = 0; r < 36; r++) IA[i][r] += Ip[r];
    if (with_bias) {
      sv Iv, b;
      sm_vec(Iv, Ip, v[i]);
      force_cross(b, v[i], Iv);
      const double* w = &v[i][0];
      v3 vc, t3, Iw_w;
      v3cross(vc, w, cw); v3add(vc, vc, &v[i][3]); /* COM velocity */
      m3vec(Iw_w, Iw, w);
      if (!XARM_MB_USE_GYRO) { v3cross(t3, w, Iw_w); for (int r = 0; r < 3; r++) b[r] -= t3[r]; }
      /* external: gravity + Bullet multibody damping m v (k+k|v|), I w (k+k|w|) */
      v3 f, nC;
      double kl = XARM_MB_LINEAR_DAMPING * (1 + v3norm(vc)), ka = XARM_MB_ANGULAR_DAMPING * (1 + v3norm(w));
      for (int r = 0; r < 3; r++) { f[r] = -m->part_mass[p] * vc[r] * kl; nC[r] = -Iw_w[r] * ka; }
      f[2] -= m->part_mass[p] * XARM_GRAVITY;
      v3cross(t3, cw, f);
      for (int r = 0; r < 3; r++) { b[r] -= nC[r] + t3[r]; b[3 + r] -= f[r]; }
      for (int r = 0; r < 6; r++) pA[i][r] += b[r];
      e->flops += 250;
    }
  }
  for (int i = n - 1; i >= 0; i--) {
    sm_vec(U[i], IA[i], k->S[i]);
    D[i] = sv_dot(k->S[i], U[i]);
    u[i] = tau[i] - sv_dot(k->S[i], pA[i]);
    int p = m->parent[i];
    if (p >= 0) {
      sm Ia;
      for (int r = 0; r < 6; r++) for (int s = 0; s < 6; s++) Ia[6 * r + s] = IA[i][6 * r + s] - U[i][r] * U[i][s] / D[i];
      sv pa, Ic;
      sm_vec(Ic, Ia, c[i]);
      for (int r = 0; r < 6; r++) pa[r] = pA[i][r] + Ic[r] + U[i][r] * u[i] / D[i];
      for (int r = 0; r < 36; r++) IA[p][r] += Ia[r];
      for (int r = 0; r < 6; r++) pA[p][r] += pa[r];
    }
    e->flops += 72 + 12 + 12 + 108 + 72 + 18 + 42;
  }
  for (int i = 0; i < n; i++) {
    sv ap;
    int p = m->parent[i];
    if (p >= 0) memcpy(ap, a[p], sizeof(sv)); else memset(ap, 0, sizeof(sv));
    for (int r = 0; r < 6; r++) ap[r] += c[i][r];
    qdd[i] = (u[i] - sv_dot(U[i], ap)) / D[i];
    for (int r = 0; r < 6; r++) a[i][r] = ap[r] + k->S[i][r] * qdd[i];
    e->flops += 32;
  }
}

/* ------------------------------------------------------------------------------------------------ collision (N7)
 * Every collider is an oriented box: table tops / ground (static), free boxes, door bars, Panda finger and hand
 * hulls approximated by their AABBs.  box_box() = 15-axis SAT (face axes favoured by the 1.05
 * fudge factor), reference-face clipping of the incident face (<= 4 points kept), edge-edge closest points: the
 * published btBoxBoxDetector / ODE dBoxBox2 scheme restated, not copied. */
enum { BODY_STATIC = 0, BODY_ARM = 1, BODY_OBJ = 2, BODY_DOOR = 3 };
typedef struct {
  int body, index, link; /* BODY_ARM: index=arm, link=dof; BODY_OBJ: index=obj */
  v3 c; m3 R; v3 h;      /* world centre, rotation (columns = box axes), half extents */
  double friction;
  double stiffness, damping; /* soft contact when stiffness > 0 (Panda fingers) */
  double erp;                /* < 0: default erp2 */
} Collider;
typedef struct {
  int ca, cb;      /* collider indices: normal points from B to A */
  v3 pa, pb, n;    /* world points on A and B, unit normal */
  double depth;    /* penetration depth (>0 overlapping) */
} Contact;

static void col_axis(const Collider* b, int i, v3 out) { v3set(out, b->R[i], b->R[3 + i], b->R[6 + i]); }

static int clip_poly(double (*in)[2], int n, int axis, double sign, double lim, double (*out)[2]) {
  /* keep the half plane sign*x[axis] <= lim (Sutherland-Hodgman) */
  int m = 0;
  for (int i = 0; i < n; i++) {
    double* a = in[i]; double* b = in[(i + 1) % n];
    double da = sign * a[axis] - lim, db = sign * b[axis] - lim;
    if (da <= 0) { out[m][0] = a[0]; out[m][1] = a[1]; m++; }
    if ((da < 0 && db > 0) || (da > 0 && db < 0)) {
      double s = da / (da - db);
      out[m][0] = a[0] + s * (b[0] - a[0]); out[m][1] = a[1] + s * (b[1] - a[1]); m++;
    }
    if (m >= 8) break;
  }
  return m;
}

static int box_box(const Collider* A, const Collider* B, Contact* out, int max_out) {
  v3 Aa[3], Ba[3], T;
  for (int i = 0; i < 3; i++) { col_axis(A, i, Aa[i]); col_axis(B, i, Ba[i]); }
  v3sub(T, B->c, A->c);
  double Rm[3][3], Q[3][3], Ta[3];
  for (int i = 0; i < 3; i++) {
    Ta[i] = v3dot(T, Aa[i]);
    for (int j = 0; j < 3; j++) { Rm[i][j] = v3dot(Aa[i], Ba[j]); Q[i][j] = fabs(Rm[i][j]); }
  }
  double best = -1e30; /* largest (least negative) separation = -depth */
  int code = -1; double nsign = 1; v3 naxis = {0, 0, 0};
  /* face axes of A */
  for (int i = 0; i < 3; i++) {
    double s = fabs(Ta[i]) - (A->h[i] + B->h[0] * Q[i][0] + B->h[1] * Q[i][1] + B->h[2] * Q[i][2]);
    if (s > XARM_CONTACT_MARGIN) return 0;
    if (s > best) { best = s; code = i; nsign = Ta[i] < 0 ? -1 : 1; v3cpy(naxis, Aa[i]); }
  }
  /* face axes of B */
  for (int j = 0; j < 3; j++) {
    double tb = v3dot(T, Ba[j]);
    double s = fabs(tb) - (B->h[j] + A->h[0] * Q[0][j] + A->h[1] * Q[1][j] + A->h[2] * Q[2][j]);
    if (s > XARM_CONTACT_MARGIN) return 0;
    if (s > best) { best = s; code = 3 + j; nsign = tb < 0 ? -1 : 1; v3cpy(naxis, Ba[j]); }
  }
  /* edge x edge axes */
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      v3 ax; v3cross(ax, Aa[i], Ba[j]);
      double l = v3norm(ax);
      if (l < 1e-6) continue;
      int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      double tp = v3dot(T, ax);
      double ra = A->h[i1] * Q[i2][j] + A->h[i2] * Q[i1][j];
      double rb = B->h[j1] * Q[i][j2] + B->h[j2] * Q[i][j1];
      double s = (fabs(tp) - (ra + rb)) / l;
      if (s > XARM_CONTACT_MARGIN) return 0;
      if (s * 1.05 > best) { best = s; code = 6 + 3 * i + j; nsign = tp < 0 ? -1 : 1; v3scale(naxis, 1 / l, ax); }
    }
  if (code < 0) return 0;
  v3 nAB; v3scale(nAB, nsign, naxis); /* unit, from A towards B */
  double depth = -best;
  if (code >= 6) { /* edge-edge: one point */
    int i = (code - 6) / 3, j = (code - 6) % 3;
    v3 pa, pb;
    v3cpy(pa, A->c); v3cpy(pb, B->c);
    for (int k = 0; k < 3; k++) {
      if (k != i) v3axpy(pa, (v3dot(nAB, Aa[k]) > 0 ? 1 : -1) * A->h[k], Aa[k]);
      if (k != j) v3axpy(pb, (v3dot(nAB, Ba[k]) > 0 ? -1 : 1) * B->h[k], Ba[k]);
    }
    /* closest points of the lines pa + s Aa[i], pb + u Ba[j] */
    v3 d; v3sub(d, pb, pa);
    double uaub = v3dot(Aa[i], Ba[j]), q1 = v3dot(Aa[i], d), q2 = -v3dot(Ba[j], d), den = 1 - uaub * uaub;
    double sa = 0, sb = 0;
    if (den > 1e-4) { sa = (q1 + uaub * q2) / den; sb = (uaub * q1 + q2) / den; }
    v3axpy(pa, sa, Aa[i]); v3axpy(pb, sb, Ba[j]);
    if (max_out < 1) return 0;
    v3cpy(out[0].pa, pa); v3cpy(out[0].pb, pb); v3scale(out[0].n, -1, nAB); out[0].depth = depth;
    return 1;
  }
  /* face contact: reference box owns the axis, incident box is the other */
  const Collider* Rf = code < 3 ? A : B;
  const Collider* In = code < 3 ? B : A;
  int ra = code < 3 ? code : code - 3;
  v3 nref; /* outward normal of the reference face, pointing to the incident box */
  if (code < 3) v3cpy(nref, nAB); else v3scale(nref, -1, nAB);
  v3 Ra[3], Ia[3];
  for (int i = 0; i < 3; i++) { col_axis(Rf, i, Ra[i]); col_axis(In, i, Ia[i]); }
  /* incident face: most anti-parallel to nref */
  int ia = 0; double bd = -1;
  for (int i = 0; i < 3; i++) { double d = fabs(v3dot(nref, Ia[i])); if (d > bd) { bd = d; ia = i; } }
  double isg = v3dot(nref, Ia[ia]) > 0 ? -1 : 1;
  v3 fc; v3cpy(fc, In->c); v3axpy(fc, isg * In->h[ia], Ia[ia]);
  int i1 = (ia + 1) % 3, i2 = (ia + 2) % 3, r1 = (ra + 1) % 3, r2 = (ra + 2) % 3;
  double poly[8][2], tmp[8][2];
  static const double sg[4][2] = {{1, 1}, {-1, 1}, {-1, -1}, {1, -1}};
  v3 verts[4];
  for (int c = 0; c < 4; c++) {
    v3cpy(verts[c], fc);
    v3axpy(verts[c], sg[c][0] * In->h[i1], Ia[i1]);
    v3axpy(verts[c], sg[c][1] * In->h[i2], Ia[i2]);
    v3 d; v3sub(d, verts[c], Rf->c);
    poly[c][0] = v3dot(d, Ra[r1]); poly[c][1] = v3dot(d, Ra[r2]);
  }
  /* incident face plane in reference coordinates: depth varies linearly over (x,y) */
  int n = 4;
  n = clip_poly(poly, n, 0, 1, Rf->h[r1], tmp); if (!n) return 0;
  n = clip_poly(tmp, n, 0, -1, Rf->h[r1], poly); if (!n) return 0;
  n = clip_poly(poly, n, 1, 1, Rf->h[r2], tmp); if (!n) return 0;
  n = clip_poly(tmp, n, 1, -1, Rf->h[r2], poly); if (!n) return 0;
  /* lift the clipped 2-D points back onto the incident face plane: solve along nref */
  v3 inorm; v3scale(inorm, isg, Ia[ia]);
  double denom = v3dot(inorm, nref); /* < 0 */
  Contact cand[8]; int nc = 0;
  for (int c = 0; c < n; c++) {
    /* point on reference plane coords (x,y), height z along nref such that it lies on the incident plane */
    v3 base; v3cpy(base, Rf->c);
    v3axpy(base, poly[c][0], Ra[r1]); v3axpy(base, poly[c][1], Ra[r2]);
    v3 d; v3sub(d, fc, base);
    double z = fabs(denom) > 1e-9 ? v3dot(d, inorm) / denom : 0;
    v3 pin; v3cpy(pin, base); v3axpy(pin, z, nref); /* point on the incident face */
    double dep = Rf->h[ra] - z;
    if (dep < -XARM_CONTACT_MARGIN) continue;
    Contact* k = &cand[nc++];
    k->depth = dep;
    if (code < 3) { /* reference A, incident B: pin on B */
      v3cpy(k->pb, pin); v3cpy(k->pa, pin); v3axpy(k->pa, dep, nref); v3scale(k->n, -1, nref);
    } else {        /* reference B, incident A: pin on A */
      v3cpy(k->pa, pin); v3cpy(k->pb, pin); v3axpy(k->pb, dep, nref); v3cpy(k->n, nref);
    }
  }
  if (nc > 4) { /* manifold reduction: deepest, farthest from it, then the extremes on both sides of that line */
    int keep[4]; int i0 = 0;
    for (int c = 1; c < nc; c++) if (cand[c].depth > cand[i0].depth) i0 = c;
    keep[0] = i0;
    int ib = -1; double bdist = -1;
    for (int c = 0; c < nc; c++) { v3 d; v3sub(d, cand[c].pb, cand[i0].pb); double l = v3dot(d, d); if (c != i0 && l > bdist) { bdist = l; ib = c; } }
    keep[1] = ib;
    v3 e1; v3sub(e1, cand[ib].pb, cand[i0].pb);
    int imax = -1, imin = -1; double amax = 0, amin = 0;
    for (int c = 0; c < nc; c++) {
      if (c == i0 || c == ib) continue;
      v3 d, cr; v3sub(d, cand[c].pb, cand[i0].pb); v3cross(cr, e1, d);
      double ar = v3dot(cr, nref);
      if (imax < 0 || ar > amax) { amax = ar; imax = c; }
      if (imin < 0 || ar < amin) { amin = ar; imin = c; }
    }
    keep[2] = imax; keep[3] = imin;
    Contact red[4]; int nr = 0;
    for (int c = 0; c < 4; c++) {
      int dup = keep[c] < 0;
      for (int d = 0; d < c && !dup; d++) if (keep[d] == keep[c]) dup = 1;
      if (!dup) red[nr++] = cand[keep[c]];
    }
    memcpy(cand, red, sizeof(Contact) * nr); nc = nr;
  }
  if (nc > max_out) nc = max_out;
  memcpy(out, cand, sizeof(Contact) * nc);
  return nc;
}

/* ------------------------------------------------------------------------------------------------ world assembly */
typedef struct {
  int nv;
  int arm_off[MAXARM], obj_off[MAXOBJ], door_off;
  Kin kin[MAXARM];
  double Minv[MAXNV][MAXNV];
  double qd[MAXNV];
  m3 objR[MAXOBJ], objIinvW[MAXOBJ];
  Collider col[MAXCOLLIDER]; int ncol;
  int pair[MAXPAIR][2]; int npair;
  int col_table0, col_ground, col_obj0, col_f1[MAXARM], col_f2[MAXARM], col_hand[MAXARM], col_bar0;
} World;

typedef struct {
  double J[MAXNV], dV[MAXNV];
  double rhs, cfm, dinv, lo, hi, applied;
  int normal_row;  /* friction rows: index of their normal row */
  double mu;
} Row;

static void obj_inertia_inv_world(const Task* t, const ObjState* o, m3 R, m3 IinvW) {
  quat_to_m3(R, o->quat);
  double lx = 2 * t->obj_half[0], ly = 2 * t->obj_half[1], lz = 2 * t->obj_half[2], m = t->obj_mass;
  double Id[3] = {m / 12 * (ly * ly + lz * lz), m / 12 * (lx * lx + lz * lz), m / 12 * (lx * lx + ly * ly)};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += R[3 * i + k] * R[3 * j + k] / Id[k];
      IinvW[3 * i + j] = s;
    }
}

static void add_collider(World* w, int body, int index, int link, const v3 c, const m3 R, const v3 h, double fr,
                         double stiff, double damp, double erp) {
  Collider* k = &w->col[w->ncol++];
  k->body = body; k->index = index; k->link = link;
  v3cpy(k->c, c); memcpy(k->R, R, sizeof(m3)); v3cpy(k->h, h);
  k->friction = fr; k->stiffness = stiff; k->damping = damp; k->erp = erp;
}
static void add_pair(World* w, int a, int b) { w->pair[w->npair][0] = a; w->pair[w->npair][1] = b; w->npair++; }

/* Build kinematics, M^-1 blocks, colliders and the fixed pair list of the task (Appendix G). */
static void world_build(OrEnv* e, World* w) {
  const Task* t = &e->t;
  const Model* m = t->model;
  memset(w, 0, sizeof(*w));
  int off = 0;
  for (int a = 0; a < t->n_arms; a++) { w->arm_off[a] = off; off += m->ndof; }
  for (int o = 0; o < t->n_obj; o++) { w->obj_off[o] = off; off += 6; }
  if (t->has_door) { w->door_off = off; off += 1; }
  w->nv = off;
  for (int a = 0; a < t->n_arms; a++) {
    arm_fk(t, a, e->arm[a].q, &w->kin[a]);
    double zero[MAXDOF] = {0}, tau[MAXDOF], col[MAXDOF];
    for (int j = 0; j < m->ndof; j++) {
      memset(tau, 0, sizeof(tau)); tau[j] = 1;
      arm_aba(e, t, &w->kin[a], zero, tau, 0, col);
      for (int i = 0; i < m->ndof; i++) w->Minv[w->arm_off[a] + i][w->arm_off[a] + j] = col[i];
    }
    for (int i = 0; i < m->ndof; i++) w->qd[w->arm_off[a] + i] = e->arm[a].qd[i];
  }
  for (int o = 0; o < t->n_obj; o++) {
    obj_inertia_inv_world(t, &e->obj[o], w->objR[o], w->objIinvW[o]);
    int b = w->obj_off[o];
    for (int i = 0; i < 3; i++) {
      w->Minv[b + i][b + i] = 1 / t->obj_mass;
      for (int j = 0; j < 3; j++) w->Minv[b + 3 + i][b + 3 + j] = w->objIinvW[o][3 * i + j];
      w->qd[b + i] = e->obj[o].v[i]; w->qd[b + 3 + i] = e->obj[o].w[i];
    }
  }
  if (t->has_door) { w->Minv[w->door_off][w->door_off] = 1 / XARM_DOOR_MASS; w->qd[w->door_off] = e->door_qd; }
  /* colliders */
  m3 I3; m3ident(I3);
  w->col_table0 = w->ncol;
  for (int k = 0; k < t->n_tables; k++) {
    v3 h = {XARM_TABLE_HALF_X, XARM_TABLE_HALF_Y, XARM_TABLE_HALF_Z};
    add_collider(w, BODY_STATIC, 0, 0, t->table_c[k], I3, h, XARM_TABLE_FRICTION, 0, 0, -1);
  }
  w->col_ground = -1;
  if (t->has_ground) {
    v3 c = {0, 0, XARM_GROUND_Z - 5.0}, h = {100, 100, 5.0};
    w->col_ground = w->ncol;
    add_collider(w, BODY_STATIC, 0, 0, c, I3, h, 1.0, 0, 0, -1);
  }
  w->col_obj0 = w->ncol;
  for (int o = 0; o < t->n_obj; o++)
    add_collider(w, BODY_OBJ, o, 0, e->obj[o].pos, w->objR[o], t->obj_half, XARM_DEFAULT_FRICTION, 0, 0, -1);
  if (m->has_gripper_boxes)
    for (int a = 0; a < t->n_arms; a++) {
      const Kin* k = &w->kin[a];
      double ff = (t->friction_switch && e->grasp_cmd[a]) ? XARM_FINGER_FRICTION_GRASP : XARM_FINGER_FRICTION_FREE;
      v3 c;
      w->col_f1[a] = w->ncol;
      link_point(k, m->finger1, m->f1c, c);
      add_collider(w, BODY_ARM, a, m->finger1, c, k->R[m->finger1], m->f1h, ff, XARM_FINGER_STIFFNESS, XARM_FINGER_DAMPING, -1);
      w->col_f2[a] = w->ncol;
      link_point(k, m->finger2, m->f2c, c);
      add_collider(w, BODY_ARM, a, m->finger2, c, k->R[m->finger2], m->f2h, ff, XARM_FINGER_STIFFNESS, XARM_FINGER_DAMPING, -1);
      w->col_hand[a] = w->ncol;
      link_point(k, m->eef_dof, m->hc, c);
      add_collider(w, BODY_ARM, a, m->eef_dof, c, k->R[m->eef_dof], m->hh, XARM_DEFAULT_FRICTION, 0, 0, -1);
    }
  w->col_bar0 = -1;
  if (t->has_door) {
    static const double b1[] = XARM_DOOR_FIXED_BAR1, b2[] = XARM_DOOR_FIXED_BAR2, org[] = XARM_DOOR_ORIGIN,
                        ax[] = XARM_DOOR_AXIS, bh[] = XARM_DOOR_BAR_HALF;
    w->col_bar0 = w->ncol;
    add_collider(w, BODY_STATIC, 0, 0, b1, I3, bh, XARM_DOOR_FRICTION, 0, 0, 0.0);
    add_collider(w, BODY_STATIC, 0, 0, b2, I3, bh, XARM_DOOR_FRICTION, 0, 0, 0.0);
    v3 c; v3cpy(c, org); v3axpy(c, e->door_q, ax);
    add_collider(w, BODY_DOOR, 0, 0, c, I3, bh, XARM_DOOR_FRICTION, 0, 0, 0.0);
  }
  /* pair list, fixed order */
  for (int o = 0; o < t->n_obj; o++) {
    for (int k = 0; k < t->n_tables; k++) add_pair(w, w->col_obj0 + o, w->col_table0 + k);
    if (t->has_ground) add_pair(w, w->col_obj0 + o, w->col_ground);
  }
  for (int o = 0; o < t->n_obj; o++)
    for (int p = o + 1; p < t->n_obj; p++) add_pair(w, w->col_obj0 + o, w->col_obj0 + p);
  if (m->has_gripper_boxes)
    for (int a = 0; a < t->n_arms; a++)
      for (int o = 0; o < t->n_obj; o++) {
        add_pair(w, w->col_f1[a], w->col_obj0 + o);
        add_pair(w, w->col_f2[a], w->col_obj0 + o);
        add_pair(w, w->col_hand[a], w->col_obj0 + o);
      }
  if (t->has_door) {
    for (int o = 0; o < t->n_obj; o++) for (int b = 0; b < 3; b++) add_pair(w, w->col_obj0 + o, w->col_bar0 + b);
    for (int a = 0; a < t->n_arms; a++)
      for (int b = 0; b < 3; b++) {
        add_pair(w, w->col_f1[a], w->col_bar0 + b);
        add_pair(w, w->col_f2[a], w->col_bar0 + b);
        add_pair(w, w->col_hand[a], w->col_bar0 + b);
      }
  }
  if (t->task == XARM_TASK_HANDOVER)
    for (int a = 0; a < t->n_arms; a++)
      for (int k = 0; k < t->n_tables; k++) {
        add_pair(w, w->col_f1[a], w->col_table0 + k);
        add_pair(w, w->col_f2[a], w->col_table0 + k);
      }
}

/* dense Jacobian of "velocity of world point p on collider's body along direction d" */
static void point_jacobian(const OrEnv* e, const World* w, const Collider* c, const v3 p, const v3 d, double sign, double* J) {
  const Task* t = &e->t;
  if (c->body == BODY_ARM) {
    const Model* m = t->model;
    const Kin* k = &w->kin[c->index];
    v3 pxd; v3cross(pxd, p, d);
    for (int j = c->link; j >= 0; j = m->parent[j])
      J[w->arm_off[c->index] + j] += sign * (v3dot(&k->S[j][0], pxd) + v3dot(&k->S[j][3], d));
  } else if (c->body == BODY_OBJ) {
    int b = w->obj_off[c->index];
    v3 r, rxd; v3sub(r, p, e->obj[c->index].pos); v3cross(rxd, r, d);
    for (int i = 0; i < 3; i++) { J[b + i] += sign * d[i]; J[b + 3 + i] += sign * rxd[i]; }
  } else if (c->body == BODY_DOOR) {
    static const double ax[] = XARM_DOOR_AXIS;
    J[w->door_off] += sign * v3dot(ax, d);
  }
}

static void row_finish(OrEnv* e, const World* w, Row* r) {
  /* dV = M^-1 J^T, dinv = 1/(J dV + cfm)  (fillMultiBodyConstraint) */
  double den = 0;
  for (int i = 0; i < w->nv; i++) {
    double s = 0;
    for (int j = 0; j < w->nv; j++) s += w->Minv[i][j] * r->J[j];
    r->dV[i] = s;
  }
  for (int i = 0; i < w->nv; i++) den += r->J[i] * r->dV[i];
  r->dinv = 1.0 / (den + r->cfm);
  e->flops += 2.0 * w->nv * w->nv + 2.0 * w->nv;
}
static double row_relvel(const World* w, const Row* r) {
  double s = 0;
  for (int i = 0; i < w->nv; i++) s += r->J[i] * w->qd[i];
  return s;
}

static void plane_space(const v3 n, v3 p, v3 q) { /* btPlaneSpace1 */
  if (fabs(n[2]) > 0.7071067811865475244) {
    double a = n[1] * n[1] + n[2] * n[2], k = 1 / sqrt(a);
    v3set(p, 0, -n[2] * k, n[1] * k);
    v3set(q, a * k, -n[0] * p[2], n[0] * p[1]);
  } else {
    double a = n[0] * n[0] + n[1] * n[1], k = 1 / sqrt(a);
    v3set(p, -n[1] * k, n[0] * k, 0);
    v3set(q, -n[2] * p[1], n[2] * p[0], a * k);
  }
}

/* One internal substep (btMultiBodyDynamicsWorld::internalSingleStepSimulation):
 * collide -> unconstrained velocities (ABA) -> rows -> PGS -> integrate.  apply_damping: URDF joint damping torque,
 * which PyBullet adds once per stepSimulation call. */
static void substep(OrEnv* e, int apply_damping, int last) {
  const Task* t = &e->t;
  const Model* m = t->model;
  const double h = t->h;
  static _Thread_local World w;
  static _Thread_local Row rows[MAXROW];
  world_build(e, &w);
  /* 1. collision detection on the current poses */
  Contact contacts[MAXCONTACT]; int nc = 0, nac = 0;
  int pair_count[MAXPAIR];
  for (int p = 0; p < w.npair; p++) {
    const Collider *A = &w.col[w.pair[p][0]], *B = &w.col[w.pair[p][1]];
    pair_count[p] = 0;
    v3 d; v3sub(d, A->c, B->c);
    double ra = v3norm(A->h), rb = v3norm(B->h);
    if (v3dot(d, d) > (ra + rb + XARM_CONTACT_MARGIN) * (ra + rb + XARM_CONTACT_MARGIN)) continue;
    int room = MAXCONTACT - nc;
    int with_arm = A->body == BODY_ARM || B->body == BODY_ARM;
    if (with_arm && XARM_MAX_ARM_CONTACTS - nac < room) room = XARM_MAX_ARM_CONTACTS - nac;
    if (room <= 0) continue;
    int k = box_box(A, B, &contacts[nc], room < 4 ? room : 4);
    if (with_arm) { nac += k; e->arm_contacts += k; }
    for (int i = 0; i < k; i++) { contacts[nc + i].ca = w.pair[p][0]; contacts[nc + i].cb = w.pair[p][1]; }
    pair_count[p] = k; nc += k;
  }
  if (getenv("XARM_TRACE")) for (int c = 0; c < nc; c++)
    fprintf(stderr, "O env %ld c%d (%d,%d) pa %.6f %.6f %.6f pb %.6f %.6f %.6f n %.4f %.4f %.4f d %.6f\n", (long)e->env_index, c, w.col[contacts[c].ca].body * 10 + w.col[contacts[c].ca].link, w.col[contacts[c].cb].body,
            contacts[c].pa[0], contacts[c].pa[1], contacts[c].pa[2], contacts[c].pb[0], contacts[c].pb[1], contacts[c].pb[2], contacts[c].n[0], contacts[c].n[1], contacts[c].n[2], contacts[c].depth);
  /* grasp flags = "both finger links hold >= 1 manifold point with lego 0 / any lego" at the time of the last
   * collision pass [REF xarm_pick_and_place.py:212, xarm_handover.py:263-264] */
  if (last && m->has_gripper_boxes)
    for (int a = 0; a < t->n_arms; a++) {
      int g = 0;
      for (int o = 0; o < t->n_obj && !g; o++) {
        if (t->task == XARM_TASK_HANDOVER && o > 0) break;
        int c1 = 0, c2 = 0;
        for (int p = 0; p < w.npair; p++) {
          if (w.pair[p][1] != w.col_obj0 + o) continue;
          if (w.pair[p][0] == w.col_f1[a]) c1 += pair_count[p];
          if (w.pair[p][0] == w.col_f2[a]) c2 += pair_count[p];
        }
        g = c1 > 0 && c2 > 0;
      }
      e->grasp[a] = g;
    }
  /* 2. unconstrained velocity update */
  for (int a = 0; a < t->n_arms; a++) {
    double tau[MAXDOF], qdd[MAXDOF];
    for (int i = 0; i < m->ndof; i++) tau[i] = apply_damping ? -m->damping[i] * e->arm[a].qd[i] : 0.0;
    arm_aba(e, t, &w.kin[a], e->arm[a].qd, tau, 1, qdd);
    for (int i = 0; i < m->ndof; i++) w.qd[w.arm_off[a] + i] += qdd[i] * h;
  }
  for (int o = 0; o < t->n_obj; o++) {
    int b = w.obj_off[o];
    double kl = XARM_MB_LINEAR_DAMPING * (1 + v3norm(e->obj[o].v)), ka = XARM_MB_ANGULAR_DAMPING * (1 + v3norm(e->obj[o].w));
    for (int i = 0; i < 3; i++) {
      w.qd[b + i] += h * (-e->obj[o].v[i] * kl + (i == 2 ? -XARM_GRAVITY : 0.0));
      w.qd[b + 3 + i] += h * (-e->obj[o].w[i] * ka);
    }
    if (XARM_MB_USE_GYRO) { /* alpha -= I^-1 (w x I w) */ }
  }
  if (t->has_door) {
    double v = e->door_qd;
    double f = (apply_damping ? -XARM_DOOR_DAMPING * v : 0.0) - XARM_DOOR_MASS * v * XARM_MB_LINEAR_DAMPING * (1 + fabs(v));
    w.qd[w.door_off] += h * f / XARM_DOOR_MASS;
  }
  /* 3. rows: per arm [limits, motors, gear], door [limits, motor], then contacts */
  int nr = 0, n_noncontact;
  for (int a = 0; a < t->n_arms; a++) {
    int off = w.arm_off[a];
    const ArmState* s = &e->arm[a];
    for (int i = 0; i < m->ndof; i++)
      for (int side = 0; side < 2; side++) { /* btMultiBodyJointLimitConstraint (N5) */
        double pen = side == 0 ? s->q[i] - m->lo[i] : m->hi[i] - s->q[i];
        if (pen > 0) continue;
        Row* r = &rows[nr++]; memset(r, 0, sizeof(*r));
        r->J[off + i] = side == 0 ? 1 : -1;
        row_finish(e, &w, r);
        double rel = row_relvel(&w, r);
        r->rhs = (-pen * XARM_ERP / h - rel) * r->dinv;
        r->lo = 0; r->hi = XARM_LIMIT_MAX_IMPULSE;
      }
    for (int i = 0; i < m->ndof; i++) { /* btMultiBodyJointMotor position target (N4) */
      Row* r = &rows[nr++]; memset(r, 0, sizeof(*r));
      r->J[off + i] = 1;
      row_finish(e, &w, r);
      double cur = w.qd[off + i];
      double target = XARM_MOTOR_KP * XARM_MOTOR_ERP * (s->qt[i] - s->q[i]) / h + cur + XARM_MOTOR_KD * (0 - cur);
      r->rhs = (target - cur) * r->dinv;
      int is_arm = i < 7;
      double force = is_arm ? t->arm_force : t->finger_force;
      r->hi = force * t->time_step; r->lo = -r->hi;
    }
    if (t->gear) { /* btMultiBodyGearConstraint (N6) */
      Row* r = &rows[nr++]; memset(r, 0, sizeof(*r));
      r->J[off + m->finger1] = 1; r->J[off + m->finger2] = XARM_GEAR_RATIO;
      row_finish(e, &w, r);
      double rel = row_relvel(&w, r);
      double pos_err = XARM_GEAR_ERP * (s->q[m->finger1] + XARM_GEAR_RATIO * s->q[m->finger2]);
      r->rhs = (-pos_err * XARM_ERP / h + (0 - rel)) * r->dinv;
      r->hi = XARM_GEAR_MAX_FORCE * t->time_step; r->lo = -r->hi;
    }
  }
  if (t->has_door) {
    for (int side = 0; side < 2; side++) {
      double pen = side == 0 ? e->door_q - XARM_DOOR_LIMIT_LO : XARM_DOOR_LIMIT_HI - e->door_q;
      if (pen > 0) continue;
      Row* r = &rows[nr++]; memset(r, 0, sizeof(*r));
      r->J[w.door_off] = side == 0 ? 1 : -1;
      row_finish(e, &w, r);
      r->rhs = (-pen * XARM_ERP / h - row_relvel(&w, r)) * r->dinv;
      r->lo = 0; r->hi = XARM_LIMIT_MAX_IMPULSE;
    }
    Row* r = &rows[nr++]; memset(r, 0, sizeof(*r)); /* loader's default velocity motor: target 0, max impulse 1 */
    r->J[w.door_off] = 1;
    row_finish(e, &w, r);
    r->rhs = (0 - w.qd[w.door_off]) * r->dinv;
    r->hi = XARM_DEFAULT_MOTOR_MAX_IMPULSE; r->lo = -r->hi;
  }
  n_noncontact = nr;
  int n_normal0 = nr;
  for (int c = 0; c < nc; c++) { /* normal rows (setupMultiBodyContactConstraint) */
    const Contact* k = &contacts[c];
    const Collider *A = &w.col[k->ca], *B = &w.col[k->cb];
    Row* r = &rows[nr++]; memset(r, 0, sizeof(*r));
    point_jacobian(e, &w, A, k->pa, k->n, 1, r->J);
    point_jacobian(e, &w, B, k->pb, k->n, -1, r->J);
    double erp = XARM_ERP2, cfm = 0;
    if (A->erp >= 0) erp = A->erp;
    if (B->erp >= 0) erp = B->erp;
    if (A->stiffness > 0 || B->stiffness > 0) {
      double ks = A->stiffness > 0 && B->stiffness > 0 ? 1 / (1 / A->stiffness + 1 / B->stiffness)
                                                       : (A->stiffness > 0 ? A->stiffness : B->stiffness);
      double kd = A->damping + B->damping;
      cfm = 1 / (h * ks + kd); erp = h * ks / (h * ks + kd); cfm /= h;
    }
    r->cfm = cfm;
    row_finish(e, &w, r);
    double rel = row_relvel(&w, r);
    double pen = -k->depth + XARM_LINEAR_SLOP;
    double pos_err = 0, vel_err = -rel;
    if (pen > 0) vel_err -= pen / h; else pos_err = -pen * erp / h;
    r->rhs = (pos_err + vel_err) * r->dinv;
    r->cfm = cfm * r->dinv; /* Bullet stores m_cfm = cfm * jacDiagABInv */
    r->lo = 0; r->hi = XARM_CONTACT_MAX_IMPULSE;
  }
  int n_fric0 = nr;
  for (int c = 0; c < nc; c++) { /* friction rows: two directions from btPlaneSpace1(normal) */
    const Contact* k = &contacts[c];
    const Collider *A = &w.col[k->ca], *B = &w.col[k->cb];
    double mu = A->friction * B->friction;
    if (mu > XARM_MAX_FRICTION) mu = XARM_MAX_FRICTION;
    v3 t1, t2; plane_space(k->n, t1, t2);
    for (int d = 0; d < (XARM_TWO_FRICTION_DIRS ? 2 : 1); d++) {
      Row* r = &rows[nr++]; memset(r, 0, sizeof(*r));
      const double* dir = d == 0 ? t1 : t2;
      point_jacobian(e, &w, A, k->pa, dir, 1, r->J);
      point_jacobian(e, &w, B, k->pb, dir, -1, r->J);
      row_finish(e, &w, r);
      r->rhs = (0 - row_relvel(&w, r)) * r->dinv;
      r->normal_row = n_normal0 + c; r->mu = mu;
    }
  }
  /* 4. PGS (btMultiBodyConstraintSolver::solveSingleIteration) */
  double dqd[MAXNV]; memset(dqd, 0, sizeof(dqd));
  for (int it = 0; it < XARM_SOLVER_ITERATIONS; it++) {
    double resid = 0;
    for (int j = 0; j < n_noncontact + nc; j++) {
      int idx = j;
      if (j < n_noncontact) idx = (it & 1) ? j : n_noncontact - 1 - j;
      Row* r = &rows[idx];
      double jv = 0;
      for (int i = 0; i < w.nv; i++) jv += r->J[i] * dqd[i];
      double delta = r->rhs - r->applied * r->cfm - jv * r->dinv;
      double sum = r->applied + delta;
      if (sum < r->lo) { delta = r->lo - r->applied; sum = r->lo; }
      else if (sum > r->hi) { delta = r->hi - r->applied; sum = r->hi; }
      r->applied = sum;
      for (int i = 0; i < w.nv; i++) dqd[i] += r->dV[i] * delta;
      double rv = delta / r->dinv;
      if (rv * rv > resid) resid = rv * rv;
      e->flops += 4.0 * w.nv + 8;
    }
    for (int j = n_fric0; j < nr;) {
      if (XARM_TWO_FRICTION_DIRS) { /* resolveConeFrictionConstraintRows: both deltas from the same state, cone clamp */
        Row *ra = &rows[j], *rb = &rows[j + 1];
        double lim = ra->mu * rows[ra->normal_row].applied;
        double ja = 0, jb = 0;
        for (int i = 0; i < w.nv; i++) { ja += ra->J[i] * dqd[i]; jb += rb->J[i] * dqd[i]; }
        double da = ra->rhs - ra->applied * ra->cfm - ja * ra->dinv, db = rb->rhs - rb->applied * rb->cfm - jb * rb->dinv;
        double sa = ra->applied + da, sb = rb->applied + db;
        double len = sqrt(sa * sa + sb * sb);
        if (len > lim) { double sc = len > 0 ? lim / len : 0; sa *= sc; sb *= sc; da = sa - ra->applied; db = sb - rb->applied; }
        ra->applied = sa; rb->applied = sb;
        for (int i = 0; i < w.nv; i++) dqd[i] += ra->dV[i] * da + rb->dV[i] * db;
        double r1 = da / ra->dinv, r2 = db / rb->dinv;
        if (r1 * r1 > resid) resid = r1 * r1;
        if (r2 * r2 > resid) resid = r2 * r2;
        e->flops += 8.0 * w.nv + 30;
        j += 2;
      } else {
        Row* r = &rows[j];
        double lim = r->mu * rows[r->normal_row].applied;
        double jv = 0;
        for (int i = 0; i < w.nv; i++) jv += r->J[i] * dqd[i];
        double delta = r->rhs - jv * r->dinv, sum = r->applied + delta;
        if (sum < -lim) { delta = -lim - r->applied; sum = -lim; } else if (sum > lim) { delta = lim - r->applied; sum = lim; }
        r->applied = sum;
        for (int i = 0; i < w.nv; i++) dqd[i] += r->dV[i] * delta;
        double rv = delta / r->dinv;
        if (rv * rv > resid) resid = rv * rv;
        j += 1;
      }
    }
    e->st_iters += 1;
    if (resid <= XARM_RESIDUAL_THRESHOLD) break;
  }
  e->st_substeps += 1; e->st_rows += nr; e->st_contacts += nc; e->st_hist_nc[nc] += 1; e->st_hist_nac[nac] += 1;
  for (int i = 0; i < w.nv; i++) w.qd[i] += dqd[i];
  /* 5. integrate positions (stepPositionsMultiDof) */
  for (int a = 0; a < t->n_arms; a++)
    for (int i = 0; i < m->ndof; i++) {
      e->arm[a].qd[i] = w.qd[w.arm_off[a] + i];
      e->arm[a].q[i] += e->arm[a].qd[i] * h;
    }
  for (int o = 0; o < t->n_obj; o++) {
    ObjState* s = &e->obj[o];
    int b = w.obj_off[o];
    for (int i = 0; i < 3; i++) { s->v[i] = w.qd[b + i]; s->w[i] = w.qd[b + 3 + i]; s->pos[i] += s->v[i] * h; }
    double ang = v3norm(s->w);
    if (ang * h > XARM_ANGULAR_MOTION_THRESHOLD) ang = XARM_ANGULAR_MOTION_THRESHOLD / h;
    v3 ax;
    if (ang < 0.001) v3scale(ax, 0.5 * h - h * h * h * 0.020833333333 * ang * ang, s->w);
    else v3scale(ax, sin(0.5 * ang * h) / ang, s->w);
    double dq[4] = {ax[0], ax[1], ax[2], cos(ang * h * 0.5)}, qn[4];
    quat_mul(qn, dq, s->quat);
    double nn = sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
    for (int i = 0; i < 4; i++) s->quat[i] = qn[i] / nn;
  }
  if (t->has_door) { e->door_qd = w.qd[w.door_off]; e->door_q += e->door_qd * h; }
  if (getenv("XARM_TRACE")) {
    fprintf(stderr, "OS env %ld nc %d qd", (long)e->env_index, nc);
    for (int i = 0; i < w.nv; i++) fprintf(stderr, " %.10f", w.qd[i]);
    fprintf(stderr, "\n");
  }
}

/* p.stepSimulation() x (number of calls per env step) */
static void simulate(OrEnv* e) {
  const Task* t = &e->t;
  for (int s = 0; s < t->n_sub; s++) substep(e, t->damping_each_sub || s == 0, s == t->n_sub - 1);
}

/* ------------------------------------------------------------------------------------------------ rewards (K6)
 * float32 evaluation of the reference's NumPy expressions (what SB3's HER computes on float32 goal arrays).
 * np.linalg.norm(x, axis=-1) = sqrt(add.reduce(x*x)); add.reduce over a contiguous row of n floats is NumPy's
 * pairwise sum: plain left-to-right for n < 8, eight partial sums for n >= 8. */
static float np_sum_f32(const float* a, int n) {
  if (n < 8) { float s = a[0]; for (int i = 1; i < n; i++) s += a[i]; return s; }
  float r[8];
  for (int i = 0; i < 8; i++) r[i] = a[i];
  int i = 8;
  for (; i + 8 <= n; i += 8) for (int j = 0; j < 8; j++) r[j] += a[i + j];
  float s = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; i++) s += a[i];
  return s;
}
static float np_norm_f32(const float* a, const float* b, int n) {
  float sq[16];
  for (int i = 0; i < n; i++) { float d = a[i] - b[i]; sq[i] = d * d; }
  return sqrtf(np_sum_f32(sq, n));
}

static float reward_stateless(int task, int reward_type, int num_obj, float thr, const float* ag, const float* dg, int G) {
  switch (task) {
    case XARM_TASK_REACH: { /* [REF xarm_reach.py:107-112] */
      float d = np_norm_f32(ag, dg, 3);
      return reward_type == XARM_REWARD_SPARSE ? (d < thr ? 1.0f : 0.0f) : -d;
    }
    case XARM_TASK_PICK_AND_PLACE: { /* [REF xarm_pick_and_place.py:163-165,176-177] */
      float d = np_norm_f32(ag, dg, G);
      return reward_type == XARM_REWARD_SPARSE ? (d < thr ? 1.0f : 0.0f) : -d;
    }
    case XARM_TASK_STACK_TOWER:
    case XARM_TASK_PUSH_WITH_DOOR: { /* [REF xarm_stack_tower.py:124-129] -(d > thr).astype(f32): -1 or -0.0 */
      float d = np_norm_f32(ag, dg, G);
      return reward_type == XARM_REWARD_SPARSE ? -(d > thr ? 1.0f : 0.0f) : -d;
    }
    case XARM_TASK_HANDOVER: { /* [REF xarm_handover.py:177-183] -sum_i(||ag_i-g_i|| > thr) */
      float s = 0;
      for (int i = 0; i < num_obj; i++) s += np_norm_f32(ag + 3 * i, dg + 3 * i, 3) > thr ? 1.0f : 0.0f;
      return -s;
    }
  }
  return 0;
}

void or_compute_reward(int32_t task, int32_t reward_type, int32_t num_obj, const float* ag, const float* dg, int64_t n, float* out) {
  Task t;
  if (task_fill(&t, task, num_obj)) return;
  for (int64_t i = 0; i < n; i++)
    out[i] = reward_stateless(task, reward_type, num_obj, t.threshold, ag + i * t.goal_dim, dg + i * t.goal_dim, t.goal_dim);
}

/* ------------------------------------------------------------------------------------------------ obs (_get_obs) */
typedef struct { float obs[64]; float ag[9]; float dg[9]; } ObsOut;

static void hand_state(const OrEnv* e, int arm, v3 pos, v3 vel) {
  /* getLinkState(arm, 9, computeLinkVelocity=1): COM position [0] and COM linear velocity [6] */
  const Task* t = &e->t;
  Kin k;
  arm_fk(t, arm, e->arm[arm].q, &k);
  link_point(&k, t->model->eef_dof, t->model->hand_com, pos);
  link_point_vel(t->model, &k, t->model->eef_dof, e->arm[arm].qd, pos, vel);
}

/* pure assembly of the observation dict from the quantities the reference reads through PyBullet getters */
typedef struct {
  v3 hand_pos[MAXARM], hand_vel[MAXARM];
  double finger_q[MAXARM], finger_qd[MAXARM];
  ObjState obj[MAXOBJ];
  float goal[3 * MAXOBJ];
} ObsIn;

static void assemble_obs(const Task* t, const ObsIn* in, ObsOut* o) {
  int n = 0;
  const v3* hp = in->hand_pos; const v3* hv = in->hand_vel;
  if (t->task == XARM_TASK_REACH) { /* [REF xarm_reach.py:144-161] */
    for (int c = 0; c < 3; c++) o->obs[n++] = (float)hp[0][c];
    for (int c = 0; c < 3; c++) o->obs[n++] = (float)hv[0][c];
    o->obs[n++] = (float)in->finger_q[0];
    o->obs[n++] = (float)in->finger_qd[0];
    for (int c = 0; c < 3; c++) { o->ag[c] = (float)hp[0][c]; o->dg[c] = in->goal[c]; }
    return;
  }
  if (t->task == XARM_TASK_PICK_AND_PLACE) { /* [REF xarm_pick_and_place.py:220-248] */
    for (int c = 0; c < 3; c++) o->obs[n++] = (float)hp[0][c];
    for (int c = 0; c < 3; c++) o->obs[n++] = (float)hv[0][c];
    o->obs[n++] = (float)in->finger_q[0];
    o->obs[n++] = (float)in->finger_qd[0];
    for (int i = 0; i < t->n_obj; i++) {
      const ObjState* s = &in->obj[i];
      for (int c = 0; c < 3; c++) o->obs[n++] = (float)s->pos[c];
      for (int c = 0; c < 4; c++) o->obs[n++] = (float)s->quat[c];
      for (int c = 0; c < 3; c++) o->obs[n++] = (float)(s->v[c] - hv[0][c]);
      for (int c = 0; c < 3; c++) o->obs[n++] = (float)s->w[c];
      for (int c = 0; c < 3; c++) o->obs[n++] = (float)(s->pos[c] - hp[0][c]);
      for (int c = 0; c < 3; c++) o->ag[3 * i + c] = (float)s->pos[c];
    }
  } else { /* StackTower / PushWithDoor / Handover [REF xarm_stack_tower.py:164-199; xarm_push_with_door.py:162-193; xarm_handover.py:299-336] */
    for (int i = 0; i < t->n_obj; i++) for (int c = 0; c < 3; c++) o->obs[n++] = (float)in->obj[i].pos[c];
    for (int i = 0; i < t->n_obj; i++) for (int c = 0; c < 4; c++) o->obs[n++] = (float)in->obj[i].quat[c];
    for (int i = 0; i < t->n_obj; i++) for (int c = 0; c < 3; c++) o->obs[n++] = (float)in->obj[i].v[c];
    for (int i = 0; i < t->n_obj; i++) for (int c = 0; c < 3; c++) o->obs[n++] = (float)in->obj[i].w[c];
    for (int a = 0; a < t->n_arms; a++) {
      for (int c = 0; c < 3; c++) o->obs[n++] = (float)(hp[a][c] - t->hand_offset[c]);
      for (int c = 0; c < 3; c++) o->obs[n++] = (float)hv[a][c];
      if (t->task != XARM_TASK_PUSH_WITH_DOOR) {
        o->obs[n++] = (float)in->finger_q[a];
        o->obs[n++] = (float)in->finger_qd[a];
      }
    }
    for (int i = 0; i < t->n_obj; i++) for (int c = 0; c < 3; c++) o->ag[3 * i + c] = (float)in->obj[i].pos[c];
  }
  for (int c = 0; c < t->goal_dim; c++) o->dg[c] = in->goal[c];
}

static void get_obs(const OrEnv* e, ObsOut* o) {
  const Task* t = &e->t;
  ObsIn in;
  for (int a = 0; a < t->n_arms; a++) {
    hand_state(e, a, in.hand_pos[a], in.hand_vel[a]);
    in.finger_q[a] = e->arm[a].q[t->model->finger1];
    in.finger_qd[a] = e->arm[a].qd[t->model->finger1];
  }
  for (int i = 0; i < t->n_obj; i++) in.obj[i] = e->obj[i];
  memcpy(in.goal, e->goal, sizeof(in.goal));
  assemble_obs(t, &in, o);
}

/* staged dense rewards that read the live simulator (not batch-safe in the reference either).  Pure function of what
 * the reference reads: hand-link COM positions (getLinkState(arm, 9)[0]), the achieved / desired goal, the grasp flags.
 * PickAndPlace queries getContactPoints inside compute_reward (flags AFTER the step) [REF xarm_pick_and_place.py:166-175];
 * Handover reads self.if_xarm*_grasp, which _set_action stored BEFORE the step's 15 stepSimulation calls
 * [REF xarm_handover.py:185-199, 262-263].  D2: Handover's last branch uses d = ||ag - goal||. */
static float dense_staged_core(int task, int goal_dim, const double hand[MAXARM][3], const float* ag, const float* dg, const int* grasp) {
  float z[3] = {0, 0, 0};
  if (task == XARM_TASK_PICK_AND_PLACE) {
    float g[3] = {(float)hand[0][0], (float)hand[0][1], (float)(hand[0][2] - (0.088 - 0.021))};
    float off[3] = {0.06f, 0, 0}, d3[3];
    for (int c = 0; c < 3; c++) d3[c] = g[c] - ag[c] + off[c];
    float d_ao = np_norm_f32(d3, z, 3), d_og = np_norm_f32(ag, dg, goal_dim);
    if (!grasp[0]) return 0.25f * (1 - tanhf(d_ao));
    if (ag[2] > 0.05f) return 1.0f + 0.25f * (1 - tanhf(d_og));
    return 0.5f;
  }
  float p1[3], p2[3];
  for (int c = 0; c < 3; c++) {
    const double off = c == 2 ? 0.088 - 0.021 : 0.0;
    float h1 = (float)(hand[0][c] - off), h2 = (float)(hand[1][c] - off); /* the float32 words of _get_obs */
    p1[c] = h1 - ag[c]; p2[c] = h2 - ag[c];
  }
  p1[0] += 0.06f; p2[0] -= 0.06f;
  float d1 = np_norm_f32(p1, z, 3), d2 = np_norm_f32(p2, z, 3);
  int g1 = grasp[0], g2 = grasp[1];
  if (!g1 && !g2) return 0.25f * (1 - tanhf(d1)) / 2.25f;
  if (g1 && !g2) return ag[2] > 0.05f ? (1.0f + 0.25f * (1 - tanhf(d2))) / 2.25f : 0.5f / 2.25f;
  if (g1 && g2) return 1.5f / 2.25f;
  return (2.0f + 0.25f * (1 - tanhf(np_norm_f32(ag, dg, goal_dim)))) / 2.25f;
}
static float reward_dense_staged(const OrEnv* e, const ObsOut* o) {
  const Task* t = &e->t;
  double hand[MAXARM][3]; v3 hv;
  for (int a = 0; a < t->n_arms; a++) hand_state(e, a, hand[a], hv);
  return dense_staged_core(t->task, t->goal_dim, hand, o->ag, o->dg, t->task == XARM_TASK_PICK_AND_PLACE ? e->grasp : e->grasp_cmd);
}
float or_debug_dense_reward(int32_t task, int32_t num_obj, const double* hand /* [n_arms][3] link-9 COM */, const float* ag, const float* dg, const int32_t* grasp) {
  double h[MAXARM][3] = {{0}}; int g[MAXARM] = {0};
  int na = task == XARM_TASK_PICK_AND_PLACE ? 1 : 2;
  for (int a = 0; a < na; a++) { for (int c = 0; c < 3; c++) h[a][c] = hand[3 * a + c]; g[a] = grasp[a]; }
  return dense_staged_core(task, 3 * num_obj, h, ag, dg, g);
}

/* ------------------------------------------------------------------------------------------------ reset / goals */
static void set_joint_init(OrEnv* e, int arm, double finger) {
  const Model* m = e->t.model;
  memset(&e->arm[arm], 0, sizeof(ArmState));
  for (int i = 0; i < 7; i++) e->arm[arm].q[i] = JOINT_INIT[i + 1];
  if (m->finger2 >= 0) { e->arm[arm].q[m->finger1] = finger; e->arm[arm].q[m->finger2] = finger; }
  for (int i = 0; i < m->ndof; i++) e->arm[arm].qt[i] = e->arm[arm].q[i];
}
static void place_obj(OrEnv* e, int i, double x, double y, double z) {
  ObjState* s = &e->obj[i];
  memset(s, 0, sizeof(*s));
  s->pos[0] = x; s->pos[1] = y; s->pos[2] = z; s->quat[3] = 1;
}

static void sample_goal(OrEnv* e) {
  const Task* t = &e->t;
  switch (t->task) {
    case XARM_TASK_REACH: /* [REF xarm_reach.py:170-173] goal_space [REF :27] */
      e->goal[0] = box_sample(e, 0.3f, 0.5f); e->goal[1] = box_sample(e, -0.25f, 0.25f); e->goal[2] = box_sample(e, 0.3f, 0.4f);
      break;
    case XARM_TASK_PICK_AND_PLACE: { /* [REF xarm_pick_and_place.py:269-287] goal_space [REF :38] */
      const float lo[3] = {0.35f, -0.25f, 0.025f}, hi[3] = {0.45f, 0.25f, 0.27f};
      if (e->cfg.goal_shape == XARM_GOAL_AIR) {
        for (int i = 0; i < t->n_obj; i++) {
          float* g = &e->goal[3 * i];
          for (int c = 0; c < 3; c++) g[c] = box_sample(e, lo[c], hi[c]);
          if (rng_uniform(e) < (double)e->cfg.goal_ground_rate) g[2] = lo[2];
          for (int tries = 0; i > 0 && tries < XARM_GOAL_RESAMPLE_TRIES; tries++) {
            double mn = 1e30;
            for (int j = 0; j < i; j++) {
              double s = 0;
              for (int c = 0; c < 3; c++) { double d = (double)g[c] - (double)e->goal[3 * j + c]; s += d * d; }
              if (sqrt(s) < mn) mn = sqrt(s);
            }
            if (!(mn < 0.05)) break;
            for (int c = 0; c < 3; c++) g[c] = box_sample(e, lo[c], hi[c]);
          }
        }
      } else {
        float xy[3];
        for (int c = 0; c < 3; c++) xy[c] = box_sample(e, lo[c], hi[c]);
        for (int i = 0; i < t->n_obj; i++) {
          e->goal[3 * i] = xy[0]; e->goal[3 * i + 1] = xy[1]; e->goal[3 * i + 2] = (float)(0.025 * (2 * i + 1));
        }
      }
      break;
    }
    case XARM_TASK_STACK_TOWER: { /* [REF xarm_stack_tower.py:213-219] */
      float x = box_sample(e, -0.3f, 0.3f), y = box_sample(e, -0.2f, 0.2f);
      for (int i = 0; i < 3; i++) { e->goal[3 * i] = x; e->goal[3 * i + 1] = y; e->goal[3 * i + 2] = (float)(0.025 * (2 * i + 1)); }
      break;
    }
    case XARM_TASK_PUSH_WITH_DOOR: /* [REF xarm_push_with_door.py:207-212] */
      e->goal[0] = box_sample(e, 0.1f, 0.3f); e->goal[1] = box_sample(e, -0.2f, 0.2f); e->goal[2] = (float)0.025;
      break;
    case XARM_TASK_HANDOVER: { /* [REF xarm_handover.py:370-393] goal_space [REF :43] */
      const float lo[3] = {0.1f, -0.18f, 0.025f}, hi[3] = {0.28f, 0.18f, 0.2f};
      for (int i = 0; i < t->n_obj; i++) {
        float* g = &e->goal[3 * i];
        for (int c = 0; c < 3; c++) g[c] = box_sample(e, lo[c], hi[c]);
        for (int tries = 0; i > 0 && tries < XARM_GOAL_RESAMPLE_TRIES; tries++) {
          double mn_obj = 1e30, mn_g = 1e30;
          for (int k = 0; k < t->n_obj; k++) {
            double dx = (double)g[0] - e->obj[k].pos[0], dy = (double)g[1] - e->obj[k].pos[1];
            double d = sqrt(dx * dx + dy * dy);
            if (d < mn_obj) mn_obj = d;
          }
          for (int j = 0; j < i; j++) { double d = fabs((double)g[1] - (double)e->goal[3 * j + 1]); if (d < mn_g) mn_g = d; }
          if (!(mn_g < 0.08 || mn_obj < 0.08)) break;
          for (int c = 0; c < 3; c++) g[c] = box_sample(e, lo[c], hi[c]);
        }
        int same = rng_uniform(e) < (double)e->cfg.same_side_rate;
        if ((e->obj[i].pos[0] > 0) ^ same) g[0] = -g[0];
        if (e->cfg.goal_shape == XARM_GOAL_GROUND) g[2] = (float)0.025;
      }
      break;
    }
  }
}

static void servo_reset(OrEnv* e, const double targets[MAXARM][3], int cmd_fingers) {
  /* 5x { IK to the start pose, motor commands, stepSimulation }  [REF xarm_pick_and_place.py:252-258; xarm_handover.py:347-353] */
  const Task* t = &e->t;
  const Model* m = t->model;
  for (int rep = 0; rep < 5; rep++) {
    for (int a = 0; a < t->n_arms; a++) {
      double qn[MAXDOF];
      arm_ik(t, a, e->arm[a].q, targets[a], qn);
      for (int i = 0; i < 7; i++) e->arm[a].qt[i] = qn[i];
      if (cmd_fingers) { e->arm[a].qt[m->finger1] = 0.02; e->arm[a].qt[m->finger2] = 0.02; }
    }
    simulate(e);
  }
}

static void env_reset(OrEnv* e) {
  const Task* t = &e->t;
  e->episode += 1; e->rng_draw = 0; e->step_count = 0;
  switch (t->task) {
    case XARM_TASK_REACH: /* [REF xarm_reach.py:96-102,163-168] teleport + one stepSimulation */
      set_joint_init(e, 0, 0);
      simulate(e);
      sample_goal(e);
      break;
    case XARM_TASK_PICK_AND_PLACE: { /* [REF xarm_pick_and_place.py:121-127,250-267] */
      const double tg[MAXARM][3] = {{0.4, 0, 0.12}, {0, 0, 0}};
      servo_reset(e, tg, 1);
      for (int i = 0; i < t->n_obj; i++) {
        double ug = rng_uniform(e);
        float x = box_sample(e, 0.35f, 0.45f), y = box_sample(e, -0.25f, 0.25f);
        if (ug < (double)e->cfg.init_grasp_rate) place_obj(e, i, 0.4, 0.0, 0.025);
        else place_obj(e, i, x, y, 0.025);
      }
      simulate(e);
      sample_goal(e);
      break;
    }
    case XARM_TASK_STACK_TOWER:   /* [REF xarm_stack_tower.py:115-119,201-211] */
    case XARM_TASK_PUSH_WITH_DOOR: { /* [REF xarm_push_with_door.py:117-121,195-205] */
      for (int a = 0; a < 2; a++) set_joint_init(e, a, 0);
      for (int i = 0; i < t->n_obj; i++) {
        float x, y;
        if (t->task == XARM_TASK_STACK_TOWER) { x = box_sample(e, -0.3f, 0.3f); y = box_sample(e, -0.2f, 0.2f); }
        else { x = box_sample(e, -0.3f, -0.1f); y = box_sample(e, -0.2f, 0.2f); }
        place_obj(e, i, x, y, 0.025);
      }
      simulate(e);
      sample_goal(e);
      break;
    }
    case XARM_TASK_HANDOVER: { /* [REF xarm_handover.py:141-145,338-368] */
      const double tg[MAXARM][3] = {{-0.15, 0, 0.15}, {0.15, 0, 0.15}};
      servo_reset(e, tg, 0);
      float pos[MAXOBJ][2];
      for (int i = 0; i < t->n_obj; i++) {
        pos[i][0] = box_sample(e, 0.11f, 0.28f); pos[i][1] = box_sample(e, -0.18f, 0.2f);
        for (int tries = 0; i > 0 && tries < XARM_GOAL_RESAMPLE_TRIES; tries++) {
          double mn = 1e30;
          for (int j = 0; j < i; j++) { double d = fabs((double)pos[i][1] - (double)pos[j][1]); if (d < mn) mn = d; }
          if (!(mn < 0.05)) break;
          pos[i][0] = box_sample(e, 0.11f, 0.28f); pos[i][1] = box_sample(e, -0.18f, 0.2f);
        }
        if (rng_uniform(e) < 0.5) pos[i][0] = -pos[i][0];
        place_obj(e, i, pos[i][0], pos[i][1], 0.025);
      }
      simulate(e);
      sample_goal(e);
      break;
    }
  }
  ObsOut o; get_obs(e, &o);
  e->d_old = np_norm_f32(o.ag, o.dg, t->goal_dim); /* [REF xarm_reach.py:100] */
}

/* ------------------------------------------------------------------------------------------------ step (_set_action + step) */
/* pure part of _set_action: eef target and finger target of one arm from the current eef position / finger joint */
static void arm_command(const Task* t, int a, const float* u, const v3 eef, double finger_q, v3 target, double* grip) {
  for (int c = 0; c < 3; c++) {
    /* [REF xarm_pick_and_place.py:202-204] `np.array(a[:3]) * max_vel * dt` is float32-array x Python-scalar
     * arithmetic (two float32 roundings); the sum with the float64 eef position and the clip are float64 */
    float stepf = (float)(u[c] * (float)t->max_vel);
    stepf = (float)(stepf * (float)t->dt_cmd);
    double np_ = eef[c] + (double)stepf;
    double lo = t->pos_lo[a][c], hi = t->pos_hi[a][c];
    target[c] = np_ < lo ? lo : (np_ > hi ? hi : np_);
  }
  double g = finger_q;
  if (t->grip_cmd) {
    g = finger_q + (double)u[3] * t->dt_cmd * t->max_grip_vel; /* [REF :205-206] float32 scalar x Python floats: float64 (NumPy 1.x) */
    if (t->grip_clip) { double lo = t->grip_lo, hi = t->grip_hi; g = g < lo ? lo : (g > hi ? hi : g); }
  }
  *grip = g;
}
/* Handover: clamp |x|,|y| and keep only pitch, zeroing velocity [REF xarm_handover.py:282-297] */
static void lego_clamp(ObjState* s) {
  double x = s->pos[0], y = s->pos[1];
  const double hx = (float)0.28f, hy = (float)0.2f;
  int neg = x < 0; if (neg) x = -x;
  x = x < -hx ? -hx : (x > hx ? hx : x); y = y < -hy ? -hy : (y > hy ? hy : y);
  if (neg) x = -x;
  m3 R; quat_to_m3(R, s->quat);
  /* getEulerFromQuaternion -> pitch; getQuaternionFromEuler([0,pitch,0]) */
  double sp = -R[6]; sp = sp < -1 ? -1 : (sp > 1 ? 1 : sp);
  double pitch = asin(sp);
  s->pos[0] = x; s->pos[1] = y;
  s->quat[0] = 0; s->quat[1] = sin(pitch / 2); s->quat[2] = 0; s->quat[3] = cos(pitch / 2);
  memset(s->v, 0, sizeof(v3)); memset(s->w, 0, sizeof(v3));
}

static void set_action(OrEnv* e, const float* act_in) {
  const Task* t = &e->t;
  const Model* m = t->model;
  float act[8];
  for (int i = 0; i < t->act_dim; i++) act[i] = act_in[i] < -1 ? -1 : (act_in[i] > 1 ? 1 : act_in[i]);
  for (int a = 0; a < MAXARM; a++) e->grasp_cmd[a] = e->grasp[a]; /* getContactPoints as _set_action sees them */
  for (int a = 0; a < t->n_arms; a++) {
    const float* u = t->task == XARM_TASK_PUSH_WITH_DOOR ? act + 3 * a : act + 4 * a;
    Kin k;
    arm_fk(t, a, e->arm[a].q, &k);
    v3 target; double g;
    arm_command(t, a, u, k.p[m->eef_dof], e->arm[a].q[m->finger1], target, &g);
    double qn[MAXDOF];
    arm_ik(t, a, e->arm[a].q, target, qn);
    for (int i = 0; i < 7; i++) e->arm[a].qt[i] = qn[i];
    if (t->grip_cmd) {
      if (m->finger2 >= 0) { e->arm[a].qt[m->finger1] = g; e->arm[a].qt[m->finger2] = g; }
      else for (int i = m->finger1; i < m->ndof; i++) e->arm[a].qt[i] = g; /* Reach: joints 10..16 [REF xarm_reach.py:141-142] */
    }
  }
  if (t->lego_clamp) for (int i = 0; i < t->n_obj; i++) lego_clamp(&e->obj[i]);
}

typedef struct { float reward; uint8_t done, truncated; float success; } StepOut;

static void env_step(OrEnv* e, const float* action, ObsOut* o, StepOut* so) {
  const Task* t = &e->t;
  e->step_count += 1;
  set_action(e, action);
  simulate(e);
  get_obs(e, o);
  /* _is_success */
  float succ;
  if (t->task == XARM_TASK_HANDOVER) { /* [REF xarm_handover.py:395-402] */
    succ = 1;
    for (int i = 0; i < t->n_obj; i++) if (!(np_norm_f32(o->ag + 3 * i, o->dg + 3 * i, 3) < t->threshold)) succ = 0;
  } else if (t->task == XARM_TASK_PICK_AND_PLACE && t->n_obj > 1) { /* D5 */
    succ = 1;
    for (int i = 0; i < t->n_obj; i++) if (!(np_norm_f32(o->ag + 3 * i, o->dg + 3 * i, 3) < t->threshold)) succ = 0;
  } else {
    succ = np_norm_f32(o->ag, o->dg, t->goal_dim) < t->threshold ? 1.0f : 0.0f;
  }
  /* compute_reward */
  float rew;
  int rt = e->cfg.reward_type;
  if (t->task == XARM_TASK_REACH && rt == XARM_REWARD_DENSE_DIFF) { /* [REF xarm_reach.py:113-116] */
    float d = np_norm_f32(o->ag, o->dg, 3);
    rew = e->d_old - d; e->d_old = d;
  } else if ((t->task == XARM_TASK_PICK_AND_PLACE || t->task == XARM_TASK_HANDOVER) && rt == XARM_REWARD_DENSE) {
    rew = reward_dense_staged(e, o);
  } else {
    rew = reward_stateless(t->task, rt, t->n_obj, t->threshold, o->ag, o->dg, t->goal_dim);
  }
  int limit = e->cfg.max_episode_steps > 0 ? e->cfg.max_episode_steps : t->max_steps;
  int terminated = 0;
  if (t->task == XARM_TASK_PICK_AND_PLACE) terminated = np_norm_f32(o->ag, o->dg, t->goal_dim) < t->threshold; /* [REF :117] */
  if (t->task == XARM_TASK_HANDOVER) terminated = succ != 0; /* [REF xarm_handover.py:138] */
  int time_up = e->step_count >= limit; /* env's own counter or gym TimeLimit (a20, D3, D4) */
  so->reward = rew; so->success = succ;
  so->done = terminated || time_up;
  so->truncated = time_up && !terminated;
}

/* ------------------------------------------------------------------------------------------------ public API (ctypes) */
OrEnv* or_create(const XarmConfig* cfg, int64_t env_index) {
  OrEnv* e = (OrEnv*)calloc(1, sizeof(OrEnv));
  if (task_fill(&e->t, cfg->task, cfg->num_obj)) { free(e); return NULL; }
  e->cfg = *cfg; e->env_index = env_index;
  const Task* t = &e->t;
  /* constructor state: episode 0 draws place the objects [REF xarm_pick_and_place.py:73] */
  e->episode = 0; e->rng_draw = 0;
  for (int a = 0; a < t->n_arms; a++) {
    if (t->task == XARM_TASK_PICK_AND_PLACE) { memset(&e->arm[a], 0, sizeof(ArmState)); } /* no resetJointState in its ctor */
    else set_joint_init(e, a, t->task == XARM_TASK_HANDOVER ? 0.04 : 0.0);
  }
  for (int i = 0; i < t->n_obj; i++) {
    float x = 0, y = 0;
    switch (t->task) {
      case XARM_TASK_PICK_AND_PLACE: x = box_sample(e, 0.35f, 0.45f); y = box_sample(e, -0.25f, 0.25f); break;
      case XARM_TASK_STACK_TOWER: x = box_sample(e, -0.3f, 0.3f); y = box_sample(e, -0.2f, 0.2f); break;
      case XARM_TASK_PUSH_WITH_DOOR: x = box_sample(e, -0.3f, -0.1f); y = box_sample(e, -0.2f, 0.2f); break;
      case XARM_TASK_HANDOVER: x = box_sample(e, 0.11f, 0.28f); y = box_sample(e, -0.18f, 0.2f); break;
    }
    place_obj(e, i, x, y, 0.025);
  }
  sample_goal(e);
  return e;
}
void or_destroy(OrEnv* e) { free(e); }
int or_dims(int32_t task, int32_t num_obj, int32_t* a, int32_t* o, int32_t* g, int32_t* s) {
  Task t;
  if (task_fill(&t, task, num_obj)) return -1;
  *a = t.act_dim; *o = t.obs_dim; *g = t.goal_dim; *s = state_words(&t);
  return 0;
}
static void write_obs(const OrEnv* e, const ObsOut* o, float* obs, float* ag, float* dg) {
  if (obs) memcpy(obs, o->obs, sizeof(float) * e->t.obs_dim);
  if (ag) memcpy(ag, o->ag, sizeof(float) * e->t.goal_dim);
  if (dg) memcpy(dg, o->dg, sizeof(float) * e->t.goal_dim);
}
void or_reset(OrEnv* e, float* obs, float* ag, float* dg) {
  env_reset(e);
  ObsOut o; get_obs(e, &o); write_obs(e, &o, obs, ag, dg);
}
void or_get_obs(OrEnv* e, float* obs, float* ag, float* dg) { ObsOut o; get_obs(e, &o); write_obs(e, &o, obs, ag, dg); }
void or_step(OrEnv* e, const float* action, float* obs, float* ag, float* dg, float* reward, uint8_t* done, float* success, uint8_t* truncated) {
  ObsOut o; StepOut so;
  env_step(e, action, &o, &so);
  write_obs(e, &o, obs, ag, dg);
  if (reward) *reward = so.reward;
  if (done) *done = so.done;
  if (success) *success = so.success;
  if (truncated) *truncated = so.truncated;
}
/* state record: same float32 layout as xarm_get_state (DESIGN.md "state record"); _d variants keep doubles */
static int state_io(OrEnv* e, double* buf, int write_env) {
  const Task* t = &e->t; const Model* m = t->model; int n = 0;
#define IO(x) do { if (write_env) (x) = buf[n]; else buf[n] = (x); n++; } while (0)
  for (int a = 0; a < t->n_arms; a++) {
    for (int i = 0; i < m->ndof; i++) IO(e->arm[a].q[i]);
    for (int i = 0; i < m->ndof; i++) IO(e->arm[a].qd[i]);
    for (int i = 0; i < m->ndof; i++) IO(e->arm[a].qt[i]);
  }
  for (int o = 0; o < t->n_obj; o++) {
    for (int c = 0; c < 3; c++) IO(e->obj[o].pos[c]);
    for (int c = 0; c < 4; c++) IO(e->obj[o].quat[c]);
    for (int c = 0; c < 3; c++) IO(e->obj[o].v[c]);
    for (int c = 0; c < 3; c++) IO(e->obj[o].w[c]);
  }
  if (t->has_door) { IO(e->door_q); IO(e->door_qd); }
  for (int c = 0; c < t->goal_dim; c++) { if (write_env) e->goal[c] = (float)buf[n]; else buf[n] = e->goal[c]; n++; }
  if (write_env) { e->step_count = (int)buf[n]; e->episode = (uint32_t)buf[n + 1]; e->d_old = (float)buf[n + 2];
    for (int a = 0; a < 2; a++) { int g = (int)buf[n + 3 + a]; e->grasp[a] = g & 1; e->grasp_cmd[a] = (g >> 1) & 1; } } /* live | snapshot << 1 */
  else { buf[n] = e->step_count; buf[n + 1] = e->episode; buf[n + 2] = e->d_old;
    for (int a = 0; a < 2; a++) buf[n + 3 + a] = (e->grasp[a] ? 1 : 0) | (e->grasp_cmd[a] ? 2 : 0); }
  n += 5;
#undef IO
  return n;
}
void or_get_state(OrEnv* e, float* out) { double b[256]; int n = state_io(e, b, 0); for (int i = 0; i < n; i++) out[i] = (float)b[i]; }
void or_set_state(OrEnv* e, const float* in) { double b[256]; int n = state_words(&e->t); for (int i = 0; i < n; i++) b[i] = in[i]; state_io(e, b, 1); }
void or_get_state_d(OrEnv* e, double* out) { state_io(e, out, 0); }
void or_set_state_d(OrEnv* e, const double* in) { double b[256]; memcpy(b, in, sizeof(double) * state_words(&e->t)); state_io(e, b, 1); }
void or_contact_hist(OrEnv* e, long* nc_hist, long* nac_hist) {
  memcpy(nc_hist, e->st_hist_nc, sizeof(e->st_hist_nc)); memcpy(nac_hist, e->st_hist_nac, sizeof(e->st_hist_nac));
}
void or_solver_stats(OrEnv* e, long out[4]) { out[0] = e->st_substeps; out[1] = e->st_iters; out[2] = e->st_rows; out[3] = e->st_contacts; }
long or_arm_contacts(OrEnv* e, int reset) { long c = e->arm_contacts; if (reset) e->arm_contacts = 0; return c; }
double or_flops(OrEnv* e, int reset) { double f = e->flops; if (reset) e->flops = 0; return f; }

/* known-answer helpers: FK of one arm at base (0,0,0) */
int or_fk(int32_t task, const double* q, double* eef_pos, double* eef_R, double* hand_com) {
  Task t; Kin k;
  if (task_fill(&t, task, 1)) return -1;
  arm_fk(&t, 0, q, &k);
  v3cpy(eef_pos, k.p[t.model->eef_dof]);
  memcpy(eef_R, k.R[t.model->eef_dof], sizeof(m3));
  link_point(&k, t.model->eef_dof, t.model->hand_com, hand_com);
  return 0;
}
/* joint-space inverse mass matrix (row-major ndof x ndof) and ABA accelerations at (q, qd, tau): cross-checks */
int or_dynamics(int32_t task, const double* q, const double* qd, const double* tau, double* Minv, double* qdd) {
  Task t; Kin k; OrEnv e; memset(&e, 0, sizeof(e));
  if (task_fill(&t, task, 1)) return -1;
  int n = t.model->ndof;
  arm_fk(&t, 0, q, &k);
  double zero[MAXDOF] = {0}, tt[MAXDOF], col[MAXDOF];
  for (int j = 0; j < n; j++) {
    memset(tt, 0, sizeof(tt)); tt[j] = 1;
    arm_aba(&e, &t, &k, zero, tt, 0, col);
    for (int i = 0; i < n; i++) Minv[i * n + j] = col[i];
  }
  arm_aba(&e, &t, &k, qd, tau, 1, qdd);
  return n;
}
int or_ik(int32_t task, int arm, const double* q, const double* target, double* q_out) {
  Task t;
  if (task_fill(&t, task, 1)) return -1;
  arm_ik(&t, arm, q, target, q_out);
  return t.model->ndof;
}
/* box-box contact generation exposed for parity tests: boxes = [c(3) R(9,row-major) h(3)] */
int or_box_box(const double* A, const double* B, double* out /* up to 4 x [pa(3) pb(3) n(3) depth] */) {
  Collider a, b; Contact c[4];
  memset(&a, 0, sizeof(a)); memset(&b, 0, sizeof(b));
  v3cpy(a.c, A); memcpy(a.R, A + 3, sizeof(m3)); v3cpy(a.h, A + 12);
  v3cpy(b.c, B); memcpy(b.R, B + 3, sizeof(m3)); v3cpy(b.h, B + 12);
  int n = box_box(&a, &b, c, 4);
  for (int i = 0; i < n; i++) {
    v3cpy(out + 10 * i, c[i].pa); v3cpy(out + 10 * i + 3, c[i].pb); v3cpy(out + 10 * i + 6, c[i].n); out[10 * i + 9] = c[i].depth;
  }
  return n;
}
/* pure gym-logic hooks checked against tests/golden/reference_logic.npz (outputs of the reference's own methods) */
int or_debug_assemble_obs(int32_t task, int32_t num_obj, const double* hand_pos, const double* hand_vel, const double* finger_q,
                          const double* finger_qd, const double* obj /* n_obj x 13 */, const float* goal, float* obs, float* ag, float* dg) {
  Task t;
  if (task_fill(&t, task, num_obj)) return -1;
  ObsIn in; memset(&in, 0, sizeof(in));
  for (int a = 0; a < t.n_arms; a++) {
    v3cpy(in.hand_pos[a], hand_pos + 3 * a); v3cpy(in.hand_vel[a], hand_vel + 3 * a);
    in.finger_q[a] = finger_q[a]; in.finger_qd[a] = finger_qd[a];
  }
  for (int i = 0; i < t.n_obj; i++) {
    const double* b = obj + 13 * i;
    v3cpy(in.obj[i].pos, b); memcpy(in.obj[i].quat, b + 3, 4 * sizeof(double)); v3cpy(in.obj[i].v, b + 7); v3cpy(in.obj[i].w, b + 10);
  }
  memcpy(in.goal, goal, sizeof(float) * t.goal_dim);
  ObsOut o; memset(&o, 0, sizeof(o));
  assemble_obs(&t, &in, &o);
  memcpy(obs, o.obs, sizeof(float) * t.obs_dim); memcpy(ag, o.ag, sizeof(float) * t.goal_dim); memcpy(dg, o.dg, sizeof(float) * t.goal_dim);
  return 0;
}
int or_debug_command(int32_t task, int arm, const float* action /* full action vector */, const double* eef, double finger_q,
                     double* target, double* grip) {
  Task t;
  if (task_fill(&t, task, 1)) return -1;
  float act[8];
  for (int i = 0; i < t.act_dim; i++) act[i] = action[i] < -1 ? -1 : (action[i] > 1 ? 1 : action[i]);
  const float* u = t.task == XARM_TASK_PUSH_WITH_DOOR ? act + 3 * arm : act + 4 * arm;
  arm_command(&t, arm, u, eef, finger_q, target, grip);
  return 0;
}
void or_debug_lego_clamp(double* pos, double* quat) {
  ObjState s; memset(&s, 0, sizeof(s));
  v3cpy(s.pos, pos); memcpy(s.quat, quat, 4 * sizeof(double));
  lego_clamp(&s);
  v3cpy(pos, s.pos); memcpy(quat, s.quat, 4 * sizeof(double));
}
void or_philox(uint64_t seed, uint64_t env, uint32_t episode, uint32_t block, uint32_t out[4]) {
  uint32_t ctr[4] = {(uint32_t)env, (uint32_t)(env >> 32), episode, block};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  philox4x32_10(ctr, key);
  memcpy(out, ctr, sizeof(ctr));
}

/* ------------------------------------------------------------------------------------------------ CPU baseline
 * n_envs independent envs over n_threads host threads, random U(-1,1) actions, auto-reset: the stand-in for
 * "SubprocVecEnv of PyBullet envs on the host cores" (PyBullet is not installable). */
typedef struct { const XarmConfig* cfg; int64_t first, count; int steps; double env_steps; } BenchArg;
static void* bench_thread(void* p) {
  BenchArg* b = (BenchArg*)p;
  for (int64_t i = 0; i < b->count; i++) {
    OrEnv* e = or_create(b->cfg, b->first + i);
    env_reset(e);
    uint64_t s = 0x9E3779B97F4A7C15ull * (uint64_t)(b->first + i + 1);
    for (int k = 0; k < b->steps; k++) {
      float act[8]; ObsOut o; StepOut so;
      for (int c = 0; c < e->t.act_dim; c++) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        act[c] = (float)((double)(s >> 11) * (2.0 / 9007199254740992.0) - 1.0);
      }
      env_step(e, act, &o, &so);
      if (so.done) env_reset(e);
      b->env_steps += 1;
    }
    or_destroy(e);
  }
  return NULL;
}
double or_bench(const XarmConfig* cfg, int64_t n_envs, int steps, int n_threads, double* env_steps_out) {
  pthread_t th[256]; BenchArg arg[256];
  if (n_threads > 256) n_threads = 256;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int i = 0; i < n_threads; i++) {
    arg[i].cfg = cfg; arg[i].steps = steps; arg[i].env_steps = 0;
    arg[i].first = n_envs * i / n_threads; arg[i].count = n_envs * (i + 1) / n_threads - arg[i].first;
    pthread_create(&th[i], NULL, bench_thread, &arg[i]);
  }
  double total = 0;
  for (int i = 0; i < n_threads; i++) { pthread_join(th[i], NULL); total += arg[i].env_steps; }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  *env_steps_out = total;
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------------------------------------ batch stepping (tests)
 * n independent oracle envs stepped by n_threads host threads: what the statistics tests at >= 32 768 envs need (a
 * Python loop over single-env calls does ~2 k env-steps/s).  op 0: reset, 1: step, 2: get_obs. */
typedef struct {
  OrEnv** envs; int64_t first, count; int op, A, O, G;
  const float* actions; float *obs, *ag, *dg, *reward, *success; uint8_t *done, *truncated;
} BatchArg;
static void* batch_thread(void* p) {
  BatchArg* b = (BatchArg*)p;
  for (int64_t i = b->first; i < b->first + b->count; i++) {
    OrEnv* e = b->envs[i];
    float* o = b->obs ? b->obs + i * b->O : NULL; float* a = b->ag ? b->ag + i * b->G : NULL; float* d = b->dg ? b->dg + i * b->G : NULL;
    if (b->op == 0) or_reset(e, o, a, d);
    else if (b->op == 2) or_get_obs(e, o, a, d);
    else or_step(e, b->actions + i * b->A, o, a, d, b->reward ? b->reward + i : NULL, b->done ? b->done + i : NULL,
                 b->success ? b->success + i : NULL, b->truncated ? b->truncated + i : NULL);
  }
  return NULL;
}
void or_batch(OrEnv** envs, int64_t n, int op, const float* actions, float* obs, float* ag, float* dg, float* reward, uint8_t* done,
              float* success, uint8_t* truncated, int n_threads) {
  if (n <= 0) return;
  pthread_t th[256]; BatchArg arg[256];
  if (n_threads > 256) n_threads = 256;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n) n_threads = (int)n;
  const Task* t = &envs[0]->t;
  for (int i = 0; i < n_threads; i++) {
    BatchArg b = {envs, n * i / n_threads, 0, op, t->act_dim, t->obs_dim, t->goal_dim, actions, obs, ag, dg, reward, success, done, truncated};
    b.count = n * (i + 1) / n_threads - b.first;
    arg[i] = b;
    pthread_create(&th[i], NULL, batch_thread, &arg[i]);
  }
  for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
}
