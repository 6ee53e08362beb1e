// xarm_tasks.cuh - compile-time task descriptions (every number cites the reference ctor).
#pragma once
#include "../../include/xarm_abi.h"
#include "xarm_model.cuh"

// body codes used by contacts
enum : int {
  BC_STATIC = 0,
  BC_ARM0_HAND = 1, BC_ARM0_F1 = 2, BC_ARM0_F2 = 3,
  BC_ARM1_HAND = 4, BC_ARM1_F1 = 5, BC_ARM1_F2 = 6,
  BC_OBJ0 = 7, BC_OBJ1 = 8, BC_OBJ2 = 9,
  BC_DOOR = 10
};
XHD bool bc_is_arm(int c) { return c >= BC_ARM0_HAND && c <= BC_ARM1_F2; }
XHD int bc_arm(int c) { return c >= BC_ARM1_HAND ? 1 : 0; }
XHD bool bc_is_obj(int c) { return c >= BC_OBJ0 && c <= BC_OBJ2; }

#define XARM_MAXC XARM_MAX_CONTACTS
#define XARM_MAXAC XARM_MAX_ARM_CONTACTS

template <int TASK, int NOBJ_>
struct TaskT;

// XarmReachEnv [REF xarm_reach.py:15-35]
template <>
struct TaskT<XARM_TASK_REACH, 0> {
  using MD = ModelXG;
  static constexpr int TASK = XARM_TASK_REACH, NARM = 1, NOBJ = 0, NTABLE = 1, A = 4, O = 8, G = 3, MAXC = 4;
  static constexpr bool HAS_DOOR = false, HAS_GROUND = false, DAMP_EACH = false, GRIP_CMD = true, GRIP_CLIP = false;
  static constexpr bool FRICTION_SWITCH = false, LEGO_CLAMP = false, FINGER_TABLE = false;
  static constexpr int NSUB = 20, NIK = 20, MAX_STEPS = 25;
  static constexpr double TIME_STEP = 1. / 240, H = TIME_STEP / 20, DT_CMD = TIME_STEP * 20;
  static constexpr double MAX_VEL = 1, MAX_GRIP_VEL = 20, ARM_FORCE = 5 * 240., FINGER_FORCE = 5 * 240.;
  static constexpr float GRIP_LO = 0, GRIP_HI = 0, THRESHOLD = 0.05f;
  static constexpr float OBJ_HX = 0, OBJ_HY = 0, OBJ_HZ = 0, OBJ_MASS = 1, HAND_OFF_Z = 0;
  static XHD float base_x(int) { return 0.f; }
  static XHD float pos_lo(int, int c) { return c == 0 ? 0.2f : (c == 1 ? -0.4f : 0.2f); }
  static XHD float pos_hi(int, int c) { return c == 0 ? 0.8f : (c == 1 ? 0.4f : 0.6f); }
  static XHD float table_x(int) { return 0.f; }
};

// XarmPickAndPlace [REF xarm_pick_and_place.py:17-50]
template <int NOBJ_>
struct TaskT<XARM_TASK_PICK_AND_PLACE, NOBJ_> {
  using MD = ModelPD;
  static constexpr int TASK = XARM_TASK_PICK_AND_PLACE, NARM = 1, NOBJ = NOBJ_, NTABLE = 1, A = 4, O = 8 + 16 * NOBJ_, G = 3 * NOBJ_;
  static constexpr int MAXC = NOBJ_ == 1 ? 16 : XARM_MAXC;  // one lego: table + finger1 + finger2 + hand manifolds, 4 points each
  static constexpr bool HAS_DOOR = false, HAS_GROUND = false, DAMP_EACH = false, GRIP_CMD = true, GRIP_CLIP = true;
  static constexpr bool FRICTION_SWITCH = true, LEGO_CLAMP = false, FINGER_TABLE = false;
  static constexpr int NSUB = 15, NIK = 15, MAX_STEPS = 50;
  static constexpr double TIME_STEP = 1. / 60, H = TIME_STEP / 15, DT_CMD = TIME_STEP * 15;
  static constexpr double MAX_VEL = 0.25, MAX_GRIP_VEL = 0.08, ARM_FORCE = XARM_MOTOR_DEFAULT_FORCE, FINGER_FORCE = 1000;
  static constexpr float GRIP_LO = 0.01f, GRIP_HI = 0.04f, THRESHOLD = 0.05f;
  static constexpr float OBJ_HX = 0.025f, OBJ_HY = 0.025f, OBJ_HZ = 0.04f, OBJ_MASS = 0.5f, HAND_OFF_Z = 0;
  static XHD float base_x(int) { return 0.f; }
  static XHD float pos_lo(int, int c) { return c == 0 ? 0.3f : (c == 1 ? -0.3f : 0.15f); }
  static XHD float pos_hi(int, int c) { return c == 0 ? 0.5f : (c == 1 ? 0.3f : 0.4f); }
  static XHD float table_x(int) { return 0.f; }
};

// shared by XarmStackTowerEnv [REF xarm_stack_tower.py:14-43] and XarmPushWithDoorEnv [REF xarm_push_with_door.py:14-42]
struct TwoArmTable {
  using MD = ModelPD;
  static constexpr int NARM = 2, NTABLE = 1, MAXC = XARM_MAXC;
  static constexpr bool HAS_GROUND = false, DAMP_EACH = false, GRIP_CLIP = true;
  static constexpr bool FRICTION_SWITCH = false, LEGO_CLAMP = false, FINGER_TABLE = false;
  static constexpr int NSUB = 15, NIK = 15, MAX_STEPS = 50;
  static constexpr double TIME_STEP = 1. / 60, H = TIME_STEP / 15, DT_CMD = TIME_STEP * 15;
  static constexpr double MAX_VEL = 0.25, MAX_GRIP_VEL = 1, ARM_FORCE = XARM_MOTOR_DEFAULT_FORCE, FINGER_FORCE = XARM_MOTOR_DEFAULT_FORCE;
  static constexpr float GRIP_LO = 0.021f, GRIP_HI = 0.04f;
  static constexpr float OBJ_HX = 0.025f, OBJ_HY = 0.025f, OBJ_HZ = 0.025f, OBJ_MASS = 0.1f, HAND_OFF_Z = 0;
  static XHD float base_x(int a) { return a == 0 ? -0.6f : 0.6f; }
  static XHD float pos_lo(int a, int c) { return c == 0 ? (a == 0 ? -0.4f : -0.3f) : (c == 1 ? -0.3f : 0.125f); }
  static XHD float pos_hi(int a, int c) { return c == 0 ? (a == 0 ? 0.3f : 0.4f) : (c == 1 ? 0.3f : 0.4f); }
  static XHD float table_x(int) { return 0.f; }
};
template <>
struct TaskT<XARM_TASK_STACK_TOWER, 3> : TwoArmTable {
  static constexpr int TASK = XARM_TASK_STACK_TOWER, NOBJ = 3, A = 8, O = 55, G = 9;
  static constexpr bool HAS_DOOR = false, GRIP_CMD = true;
  static constexpr float THRESHOLD = (float)(0.03 * 3);
};
template <>
struct TaskT<XARM_TASK_PUSH_WITH_DOOR, 1> : TwoArmTable {
  static constexpr int TASK = XARM_TASK_PUSH_WITH_DOOR, NOBJ = 1, A = 6, O = 25, G = 3;
  static constexpr bool HAS_DOOR = true, GRIP_CMD = false;  // D1: the reference's finger commands raise NameError
  static constexpr float THRESHOLD = (float)(0.03 * 1);
};

// XarmHandover [REF xarm_handover.py:25-57,79-83]
template <int NOBJ_>
struct TaskT<XARM_TASK_HANDOVER, NOBJ_> {
  using MD = ModelPD;
  static constexpr int TASK = XARM_TASK_HANDOVER, NARM = 2, NOBJ = NOBJ_, NTABLE = 2, A = 8, O = 13 * NOBJ_ + 16, G = 3 * NOBJ_, MAXC = XARM_MAXC;
  static constexpr bool HAS_DOOR = false, HAS_GROUND = true, DAMP_EACH = true, GRIP_CMD = true, GRIP_CLIP = true;
  static constexpr bool FRICTION_SWITCH = true, LEGO_CLAMP = true, FINGER_TABLE = true;
  static constexpr int NSUB = 15, NIK = 15, MAX_STEPS = 100;
  static constexpr double TIME_STEP = 1. / 240, H = TIME_STEP, DT_CMD = TIME_STEP * 15;
  static constexpr double MAX_VEL = 1.8, MAX_GRIP_VEL = 1, ARM_FORCE = XARM_MOTOR_DEFAULT_FORCE, FINGER_FORCE = XARM_MOTOR_DEFAULT_FORCE;
  static constexpr float GRIP_LO = 0.020f, GRIP_HI = 0.04f, THRESHOLD = 0.05f;
  static constexpr float OBJ_HX = 0.075f, OBJ_HY = 0.025f, OBJ_HZ = 0.025f, OBJ_MASS = 0.5f;
  static constexpr float HAND_OFF_Z = (float)(0.088 - 0.021);
  static XHD float base_x(int a) { return a == 0 ? -0.6f : 0.6f; }
  static XHD float pos_lo(int a, int c) { return c == 0 ? (a == 0 ? -0.3f : 0.0f) : (c == 1 ? -0.2f : 0.1f); }
  static XHD float pos_hi(int a, int c) { return c == 0 ? (a == 0 ? 0.0f : 0.3f) : (c == 1 ? 0.2f : 0.22f); }
  static XHD float table_x(int k) { return k == 0 ? -0.85f : 0.85f; }
};

template <class T>
XHD int state_words() {
  return T::NARM * 3 * T::MD::N + T::NOBJ * 13 + (T::HAS_DOOR ? 2 : 0) + T::G + 5;
}
