/* xarm_constants.h - every PyBullet/Bullet default the env step depends on, in one place.
 *
 * The reference delegates all physics to the third-party `pybullet` wheel (unpinned, [REF setup.py:17]); no Bullet
 * source exists in the build container, so the values below are RECALLED from Bullet3 (era pybullet 3.1-3.2) and
 * are the calibration knobs of this repo.  The one value pinned by a reference artefact
 * is the multibody linear damping (G1: terminal free-fall velocity -15.1604595 m/s).
 * Shared by the CPU oracle (oracle/xarm_oracle.c, double) and the CUDA kernels (float): constants only, no logic. */
#ifndef XARM_CONSTANTS_H
#define XARM_CONSTANTS_H

/* ---- world (PhysicsServerCommandProcessor::createEmptyDynamicsWorld) ---- */
#define XARM_GRAVITY 9.8                 /* [REF xarm_pick_and_place.py:110] setGravity(0,0,-9.8) */
#ifndef XARM_SOLVER_ITERATIONS
#define XARM_SOLVER_ITERATIONS 50        /* numSolverIterations default */
#endif
#define XARM_RESIDUAL_THRESHOLD 1e-7     /* leastSquaresResidualThreshold: exit when max (delta v)^2 <= this */
#define XARM_ERP 0.2                     /* btContactSolverInfo::m_erp (joint-limit rows) */
#define XARM_ERP2 0.08                   /* m_erp2 as set by pybullet (contact rows) */
#define XARM_LINEAR_SLOP 1e-5            /* m_linearSlop as set by pybullet */
#define XARM_MAX_FRICTION 10.0           /* combined lateral friction clamp (MAX_FRICTION) */
#define XARM_DEFAULT_FRICTION 0.5        /* lateral friction of a body without <contact> */
#define XARM_TABLE_FRICTION 1.0          /* pybullet_data table/table.urdf <lateral_friction value="1.0"/> */
#define XARM_TWO_FRICTION_DIRS 1         /* SOLVER_USE_2_FRICTION_DIRECTIONS with the implicit cone (uncertain: switch) */
#define XARM_CONTACT_WARMSTART 0         /* multibody contact warm starting is disabled in btMultiBodyConstraintSolver */
#define XARM_CONTACT_MAX_IMPULSE 1e10
#define XARM_CONTACT_MARGIN 0.0          /* boxes collide when they overlap (btBoxBoxDetector: no speculative points) */
#define XARM_MAX_CONTACTS 24             /* per env and substep; pairs are visited in a fixed order, overflow is dropped */
#define XARM_MAX_ARM_CONTACTS 12         /* of which at most this many touch a gripper link */

/* ---- btMultiBody ---- */
#define XARM_MB_LINEAR_DAMPING 0.04      /* m_linearDamping: F = m v (k + k|v|)  -- pinned by G1 */
#define XARM_MB_ANGULAR_DAMPING 0.04     /* m_angularDamping: T = I w (k + k|w|) */
#define XARM_MB_USE_GYRO 0               /* m_useGyroTerm(false): the w x (I w) term is left out */
#define XARM_ANGULAR_MOTION_THRESHOLD 0.7853981633974483 /* 0.5*SIMD_HALF_PI: clamp on |w| h in the quaternion update */

/* ---- joint motors (setJointMotorControl2, POSITION_CONTROL) ---- */
#define XARM_MOTOR_KP 0.1                /* positionGain default */
#define XARM_MOTOR_KD 1.0                /* velocityGain default */
#define XARM_MOTOR_ERP 1.0
#define XARM_MOTOR_DEFAULT_FORCE 100000.0 /* pybullet.c default `force` when not passed */
#define XARM_DEFAULT_MOTOR_MAX_IMPULSE 1.0 /* loader-created velocity motor (target 0) on never-commanded joints */
#define XARM_LIMIT_MAX_IMPULSE 100.0     /* btMultiBodyConstraint::m_maxAppliedImpulse default (joint-limit rows) */

/* ---- gear constraint [REF xarm_pick_and_place.py:78-79] changeConstraint(gearRatio=-1, erp=0.1, maxForce=50) ---- */
#define XARM_GEAR_RATIO (-1.0)
#define XARM_GEAR_ERP 0.1
#define XARM_GEAR_MAX_FORCE 50.0

/* ---- Panda finger <contact> block [REF xarm7_pd.urdf:343-349,369-375] ---- */
#define XARM_FINGER_STIFFNESS 30000.0
#define XARM_FINGER_DAMPING 1000.0
#define XARM_FINGER_FRICTION_FREE 1.0    /* [REF xarm_pick_and_place.py:217-218] */
#define XARM_FINGER_FRICTION_GRASP 100.0 /* [REF xarm_pick_and_place.py:214-215] */

/* ---- IK (calculateInverseKinematics -> IKTrajectoryHelper IK2_VEL_DLS_WITH_ORIENTATION) ---- */
#define XARM_IK_DAMPING 0.5              /* per-joint damping added to the diagonal of J^T J */
#define XARM_IK_MAX_STEP 0.7853981633974483 /* MaxAngleDLS = 45 deg */
#define XARM_IK_RESIDUAL 1e-4            /* residualThreshold on the position error */

/* ---- scene geometry: pybullet_data table/table.urdf loaded at z=-0.625 => top face is z=0 over 1.5 x 1.0 m
 *      (witness: [REF gym_xarm/envs/urdf/my_table.urdf:22-30]) ---- */
#define XARM_TABLE_HALF_X 0.75
#define XARM_TABLE_HALF_Y 0.5
#define XARM_TABLE_HALF_Z 0.025
#define XARM_GROUND_Z (-0.625)           /* plane.urdf at z=-0.625 [REF xarm_handover.py:79] */

/* ---- RNG contract: Philox4x32-10, key=(seed lo, seed hi), counter=(env lo, env hi, episode, draw) */
#define XARM_PHILOX_M0 0xD2511F53u
#define XARM_PHILOX_M1 0xCD9E8D57u
#define XARM_PHILOX_W0 0x9E3779B9u
#define XARM_PHILOX_W1 0xBB67AE85u
#define XARM_GOAL_RESAMPLE_TRIES 32      /* bound on the reference's rejection loops */

#endif /* XARM_CONSTANTS_H */
