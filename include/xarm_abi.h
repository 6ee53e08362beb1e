/* xarm_abi.h - C ABI of libxarm_b200.so: the drop-in boundary for the gym-xarm environment step.
 *
 * The reference has no FFI of its own: its envs are Python classes that call the third-party `pybullet` C-API
 * ~35-60 times per step.  This ABI replaces that whole call sequence with one batched call per
 * gym method; each entry point cites the reference method it stands for.  Plain pointers and sizes only - no
 * torch types.  Device pointers are owned by the caller (torch allocates them; addresses stay fixed so the
 * step can be replayed as a CUDA graph).  Every function returns 0 on success or a negative XARM_E_* code and
 * leaves a message for xarm_last_error(); nothing throws or aborts across the boundary.
 *
 * Threading: a handle is not thread-safe; distinct handles (one per GPU / per process) are independent.
 * xarm_step/xarm_reset are asynchronous on the given stream; the *_host variants synchronise that stream. */
#ifndef XARM_ABI_H
#define XARM_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XARM_ABI_VERSION 2

/* tasks: the five env classes of the hot path */
enum {
  XARM_TASK_REACH = 0,           /* XarmReachEnv        [REF gym_xarm/envs/xarm_reach.py:9]           */
  XARM_TASK_PICK_AND_PLACE = 1,  /* XarmPickAndPlace    [REF gym_xarm/envs/xarm_pick_and_place.py:16] */
  XARM_TASK_STACK_TOWER = 2,     /* XarmStackTowerEnv   [REF gym_xarm/envs/xarm_stack_tower.py:13]    */
  XARM_TASK_PUSH_WITH_DOOR = 3,  /* XarmPushWithDoorEnv [REF gym_xarm/envs/xarm_push_with_door.py:13] */
  XARM_TASK_HANDOVER = 4,        /* XarmHandover        [REF gym_xarm/envs/xarm_handover.py:19]       */
  XARM_NUM_TASKS = 5
};

/* reward types: config['reward_type'] strings of the reference */
enum {
  XARM_REWARD_SPARSE = 0,     /* all tasks */
  XARM_REWARD_DENSE = 1,      /* Reach: -d; PickAndPlace/Handover: staged (reads sim state); Stack/Push: -d */
  XARM_REWARD_DENSE_O2G = 2,  /* PickAndPlace only: -d   [REF xarm_pick_and_place.py:176-177] */
  XARM_REWARD_DENSE_DIFF = 3  /* Reach only: d_old - d   [REF xarm_reach.py:113-116] */
};

enum { XARM_GOAL_AIR = 0, XARM_GOAL_GROUND = 1 };

enum {
  XARM_OK = 0,
  XARM_E_INVALID = -1,   /* bad argument / unsupported config (Python side raises ValueError/NotImplementedError) */
  XARM_E_CUDA = -2,      /* CUDA runtime error (sticky errors surface at the next call) */
  XARM_E_STATE = -3,     /* call order violated (e.g. step before bind) */
  XARM_E_NOMEM = -4
};

/* mirrors the `config` dict of the reference constructors [REF xarm_pick_and_place.py:17,53,68,164,261,272,276;
 * xarm_handover.py:25,60,85,380,387,391; xarm_reach.py:25] plus the batching/seeding keys this repo adds */
typedef struct XarmConfig {
  int32_t task;               /* XARM_TASK_* */
  int32_t reward_type;        /* XARM_REWARD_* */
  int32_t num_obj;            /* config['num_obj'] (PickAndPlace, Handover); fixed 3 / 1 for StackTower / PushWithDoor */
  int32_t goal_shape;         /* config['goal_shape']: XARM_GOAL_AIR | XARM_GOAL_GROUND */
  float init_grasp_rate;      /* config['init_grasp_rate'] */
  float goal_ground_rate;     /* config['goal_ground_rate'] */
  float same_side_rate;       /* config['same_side_rate'] */
  int32_t use_stand;          /* config['use_stand']: must be 0 - xarm_create refuses 1 (the stand collider is not built) */
  int32_t max_episode_steps;  /* 0 = the registered TimeLimit (25/50/50/50/100) [REF gym_xarm/__init__.py:6-22] */
  int32_t auto_reset;         /* 1: VecEnv semantics - a finished env is reset inside xarm_step */
  int32_t device;             /* CUDA device ordinal */
  int32_t stagger_phases;     /* 1: xarm_create / an explicit xarm_reset start env i at step counter (global index mod episode
                                 length), so that time-limit endings - and with them the auto-reset work - spread evenly over the
                                 steps instead of arriving as one wave every episode length (no reference counterpart: SubprocVecEnv
                                 workers drift apart by themselves; off in every parity test) */
  int64_t num_envs;           /* envs in this slab */
  int64_t env_index_base;     /* global index of env 0 of the slab: RNG streams are keyed by the global index */
  uint64_t seed;
} XarmConfig;

/* Caller-owned DEVICE buffers (float32 unless noted), row-major [num_envs, dim]. */
typedef struct XarmBuffers {
  const float* actions;    /* [N, A]  in   */
  float* observation;      /* [N, O]  out: obs['observation'] (after auto-reset: first obs of the new episode) */
  float* achieved_goal;    /* [N, G]  out */
  float* desired_goal;     /* [N, G]  out */
  float* reward;           /* [N]     out */
  uint8_t* done;           /* [N]     out: terminated or truncated */
  float* success;          /* [N]     out: info['is_success'] */
  uint8_t* truncated;      /* [N]     out: info['TimeLimit.truncated'] */
  float* terminal_observation; /* [N, O+2G] out (may be NULL): observation|achieved|desired of the step BEFORE auto-reset */
} XarmBuffers;

typedef struct XarmHandle XarmHandle;

/* dims of a task: action A, observation O, goal G, state words per env S (layout in DESIGN.md "state record") */
int xarm_task_dims(int32_t task, int32_t num_obj, int32_t* act_dim, int32_t* obs_dim, int32_t* goal_dim, int32_t* state_words);

/* Env.__init__(config)  [REF xarm_pick_and_place.py:17-103; xarm_reach.py:15-77; xarm_handover.py:25-124;
 * xarm_stack_tower.py:14-97; xarm_push_with_door.py:14-99] - allocates the SoA state slab on cfg->device. */
int xarm_create(const XarmConfig* cfg, XarmHandle** out);
/* Env.close()  [REF xarm_reach.py:125-126, xarm_handover.py:147-148] */
int xarm_destroy(XarmHandle* h);
int xarm_bind(XarmHandle* h, const XarmBuffers* bufs);

/* Env.reset()  [REF xarm_pick_and_place.py:121-127,250-287; xarm_reach.py:96-102,163-173;
 * xarm_stack_tower.py:115-119,201-219; xarm_push_with_door.py:117-121,195-212; xarm_handover.py:141-145,338-393].
 * mask: device uint8[N] selecting envs (NULL = all).  Writes observation/achieved_goal/desired_goal. */
int xarm_reset(XarmHandle* h, const uint8_t* mask, void* stream);

/* Env.step(action)  [REF xarm_pick_and_place.py:107-119,199-248; xarm_reach.py:81-94,131-161;
 * xarm_stack_tower.py:101-113,142-199; xarm_push_with_door.py:103-115,144-193; xarm_handover.py:128-139,244-336]
 * for every env of the slab: _set_action (IK + motor targets + grasp friction), stepSimulation x n_substeps,
 * _get_obs, _is_success, compute_reward, done, TimeLimit, optional auto-reset.  Replays the captured CUDA graph
 * when xarm_graph_capture() succeeded for this stream. */
int xarm_step(XarmHandle* h, void* stream);

/* Same step with HOST buffers (the path SB3's numpy VecEnv uses): copies actions H2D, steps (replaying a CUDA graph of its
 * own, captured at the first call), copies observation|achieved|desired|reward|done|success|truncated D2H, synchronises.
 * NULL outputs are skipped.  terminal_observation [N, O + 2 G] (may be NULL): rows of the envs that finished in this step
 * receive observation|achieved|desired of the step BEFORE the auto-reset (SB3's infos[i]['terminal_observation']); other
 * rows are left untouched.  Only the finished rows cross PCIe (gathered on the device).  stream NULL = the handle's own. */
int xarm_step_host(XarmHandle* h, const float* actions, float* observation, float* achieved_goal, float* desired_goal,
                   float* reward, uint8_t* done, float* success, uint8_t* truncated, float* terminal_observation, void* stream);
int xarm_reset_host(XarmHandle* h, float* observation, float* achieved_goal, float* desired_goal, void* stream);

/* Env.compute_reward(achieved_goal, desired_goal, info) for HER relabelling: batch form, device pointers
 * [REF xarm_reach.py:107-116; xarm_pick_and_place.py:155-190; xarm_stack_tower.py:124-129;
 *  xarm_push_with_door.py:126-131; xarm_handover.py:153-183].  Only the state-free reward types. */
int xarm_compute_reward(int32_t task, int32_t reward_type, int32_t num_obj, const float* achieved_goal,
                        const float* desired_goal, int64_t n, float* out, void* stream);

/* Simulator checkpoint (the reference has none; needed for parity injection): HOST float32
 * [num_envs, state_words] records. */
int xarm_get_state(XarmHandle* h, float* host_out);
int xarm_set_state(XarmHandle* h, const float* host_in);
/* recompute observation/achieved_goal/desired_goal from the current state (_get_obs) */
int xarm_get_obs(XarmHandle* h, void* stream);

/* Capture xarm_step (step + auto-reset) into a CUDA graph on `stream`; later xarm_step calls on that stream replay it. */
int xarm_graph_capture(XarmHandle* h, void* stream);

/* Episode statistics accumulated on the device since the last call (reset on read):
 * out[0]=episodes, out[1]=sum return, out[2]=sum length, out[3]=sum success, out[4]=diverged (NaN-guard resets). */
int xarm_episode_stats(XarmHandle* h, double out[5], void* stream);

/* Profiling aid (bench.py): device-side %globaltimer stamps around every launch of the step pipeline, accumulated per
 * kernel.  xarm_set_profiling invalidates a captured graph (capture again).  xarm_kernel_times writes text lines
 * "branch kernel launches total_us" (branch M = main, E = early branch + its auto-reset passes, L = late tail) into
 * out[cap] and returns the number of lines.  [no reference counterpart: PyBullet has no per-stage timers] */
int xarm_set_profiling(XarmHandle* h, int32_t on);
int xarm_kernel_times(XarmHandle* h, char* out, int64_t cap);
/* Measured FP32 SIMT peak of `device` in TFLOP/s (the roofline this path is bound by is FP32 issue, and
 * MEASURED_PEAKS.json has no FP32 figure): 8 independent FMA chains per thread at full occupancy, best of 5 launches of
 * ~10 ms, CUDA events.  Synchronises the device. */
int xarm_measure_fp32_peak(int32_t device, double* tflops);
/* ---- caller side of the path: SB3 VecExtractDictObs + VecNormalize on the device
 * [REF benchmark/train.py:44-62,74-75: make_vec_env -> VecNormalize(env, norm_obs=True, norm_reward=True, clip_obs=10.)].
 * Semantics of stable-baselines3 1.x VecNormalize, pinned by the reference's saved benchmark/saved_data/
 * XarmPDHandoverNoGoal-v1/vec_normalize.pkl (epsilon 1e-8, gamma 0.99, clip 10/10, RunningMeanStd count 1e-4 at start;
 * after 6250 steps of 4 envs obs_rms.count = 25000.0001 and ret_rms.count = 25004.0001: reset() feeds the zeroed
 * returns to ret_rms and does not touch obs_rms).  Running statistics are float64 on the device. */
typedef struct XarmVecNormConfig {
  int64_t num_envs;
  int32_t obs_dim;        /* flat observation (obs['observation']) */
  int32_t device;
  float gamma;            /* 0.99 */
  float clip_obs;         /* 10 */
  float clip_reward;      /* 10 */
  float epsilon;          /* 1e-8 */
  int32_t norm_obs, norm_reward, training, reserved;
} XarmVecNormConfig;
typedef struct XarmVecNorm XarmVecNorm;
int xarm_vecnorm_create(const XarmVecNormConfig* cfg, XarmVecNorm** out);
int xarm_vecnorm_destroy(XarmVecNorm* v);
/* VecNormalize.reset(): returns := 0, ret_rms.update(returns) when training, obs_out = normalize_obs(obs).  Device pointers. */
int xarm_vecnorm_reset(XarmVecNorm* v, const float* obs, float* obs_out, void* stream);
/* VecNormalize.step_wait() after the env step: obs_rms.update(obs), obs_out = clip((obs - mean) / sqrt(var + eps)),
 * returns = returns * gamma + reward, ret_rms.update(returns), reward_out = clip(reward / sqrt(ret var + eps)),
 * returns[done] = 0.  obs [N, obs_dim], reward [N], done uint8 [N]: device pointers; outputs may alias the inputs. */
int xarm_vecnorm_step(XarmVecNorm* v, const float* obs, const float* reward, const uint8_t* done, float* obs_out,
                      float* reward_out, void* stream);
int xarm_vecnorm_set_training(XarmVecNorm* v, int32_t training);
/* VecNormalize.normalize_obs / unnormalize_obs (inverse = 1) and normalize_reward on ARBITRARY device batches of n rows: apply
 * only - neither the running statistics nor the discounted returns change.  Rows may be strided (row_stride / out_stride floats
 * between rows, >= obs_dim): the observation part of a terminal-observation slab [N, O + 2 G] normalises without a copy. */
int xarm_vecnorm_normalize_obs(XarmVecNorm* v, const float* obs, int64_t n, int64_t row_stride, float* out, int64_t out_stride,
                               int32_t inverse, void* stream);
int xarm_vecnorm_normalize_reward(XarmVecNorm* v, const float* reward, int64_t n, float* out, void* stream);
/* obs_rms / ret_rms (HOST float64: mean[obs_dim], var[obs_dim], count; mean, var, count) - VecNormalize.save / load */
int xarm_vecnorm_get_stats(XarmVecNorm* v, double* obs_mean, double* obs_var, double* obs_count, double* ret_stats3);
int xarm_vecnorm_set_stats(XarmVecNorm* v, const double* obs_mean, const double* obs_var, double obs_count, const double* ret_stats3);

/* ---- caller side of the path: hindsight experience replay on the device
 * [REF benchmark/train.py:81-97: HerReplayBuffer(n_sampled_goal=4, goal_selection_strategy="future",
 *  max_episode_length=100, online_sampling=True), whose relabelling calls env.compute_reward on batches].
 * Semantics of stable-baselines3 1.x HerReplayBuffer ('future' strategy, online sampling), with three stated differences:
 * (1) every env owns a ring of `episodes_per_env` episode slots (SB3: one global ring; it supports a single env only);
 * (2) indices come from Philox4x32-10 keyed by (seed; sample, call, try) instead of np.random, so a batch is a pure function
 *     of the buffer and the call number (oracle/her_oracle.py restates it);
 * (3) the desired goal is kept once per episode (the reference's envs draw it in reset() only).
 * All data pointers are DEVICE pointers. */
typedef struct XarmHerConfig {
  int64_t num_envs;
  int32_t episodes_per_env;     /* ring slots per env (>= 2: one is always being written) */
  int32_t max_episode_length;   /* T: longer episodes are closed at T transitions */
  int32_t obs_dim, goal_dim, action_dim;
  int32_t task, reward_type, num_obj;   /* for compute_reward: state-free reward types only (as xarm_compute_reward) */
  int32_t n_sampled_goal;       /* her_ratio = 1 - 1 / (n_sampled_goal + 1) */
  int32_t device;
  uint64_t seed;
} XarmHerConfig;
typedef struct XarmHer XarmHer;
int xarm_her_create(const XarmHerConfig* cfg, XarmHer** out);
int xarm_her_destroy(XarmHer* h);
/* after Env.reset(): first row and goal of the episode each (masked) env starts; an unfinished episode is dropped */
int xarm_her_begin(XarmHer* h, const float* observation, const float* achieved_goal, const float* desired_goal,
                   const uint8_t* mask_or_null, void* stream);
/* HerReplayBuffer.add after Env.step: observation / achieved_goal / desired_goal [N, O | G | G] as the step left them (after
 * the auto-reset where done), terminal [N, O + 2 G] = the finishing step's [obs | ag | dg] (XarmBuffers.terminal_observation;
 * NULL without auto-reset), action [N, A], reward [N], done [N] u8, truncated [N] u8 or NULL (stored done = done & !truncated). */
int xarm_her_add(XarmHer* h, const float* observation, const float* achieved_goal, const float* desired_goal,
                 const float* terminal, const float* action, const float* reward, const uint8_t* done,
                 const uint8_t* truncated, void* stream);
/* HerReplayBuffer.sample(batch): observation, next_observation [B, O]; achieved_goal, next_achieved_goal, desired_goal [B, G];
 * action [B, A]; reward [B]; done [B] u8; index_or_null [B, 4] int32 = (env, ring slot, transition, future transition or -1).
 * Samples that found no finished episode in 64 draws are zero-filled with env = -1 and counted (xarm_her_stats). */
int xarm_her_sample(XarmHer* h, int64_t batch, float* observation, float* achieved_goal, float* desired_goal, float* action,
                    float* next_observation, float* next_achieved_goal, float* reward, uint8_t* done, int32_t* index_or_null,
                    void* stream);
/* out[0] = invalid samples since the last call, out[1] = finished episodes stored so far, out[2] = transitions added,
 * out[3] = sample calls so far.  Synchronises the device. */
int xarm_her_stats(XarmHer* h, int64_t out[4]);

/* kernels launched by this library since load (claim for bench.py's gpu_launches) */
int64_t xarm_launch_count(void);
const char* xarm_last_error(void);
int xarm_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* XARM_ABI_H */
