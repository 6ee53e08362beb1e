// xarm_her.cuh - hindsight-experience-replay episode store and 'future' relabelling on the device
// [REF benchmark/train.py:81-97: HerReplayBuffer(n_sampled_goal=4, goal_selection_strategy="future",
//  max_episode_length=100, online_sampling=True)], semantics of stable-baselines3 1.x HerReplayBuffer (include/xarm_abi.h).
// HBM bound: add = one read of the step's outputs + one write of the same words; sample = gathered records -> dense batch.
//
// Layout: an array of step records, an episode contiguous, an env's ring of K episodes contiguous:
//   rec [N][K][T+1][RW]  RW = O + G + A + 2 words: [ obs_t (O) | achieved_goal_t (G) | action_t (A) | reward_t | done_t (0.f / 1.f) ]
//                        record t+1 starts with next_obs / next_achieved_goal of transition t (nothing is stored twice), so
//                        * add writes ONE contiguous run of RW words per env: the tail of record t (action, reward, done) and the
//                          head of record t+1 (obs, achieved goal);
//                        * a sample reads ONE contiguous run of RW + O + G words (record t and the head of record t+1) plus the
//                          future step's achieved goal - two scattered accesses instead of the eight of a structure of arrays
//                          (the first version of this file: 1.0 GB of DRAM reads per 1 M samples for 0.28 GB algorithmic).
//   dg  [K][N][G]        desired goal, one per episode: the reference's envs draw the goal in reset() only
//                        [REF xarm_pick_and_place.py:121-127; xarm_reach.py:96-102; xarm_handover.py:141-151]
//   ep_len [K][N]        transitions of a finished episode; 0 = empty or being written
//   cur_k [N], cur_t [N] ring position and length of the episode each env is writing
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct HerBuf {
  float *rec, *dg;
  int *ep_len, *cur_k, *cur_t;
  unsigned long long* counters;   // [0] invalid samples since the last read, [1] finished episodes stored, [2] transitions added
  int64_t N;
  int K, T, O, G, A;
};
#define XARM_HER_MAX_OBS 128
#define XARM_HER_MAX_ACT 32

// Record numbers are 32-bit (xarm_her_create checks N K (T+1) < 2^32): one 32-bit multiply-add and one widening multiply per
// record instead of chains of 64-bit multiplies - these kernels are issue bound on exactly that arithmetic.
XD size_t her_rec(const HerBuf& h, int k, int t, int64_t i, int RW) {
  return (size_t)(((uint32_t)i * (uint32_t)h.K + (uint32_t)k) * (uint32_t)(h.T + 1) + (uint32_t)t) * (uint32_t)RW;
}

// HerReplayBuffer.add for one transition of every env.  Inputs are the buffers an env step leaves behind: obs / ag / dg
// AFTER the step (after the auto-reset where done), terminal = [obs | ag | dg] of the finishing step (or null: no auto-reset).
// One warp per env, lanes over the RW words of the run [action | reward | done | next_obs | next_ag]; cur_k / cur_t / done are
// warp-uniform loads; every load is issued before the first store.  NR = ceil(RW / 32) (template: the loops unroll to predicated
// loads).  The bookkeeping is k_her_advance's (a second launch: every warp reads cur_k / cur_t).
template <int NR>
__global__ void __launch_bounds__(256) k_her_store(HerBuf h, const float* __restrict__ obs, const float* __restrict__ ag,
                                                   const float* __restrict__ dg, const float* __restrict__ terminal,
                                                   const float* __restrict__ action, const float* __restrict__ reward,
                                                   const uint8_t* __restrict__ done, const uint8_t* __restrict__ truncated) {
  const int O = h.O, G = h.G, A = h.A, RW = O + G + A + 2;
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));   // 32-bit arithmetic: N < 2^31
  if (i >= h.N) return;
  const int k = h.cur_k[i], t = h.cur_t[i];
  const bool d = done[i] != 0;
  const bool term = d && terminal != nullptr;
  const bool close = d || t + 1 == h.T;
  const float* trow = terminal + i * (O + 2 * G);
  const float* src_o = term ? trow : obs + i * O;
  const float* src_g = term ? trow + O : ag + i * G;
  const bool tr = truncated && truncated[i];
  float v[NR], v2[NR];
#pragma unroll
  for (int j = 0; j < NR; j++) {
    const int w = lane + 32 * j;
    float x = 0.f;
    if (w < A) x = action[i * A + w];
    else if (w == A) x = reward[i];
    else if (w == A + 1) x = (d && !tr) ? 1.f : 0.f;   // SB3 handle_timeout_termination: done * (1 - timeout)
    else if (w < A + 2 + O) x = src_o[w - A - 2];
    else if (w < RW) x = src_g[w - A - 2 - O];
    v[j] = x;
    // first record of the next episode = the observation after the auto-reset
    v2[j] = (close && w < O + G) ? (w < O ? obs[i * O + w] : ag[i * G + w - O]) : 0.f;
  }
  const float vd2 = (close && lane < G) ? dg[i * G + lane] : 0.f;
  float* run = h.rec + her_rec(h, k, t, i, RW) + (O + G);   // tail of record t, then the head of record t+1
#pragma unroll
  for (int j = 0; j < NR; j++) {
    const int w = lane + 32 * j;
    if (w < RW) run[w] = v[j];
  }
  if (!close) return;                   // the episode goes on: nothing else to store
  const int k2 = (k + 1 == h.K) ? 0 : k + 1;
  float* head = h.rec + her_rec(h, k2, 0, i, RW);
#pragma unroll
  for (int j = 0; j < NR; j++) {
    const int w = lane + 32 * j;
    if (w < O + G) head[w] = v2[j];
  }
  if (lane < G) h.dg[((int64_t)k2 * h.N + i) * G + lane] = vd2;
}

__global__ void __launch_bounds__(256) k_her_advance(HerBuf h, const uint8_t* __restrict__ done) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= h.N) return;
  const int k = h.cur_k[i], t = h.cur_t[i];
  unsigned long long eps = 0;
  if (done[i] != 0 || t + 1 == h.T) {   // an episode longer than max_episode_length is closed there (SB3 raises instead)
    const int k2 = (k + 1 == h.K) ? 0 : k + 1;
    h.ep_len[(int64_t)k * h.N + i] = t + 1;
    h.ep_len[(int64_t)k2 * h.N + i] = 0;   // the oldest episode of this env's ring is overwritten from now on
    h.cur_k[i] = k2; h.cur_t[i] = 0;
    eps = 1;
  } else {
    h.cur_t[i] = t + 1;
  }
  // statistics: one atomic per warp
  const unsigned m = __activemask();
  const unsigned e = __ballot_sync(m, eps != 0);
  if ((threadIdx.x & 31) == (__ffs(m) - 1)) {
    if (e) atomicAdd(&h.counters[1], (unsigned long long)__popc(e));
    atomicAdd(&h.counters[2], (unsigned long long)__popc(m));
  }
}

// start of an episode outside the auto-reset path (after Env.reset()): head of record 0 and the goal of the episode being written
__global__ void __launch_bounds__(256) k_her_begin(HerBuf h, const float* __restrict__ obs, const float* __restrict__ ag,
                                                   const float* __restrict__ dg, const uint8_t* __restrict__ mask) {
  const int O = h.O, G = h.G, RW = O + G + h.A + 2;
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));   // 32-bit arithmetic: N < 2^31
  if (i >= h.N) return;
  if (mask && !mask[i]) return;
  const int k = h.cur_k[i];
  if (lane == 0) h.cur_t[i] = 0;            // a partial episode is dropped
  float* head = h.rec + her_rec(h, k, 0, i, RW);
  for (int w = lane; w < O + G; w += 32) head[w] = w < O ? obs[i * O + w] : ag[i * G + w - O];
  if (lane < G) h.dg[((int64_t)k * h.N + i) * G + lane] = dg[i * G + lane];
}

// ---- sampling.  Sample b of call c draws Philox4x32-10 blocks with counter (b lo, b hi, c, try) and key = seed: word 0 -> env,
// word 1 -> ring position, word 2 -> transition, word 3 -> future transition; a try that lands on an empty / unfinished
// episode is repeated with the next block (rejection => uniform over the finished episodes, as SB3 draws episode_indices).
// Ranges are mapped by the high word of a 32 x 32-bit product.  HER samples (the first int(her_ratio * batch) of the batch,
// her_ratio = 1 - 1 / (n_sampled_goal + 1)) of an episode longer than one transition draw t in [0, L-1) and the future index
// in [t+1, L); the others - and HER samples of one-transition episodes - draw t in [0, L) and keep goal and reward.
#define XARM_HER_MAX_TRIES 64
#define XARM_HER_FUSED_MAX_BATCH 2048   // up to here one fused launch draws and gathers (xarm_her_sample)
XD int4 her_draw(const HerBuf& h, int64_t b, int64_t n_her, uint64_t seed, uint32_t call) {
  for (uint32_t tr = 0; tr < XARM_HER_MAX_TRIES; tr++) {
    uint32_t c[4] = {(uint32_t)b, (uint32_t)((uint64_t)b >> 32), call, tr};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const int env = (int)__umulhi(c[0], (uint32_t)h.N);
    const int k = (int)__umulhi(c[1], (uint32_t)h.K);
    const int L = h.ep_len[(int64_t)k * h.N + env];
    if (L <= 0) continue;
    const bool her = b < n_her && L > 1;
    const int t = (int)__umulhi(c[2], (uint32_t)(her ? L - 1 : L));
    const int tf = her ? t + 1 + (int)__umulhi(c[3], (uint32_t)(L - 1 - t)) : -1;
    return make_int4(env, k, t, tf);
  }
  atomicAdd(&h.counters[0], 1ull);
  return make_int4(-1, -1, -1, -1);
}
__global__ void __launch_bounds__(256) k_her_index(HerBuf h, int64_t batch, int64_t n_her, uint64_t seed, uint32_t call,
                                                   int4* __restrict__ index) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  index[b] = her_draw(h, b, n_her, seed, call);
}

// gather + relabel + reward: one warp per sample, lanes over the RW + O + G words of the run [record t | head of record t+1];
// the run, the future goal and the episode goal are loaded before the first store.  NR = ceil((RW + O + G) / 32).  Outputs are SB3's
// DictReplayBufferSamples fields.  FUSED (small batches, which are launch bound: SB3 trains on 256): lane 0 draws the index itself
// (same her_draw, same result) and no index kernel is launched; large batches keep the separate index pass, where 32 samples
// share a warp's Philox instructions.
template <int NR, bool FUSED>
__global__ void __launch_bounds__(256) k_her_gather(HerBuf h, int64_t batch, int4* __restrict__ index, int64_t n_her, uint64_t seed,
                                                    uint32_t call, int task, int reward_type,
                                                    int num_obj, float* __restrict__ o_obs, float* __restrict__ o_ag,
                                                    float* __restrict__ o_dg, float* __restrict__ o_act, float* __restrict__ o_nobs,
                                                    float* __restrict__ o_nag, float* __restrict__ o_rew, uint8_t* __restrict__ o_done) {
  const int O = h.O, G = h.G, A = h.A, RW = O + G + A + 2, RL = RW + O + G;
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));   // 32-bit arithmetic: batch < 2^31
  if (b >= batch) return;
  int4 s = make_int4(-1, -1, -1, -1);
  if (FUSED) {
    if (lane == 0) {
      s = her_draw(h, b, n_her, seed, call);
      if (index) index[b] = s;
    }
    s.x = __shfl_sync(0xffffffffu, s.x, 0); s.y = __shfl_sync(0xffffffffu, s.y, 0);
    s.z = __shfl_sync(0xffffffffu, s.z, 0); s.w = __shfl_sync(0xffffffffu, s.w, 0);
  } else {
    s = index[b];
  }
  const bool ok = s.x >= 0, her = ok && s.w >= 0;
  const int64_t i = ok ? (int64_t)(uint32_t)s.x : 0;
  const int k = ok ? s.y : 0, t = ok ? s.z : 0;
  const float* run = h.rec + her_rec(h, k, t, i, RW);
  float v[NR];
#pragma unroll
  for (int j = 0; j < NR; j++) {
    const int w = lane + 32 * j;
    v[j] = (ok && w < RL) ? run[w] : 0.f;
  }
  const bool lg = ok && lane < G;
  const float vna = lg ? run[RW + O + lane] : 0.f;   // next achieved goal again, in the low lanes (an L1 hit) for the reward
  // desired goal: the achieved goal of the future step for a relabelled sample
  const float vd = lg ? (her ? h.rec[her_rec(h, k, s.w, i, RW) + O + lane] : h.dg[((int64_t)k * h.N + i) * G + lane]) : 0.f;
  float vr = 0.f;
  if (her) {   // env.compute_reward(next_achieved_goal, new desired_goal, info): lane 0 collects the goals (warp-uniform branch)
    float a[9], d[9];
#pragma unroll
    for (int g = 0; g < 9; g++)
      if (g < G) { a[g] = __shfl_sync(0xffffffffu, vna, g); d[g] = __shfl_sync(0xffffffffu, vd, g); }
    if (lane == 0) {   // constant G at the call sites: the goal arrays stay in registers and the distance loops unroll
      const float thr = task_threshold(task);
      if (G == 3) vr = reward_stateless(task, reward_type, num_obj, thr, a, d, 3);
      else if (G == 6) vr = reward_stateless(task, reward_type, num_obj, thr, a, d, 6);
      else if (G == 9) vr = reward_stateless(task, reward_type, num_obj, thr, a, d, 9);
      else vr = reward_stateless(task, reward_type, num_obj, thr, a, d, G);
    }
    vr = __shfl_sync(0xffffffffu, vr, 0);
  }
#pragma unroll
  for (int j = 0; j < NR; j++) {
    const int w = lane + 32 * j;
    const float x = v[j];
    if (w < O) o_obs[b * O + w] = x;
    else if (w < O + G) o_ag[b * G + w - O] = x;
    else if (w < O + G + A) o_act[b * A + w - O - G] = x;
    else if (w == O + G + A) o_rew[b] = her ? vr : x;
    else if (w == O + G + A + 1) o_done[b] = (uint8_t)(x != 0.f);
    else if (w < RW + O) o_nobs[b * O + w - RW] = x;
    else if (w < RL) o_nag[b * G + w - RW - O] = x;
  }
  if (lane < G) o_dg[b * G + lane] = vd;
}
