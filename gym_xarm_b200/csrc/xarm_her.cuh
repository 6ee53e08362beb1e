// xarm_her.cuh - hindsight-experience-replay episode store and 'future' relabelling on the device
// [REF benchmark/train.py:81-97: HerReplayBuffer(n_sampled_goal=4, goal_selection_strategy="future",
//  max_episode_length=100, online_sampling=True)], semantics of stable-baselines3 1.x HerReplayBuffer (include/xarm_abi.h).
// HBM bound: add = one read of the step's outputs + one write of the same words; sample = gathered rows -> dense batch.
//
// Layout (time-major, so that the envs of a slab, which step together, write neighbouring addresses):
//   obs [K][T+1][N][O]   row t = observation before transition t; row t+1 = observation after it (next_obs of t)
//   ag  [K][T+1][N][G]   achieved goals, same indexing
//   act [K][T][N][A], rew [K][T][N], done [K][T][N] (uint8)
//   dg  [K][N][G]        desired goal, one per episode: the reference's envs draw the goal in reset() only
//                        [REF xarm_pick_and_place.py:121-127; xarm_reach.py:96-102; xarm_handover.py:141-151]
//   ep_len [K][N]        transitions of a finished episode; 0 = empty or being written
//   cur_k [N], cur_t [N] ring position and length of the episode each env is writing
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct HerBuf {
  float *obs, *ag, *dg, *act, *rew;
  uint8_t* done;
  int *ep_len, *cur_k, *cur_t;
  unsigned long long* counters;   // [0] invalid samples since the last read, [1] finished episodes stored, [2] transitions added
  int64_t N;
  int K, T, O, G, A;
};

// Row numbers are 32-bit (xarm_her_create checks K (T+1) N < 2^32): one 32-bit multiply-add per row and one widening multiply per
// array instead of chains of 64-bit multiplies - the kernels are issue bound on exactly this arithmetic.
XD size_t her_row(const HerBuf& h, int k, int t, int64_t i) { return (uint32_t)(k * (h.T + 1) + t) * (uint32_t)h.N + (uint32_t)i; }   // obs / ag rows
XD size_t her_tr(const HerBuf& h, int k, int t, int64_t i) { return (uint32_t)(k * h.T + t) * (uint32_t)h.N + (uint32_t)i; }          // act / rew / done rows

// HerReplayBuffer.add for one transition of every env.  Inputs are the buffers an env step leaves behind: obs / ag / dg
// AFTER the step (after the auto-reset where done), terminal = [obs | ag | dg] of the finishing step (or null: no auto-reset).
// One warp per env, lanes over the words of a row (no index arithmetic beyond the row base; cur_k / cur_t / done are
// warp-uniform loads); every load of the row is issued before the first store.  The bookkeeping is k_her_advance's (a second
// launch: every warp reads cur_k / cur_t).
// HER_OBS_REGS = ceil(O / 32) rounded up to 1, 2 or 4: observation words a lane carries (template: the row loops unroll to
// exactly that many predicated loads; the first version unrolled 4 for every O and was issue bound at 195 instructions per warp).
#define XARM_HER_MAX_OBS 128
template <int HER_OBS_REGS>
__global__ void __launch_bounds__(256) k_her_store(HerBuf h, const float* __restrict__ obs, const float* __restrict__ ag,
                                                   const float* __restrict__ dg, const float* __restrict__ terminal,
                                                   const float* __restrict__ action, const float* __restrict__ reward,
                                                   const uint8_t* __restrict__ done, const uint8_t* __restrict__ truncated) {
  const int O = h.O, G = h.G, A = h.A;
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
  float vo[HER_OBS_REGS], vo2[HER_OBS_REGS];
#pragma unroll
  for (int j = 0; j < HER_OBS_REGS; j++) {
    const int w = lane + 32 * j;
    vo[j] = w < O ? src_o[w] : 0.f;
    vo2[j] = (close && w < O) ? obs[i * O + w] : 0.f;
  }
  const float vg = lane < G ? src_g[lane] : 0.f;
  const float va = lane < A ? action[i * A + lane] : 0.f;
  const float vg2 = (close && lane < G) ? ag[i * G + lane] : 0.f;
  const float vd2 = (close && lane < G) ? dg[i * G + lane] : 0.f;
  const float vr = reward[i];
  const bool tr = truncated && truncated[i];
  const size_t r1 = her_row(h, k, t + 1, i), r0 = her_tr(h, k, t, i);
#pragma unroll
  for (int j = 0; j < HER_OBS_REGS; j++) {
    const int w = lane + 32 * j;
    if (w < O) h.obs[r1 * O + w] = vo[j];
  }
  if (lane < G) h.ag[r1 * G + lane] = vg;
  if (lane < A) h.act[r0 * A + lane] = va;
  if (lane == 0) {
    h.rew[r0] = vr;
    h.done[r0] = (uint8_t)(d && !tr);   // SB3 handle_timeout_termination: done * (1 - timeout)
  }
  if (!close) return;                   // the episode goes on: nothing else to store
  const int k2 = (k + 1 == h.K) ? 0 : k + 1;  // first row of the next episode = the observation after the auto-reset
  const size_t r2 = her_row(h, k2, 0, i);
#pragma unroll
  for (int j = 0; j < HER_OBS_REGS; j++) {
    const int w = lane + 32 * j;
    if (w < O) h.obs[r2 * O + w] = vo2[j];
  }
  if (lane < G) {
    h.ag[r2 * G + lane] = vg2;
    h.dg[((int64_t)k2 * h.N + i) * G + lane] = vd2;
  }
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

// start of an episode outside the auto-reset path (after Env.reset()): row 0 and the goal of the episode being written
__global__ void __launch_bounds__(256) k_her_begin(HerBuf h, const float* __restrict__ obs, const float* __restrict__ ag,
                                                   const float* __restrict__ dg, const uint8_t* __restrict__ mask) {
  const int O = h.O, G = h.G;
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));   // 32-bit arithmetic: N < 2^31
  if (i >= h.N) return;
  if (mask && !mask[i]) return;
  const int k = h.cur_k[i];
  if (lane == 0) h.cur_t[i] = 0;            // a partial episode is dropped
  const size_t r = her_row(h, k, 0, i);
  for (int w = lane; w < O; w += 32) h.obs[r * O + w] = obs[i * O + w];
  if (lane < G) {
    h.ag[r * G + lane] = ag[i * G + lane];
    h.dg[((int64_t)k * h.N + i) * G + lane] = dg[i * G + lane];
  }
}

// ---- sampling.  Sample b of call c draws Philox4x32-10 blocks with counter (b lo, b hi, c, try) and key = seed: word 0 -> env,
// word 1 -> ring position, word 2 -> transition, word 3 -> future transition; a try that lands on an empty / unfinished
// episode is repeated with the next block (rejection => uniform over the finished episodes, as SB3 draws episode_indices).
// Ranges are mapped by the high word of a 32 x 32-bit product.  HER samples (the first int(her_ratio * batch) of the batch,
// her_ratio = 1 - 1 / (n_sampled_goal + 1)) of an episode longer than one transition draw t in [0, L-1) and the future index
// in [t+1, L); the others - and HER samples of one-transition episodes - draw t in [0, L) and keep goal and reward.
#define XARM_HER_MAX_TRIES 64
__global__ void __launch_bounds__(256) k_her_index(HerBuf h, int64_t batch, int64_t n_her, uint64_t seed, uint32_t call,
                                                   int4* __restrict__ index) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  int4 r = make_int4(-1, -1, -1, -1);
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
    r = make_int4(env, k, t, tf);
    break;
  }
  if (r.x < 0) atomicAdd(&h.counters[0], 1ull);
  index[b] = r;
}

// gather + relabel + reward: one warp per sample, lanes over the words of a row; the eight row segments are loaded before the
// first store (eight independent gathers in flight per warp).  Outputs are SB3's DictReplayBufferSamples fields.
template <int HER_OBS_REGS>
__global__ void __launch_bounds__(256) k_her_gather(HerBuf h, int64_t batch, const int4* __restrict__ index, int task, int reward_type,
                                                    int num_obj, float* __restrict__ o_obs, float* __restrict__ o_ag,
                                                    float* __restrict__ o_dg, float* __restrict__ o_act, float* __restrict__ o_nobs,
                                                    float* __restrict__ o_nag, float* __restrict__ o_rew, uint8_t* __restrict__ o_done) {
  const int O = h.O, G = h.G, A = h.A;
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));   // 32-bit arithmetic: batch < 2^31
  if (b >= batch) return;
  const int4 s = index[b];
  const bool ok = s.x >= 0, her = ok && s.w >= 0;
  const int64_t i = ok ? (int64_t)(uint32_t)s.x : 0;
  const int k = ok ? s.y : 0, t = ok ? s.z : 0;
  const size_t r0 = her_row(h, k, t, i), r1 = her_row(h, k, t + 1, i), q = her_tr(h, k, t, i);
  float vo[HER_OBS_REGS], vn[HER_OBS_REGS];
#pragma unroll
  for (int j = 0; j < HER_OBS_REGS; j++) {
    const int w = lane + 32 * j;
    vo[j] = (ok && w < O) ? h.obs[r0 * O + w] : 0.f;
    vn[j] = (ok && w < O) ? h.obs[r1 * O + w] : 0.f;
  }
  const bool lg = ok && lane < G;
  const float va = lg ? h.ag[r0 * G + lane] : 0.f;
  const float vna = lg ? h.ag[r1 * G + lane] : 0.f;
  // desired goal: the achieved goal of the future step for a relabelled sample
  const float vd = lg ? (her ? h.ag[her_row(h, k, s.w, i) * G + lane] : h.dg[((int64_t)k * h.N + i) * G + lane]) : 0.f;
  const float vact = (ok && lane < A) ? h.act[q * A + lane] : 0.f;
  float vr = ok ? h.rew[q] : 0.f;
  const uint8_t vdone = ok ? h.done[q] : (uint8_t)0;
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
  }
#pragma unroll
  for (int j = 0; j < HER_OBS_REGS; j++) {
    const int w = lane + 32 * j;
    if (w < O) { o_obs[b * O + w] = vo[j]; o_nobs[b * O + w] = vn[j]; }
  }
  if (lane < G) { o_ag[b * G + lane] = va; o_nag[b * G + lane] = vna; o_dg[b * G + lane] = vd; }
  if (lane < A) o_act[b * A + lane] = vact;
  if (lane == 0) { o_rew[b] = vr; o_done[b] = vdone; }
}
