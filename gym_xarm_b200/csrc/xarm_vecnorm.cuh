// xarm_vecnorm.cuh - VecExtractDictObs + VecNormalize of the reference's training script on the device
// [REF benchmark/train.py:44-62,74-75], semantics of stable-baselines3 1.x (include/xarm_abi.h).  HBM bound: one read
// of the observation batch for the moments, one read (L2) + one write for the normalised batch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define XARM_VN_MAX_OBS 128
struct VnStats {          // device, float64
  double mean[XARM_VN_MAX_OBS], var[XARM_VN_MAX_OBS], count;   // obs_rms
  double rmean, rvar, rcount;                                  // ret_rms
  double acc[2 * XARM_VN_MAX_OBS + 2];                         // batch sums (shifted by the running mean) and sums of squares
  float fmean[XARM_VN_MAX_OBS], finv[XARM_VN_MAX_OBS], rinv;   // what the apply pass reads
};

// batch moments of obs (per column) and, for step, of the updated returns.  The launch uses a thread count that is a
// multiple of obs_dim, so a thread's column is fixed over its grid-stride loop and its two sums stay in registers.
__global__ void __launch_bounds__(256) k_vn_moments(VnStats* st, const float* __restrict__ obs, int64_t n, int O, float* __restrict__ ret,
                                                    const float* __restrict__ reward, float gamma, int do_obs, int ret_mode) {
  __shared__ double sh[2 * XARM_VN_MAX_OBS + 2];
  for (int k = threadIdx.x; k < 2 * O + 2; k += blockDim.x) sh[k] = 0.0;
  __syncthreads();
  const int64_t T = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (do_obs) {
    const int col = (int)(t0 % O);
    const double m0 = st->mean[col];
    double s1 = 0.0, s2 = 0.0;
    for (int64_t e = t0; e < n * O; e += T) { const double d = (double)obs[e] - m0; s1 += d; s2 += d * d; }
    atomicAdd(&sh[col], s1); atomicAdd(&sh[O + col], s2);
  }
  if (ret_mode) {  // 1: returns = returns * gamma + reward (step); 2: returns = 0 (reset)
    double s1 = 0.0, s2 = 0.0;
    for (int64_t i = t0; i < n; i += T) {
      const float r = ret_mode == 1 ? ret[i] * gamma + reward[i] : 0.f;
      ret[i] = r;
      s1 += (double)r; s2 += (double)r * (double)r;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { s1 += __shfl_down_sync(0xffffffffu, s1, off); s2 += __shfl_down_sync(0xffffffffu, s2, off); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&sh[2 * O], s1); atomicAdd(&sh[2 * O + 1], s2); }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 2 * O + 2; k += blockDim.x) if (sh[k] != 0.0) atomicAdd(&st->acc[k], sh[k]);
}

// RunningMeanStd.update_from_moments for every column and for the returns; refresh what the apply pass reads
__global__ void k_vn_finalize(VnStats* st, int64_t n, int O, float eps, int upd_obs, int upd_ret) {
  const int k = threadIdx.x;
  if (k < O) {
    if (upd_obs) {
      const double bm_shift = st->acc[k] / (double)n;                       // batch mean - running mean
      const double bvar = st->acc[O + k] / (double)n - bm_shift * bm_shift; // population variance of the batch
      const double cnt = st->count, tot = cnt + (double)n;
      const double delta = bm_shift;
      const double new_mean = st->mean[k] + delta * (double)n / tot;
      const double M2 = st->var[k] * cnt + bvar * (double)n + delta * delta * cnt * (double)n / tot;
      st->mean[k] = new_mean; st->var[k] = M2 / tot;
    }
    st->fmean[k] = (float)st->mean[k];
    st->finv[k] = (float)(1.0 / sqrt(st->var[k] + (double)eps));
    st->acc[k] = 0.0; st->acc[O + k] = 0.0;
  }
  __syncthreads();
  if (k == 0) {
    if (upd_obs) st->count += (double)n;
    if (upd_ret) {
      const double bmean = st->acc[2 * O] / (double)n, bvar = st->acc[2 * O + 1] / (double)n - bmean * bmean;
      const double cnt = st->rcount, tot = cnt + (double)n, delta = bmean - st->rmean;
      const double M2 = st->rvar * cnt + bvar * (double)n + delta * delta * cnt * (double)n / tot;
      st->rmean += delta * (double)n / tot; st->rvar = M2 / tot; st->rcount = tot;
    }
    st->rinv = (float)(1.0 / sqrt(st->rvar + (double)eps));
    st->acc[2 * O] = 0.0; st->acc[2 * O + 1] = 0.0;
  }
}

// normalize_obs / normalize_reward, returns[done] = 0
__global__ void __launch_bounds__(256) k_vn_apply(const VnStats* __restrict__ st, const float* __restrict__ obs, float* __restrict__ obs_out, int64_t n, int O,
                                                  const float* __restrict__ reward, float* __restrict__ reward_out, const uint8_t* __restrict__ done,
                                                  float* __restrict__ ret, float clip_obs, float clip_reward, int norm_obs, int norm_reward) {
  __shared__ float sm[2 * XARM_VN_MAX_OBS];
  for (int k = threadIdx.x; k < O; k += blockDim.x) { sm[k] = st->fmean[k]; sm[O + k] = st->finv[k]; }
  __syncthreads();
  const int64_t T = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (obs_out) {
    const int col = (int)(t0 % O);   // fixed per thread: T is a multiple of O
    const float m = sm[col], iv = sm[O + col];
    for (int64_t e = t0; e < n * O; e += T) {
      const float x = obs[e];
      obs_out[e] = norm_obs ? fminf(fmaxf((x - m) * iv, -clip_obs), clip_obs) : x;
    }
  }
  if (reward_out) {
    const float rinv = st->rinv;
    for (int64_t i = t0; i < n; i += T) {
      const float r = reward[i];
      reward_out[i] = norm_reward ? fminf(fmaxf(r * rinv, -clip_reward), clip_reward) : r;
      if (done[i]) ret[i] = 0.f;
    }
  }
}
