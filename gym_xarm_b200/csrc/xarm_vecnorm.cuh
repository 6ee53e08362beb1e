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
__global__ void __launch_bounds__(256, 6) k_vn_moments(VnStats* st, const float* __restrict__ obs, int64_t n, int O, float* __restrict__ ret,
                                                    const float* __restrict__ reward, float gamma, int do_obs, int ret_mode) {
  // per-warp accumulators (shared-memory double atomics are CAS loops: keep the contention to the <= 5 lanes of a warp that
  // share a column), combined by the first 2 O + 2 threads
  constexpr int SW = 2 * XARM_VN_MAX_OBS + 2;
  __shared__ double shw[8][SW];
  double* sh = shw[threadIdx.x >> 5];
  for (int k = threadIdx.x; k < 8 * SW; k += blockDim.x) (&shw[0][0])[k] = 0.0;
  __syncthreads();
  const int64_t T = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (do_obs) {
    // 4 consecutive elements per thread and iteration (one 128-bit load); T * 4 is a multiple of O, so the four columns
    // of a thread are fixed.  float partial sums over chunks of 8 iterations, folded into doubles (shifted data: small)
    const int64_t total = n * O, q0 = t0 * 4;
    int col[4];
    float m0[4];
    double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; k++) { col[k] = (int)((q0 + k) % O); m0[k] = (float)st->mean[col[k]]; }
    float p1[4] = {0, 0, 0, 0}, p2[4] = {0, 0, 0, 0};
    int it = 0, cnt[4] = {0, 0, 0, 0};
    const int64_t step = T * 4;
    int64_t e = q0;
    // main part: two 128-bit loads in flight per thread
    for (; e + step + 3 < total; e += 2 * step) {
      const float4 va = *reinterpret_cast<const float4*>(obs + e);
      const float4 vb = *reinterpret_cast<const float4*>(obs + e + step);
      const float xa[4] = {va.x, va.y, va.z, va.w}, xb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float da = xa[k] - m0[k], db = xb[k] - m0[k];
        p1[k] += da + db; p2[k] += da * da + db * db;
        cnt[k] += 2;
      }
      if ((++it & 3) == 0) {
#pragma unroll
        for (int k = 0; k < 4; k++) { s1[k] += (double)p1[k]; s2[k] += (double)p2[k]; p1[k] = 0.f; p2[k] = 0.f; }
      }
    }
    for (; e < total; e += step) {   // tail
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (e + k < total) { const float d = obs[e + k] - m0[k]; p1[k] += d; p2[k] += d * d; cnt[k] += 1; }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      s1[k] += (double)p1[k]; s2[k] += (double)p2[k];
      // the data were shifted by float(mean): move the sums to the double mean the finalize pass assumes
      const double c = (double)m0[k] - st->mean[col[k]];
      const double t1 = s1[k] + (double)cnt[k] * c, t2 = s2[k] + 2.0 * c * s1[k] + (double)cnt[k] * c * c;
      atomicAdd(&sh[col[k]], t1); atomicAdd(&sh[O + col[k]], t2);
    }
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
  for (int k = threadIdx.x; k < 2 * O + 2; k += blockDim.x) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) v += shw[w][k];
    if (v != 0.0) atomicAdd(&st->acc[k], v);
  }
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
    const int64_t total = n * O, q0 = t0 * 4;   // four fixed columns per thread: T * 4 is a multiple of O
    float m[4], iv[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { const int col = (int)((q0 + k) % O); m[k] = sm[col]; iv[k] = sm[O + col]; }
    const int64_t step = T * 4;
    int64_t e = q0;
#define VN_NORM4(v, o)                                                                                            \
    if (norm_obs) {                                                                                               \
      o.x = fminf(fmaxf((v.x - m[0]) * iv[0], -clip_obs), clip_obs); o.y = fminf(fmaxf((v.y - m[1]) * iv[1], -clip_obs), clip_obs); \
      o.z = fminf(fmaxf((v.z - m[2]) * iv[2], -clip_obs), clip_obs); o.w = fminf(fmaxf((v.w - m[3]) * iv[3], -clip_obs), clip_obs); \
    }
    for (; e + step + 3 < total; e += 2 * step) {   // two 128-bit loads in flight per thread
      const float4 va = *reinterpret_cast<const float4*>(obs + e);
      const float4 vb = *reinterpret_cast<const float4*>(obs + e + step);
      float4 oa = va, ob = vb;
      VN_NORM4(va, oa) VN_NORM4(vb, ob)
      *reinterpret_cast<float4*>(obs_out + e) = oa;
      *reinterpret_cast<float4*>(obs_out + e + step) = ob;
    }
    for (; e < total; e += step) {
      if (e + 3 < total) {
        const float4 v = *reinterpret_cast<const float4*>(obs + e);
        float4 o = v;
        VN_NORM4(v, o)
        *reinterpret_cast<float4*>(obs_out + e) = o;
      } else {
        for (int k = 0; k < 4 && e + k < total; k++) { const float x = obs[e + k]; obs_out[e + k] = norm_obs ? fminf(fmaxf((x - m[k]) * iv[k], -clip_obs), clip_obs) : x; }
      }
    }
#undef VN_NORM4
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


// VecNormalize.normalize_obs / unnormalize_obs on any batch of rows (apply only; rows may be strided)
__global__ void __launch_bounds__(256) k_vn_rows(const VnStats* __restrict__ st, const float* __restrict__ src, int64_t n, int O, int64_t src_stride,
                                                 float* __restrict__ dst, int64_t dst_stride, float clip_obs, float eps, int norm_obs, int inverse) {
  const int64_t total = n * O;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / O;
    const int c = (int)(e - r * O);
    const float x = src[r * src_stride + c];
    float y = x;
    if (norm_obs) {
      if (!inverse) y = fminf(fmaxf((x - st->fmean[c]) * st->finv[c], -clip_obs), clip_obs);
      else y = (float)((double)x * sqrt(st->var[c] + (double)eps) + st->mean[c]);
    }
    dst[r * dst_stride + c] = y;
  }
}
__global__ void __launch_bounds__(256) k_vn_reward_rows(const VnStats* __restrict__ st, const float* __restrict__ src, int64_t n, float* __restrict__ dst,
                                                        float clip_reward, int norm_reward) {
  const float rinv = st->rinv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = norm_reward ? fminf(fmaxf(src[i] * rinv, -clip_reward), clip_reward) : src[i];
}
