// Microbenchmark (development aid): latency of ONE row step of the impulse-space joint loop (xarm_heavy.cuh, heavy_solve_dela)
// for a lone warp: x = c - s * dinv -> clamp -> delta -> broadcast from the owner lane -> s += A * delta.  Variants of the
// broadcast: SHFL.IDX (width 16), REDUX.OR over the half-warp's member mask, a shared-memory word, and none (lower bound).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_micro chain_micro.cu && ./chain_micro
#include <cuda_runtime.h>
#include <cstdio>

template <int MODE>
__global__ void k(float* out, long long* cyc, int steps, float dinv, float a0, float a1, float a2, float a3) {
  __shared__ float bc[2];
  const int l = threadIdx.x & 15;
  float s0 = 0.01f * l, s1 = 0.02f * l, s2 = 0.03f, s3 = 0.04f, c = 0.5f + l, app = 0.f;
  __syncwarp();
  const long long t0 = clock64();
  for (int it = 0; it < steps; it++) {
    const int o = it & 15;
    const float x = fmaf(-s3, dinv, c);
    const float xn = fminf(fmaxf(x, -1.f), 1.f);
    const float delta = xn - app;
    float d;
    if (MODE == 0) d = __shfl_sync(0xffffffffu, delta, o, 16);
    else if (MODE == 1) {
      const unsigned m = (threadIdx.x & 16) ? 0xffff0000u : 0x0000ffffu;
      d = __uint_as_float(__reduce_or_sync(m, l == o ? __float_as_uint(delta) : 0u));
    } else if (MODE == 2) {
      if (l == o) bc[threadIdx.x >> 4] = delta;
      __syncwarp();
      d = bc[threadIdx.x >> 4];
      __syncwarp();
    } else d = delta;
    const bool own = l == o;
    app = own ? xn : app; c = app + 0.25f;
    s0 = fmaf(a0, d, s0); s1 = fmaf(a1, d, s1); s2 = fmaf(a2, d, s2); s3 = fmaf(a3, d, s3);
  }
  const long long t1 = clock64();
  out[threadIdx.x] = s0 + s1 + s2 + s3 + app;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// plain dependent-issue probes: a chain of FFMAs, a chain of SHFLs, FMNMX pairs
__global__ void k_probe(float* out, long long* cyc, int steps, float a, float b) {
  float x = threadIdx.x * 0.001f;
  long long t0 = clock64();
  for (int i = 0; i < steps; i++) x = fmaf(x, a, b);
  long long t1 = clock64();
  float y = x;
  for (int i = 0; i < steps; i++) y = __shfl_sync(0xffffffffu, y, (i + 1) & 15, 16);
  long long t2 = clock64();
  float z = y;
  for (int i = 0; i < steps; i++) z = fminf(fmaxf(z * a, -b), b);
  long long t3 = clock64();
  out[threadIdx.x] = z;
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
}
int main() {
  float* out; long long* cyc, h[3];
  cudaMalloc(&out, 256); cudaMalloc(&cyc, 64);
  const int steps = 20000;
  const char* names[4] = {"SHFL.IDX width 16", "REDUX.OR half-warp masks", "shared-memory word", "no broadcast (bound)"};
  for (int rep = 0; rep < 2; rep++) {
    k<0><<<1, 32>>>(out, cyc, steps, 0.5f, 0.1f, 0.2f, 0.3f, 0.05f); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("%-28s %.1f cycles per row step\n", names[0], (double)h[0] / steps);
    k<1><<<1, 32>>>(out, cyc, steps, 0.5f, 0.1f, 0.2f, 0.3f, 0.05f); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("%-28s %.1f cycles per row step\n", names[1], (double)h[0] / steps);
    k<2><<<1, 32>>>(out, cyc, steps, 0.5f, 0.1f, 0.2f, 0.3f, 0.05f); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("%-28s %.1f cycles per row step\n", names[2], (double)h[0] / steps);
    k<3><<<1, 32>>>(out, cyc, steps, 0.5f, 0.1f, 0.2f, 0.3f, 0.05f); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("%-28s %.1f cycles per row step\n", names[3], (double)h[0] / steps);
    k_probe<<<1, 32>>>(out, cyc, steps, 0.999f, 0.5f); cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost);
    if (rep) printf("probes: FFMA chain %.1f | SHFL chain %.1f | FMUL+FMNMX+FMNMX chain %.1f cycles per link\n", (double)h[0] / steps, (double)h[1] / steps, (double)h[2] / steps);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
