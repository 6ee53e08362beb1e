// Microbenchmark (development aid): the arm-row PGS sweep of xarm_sim.cuh in isolation - 9 unit rows + gear, forward /
// backward alternation, everything in registers.  Answers: what IPC does this loop reach when nothing else competes
// for the instruction cache?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o pgs_micro pgs_micro.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define N 9
__host__ __device__ constexpr int tri(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }
template <int REGS_BLOCKS>
__global__ void __launch_bounds__(128, REGS_BLOCKS) k(const float* __restrict__ in, float* __restrict__ out, int n, int sweeps) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  float Mi[45], iden[N], dqd[N], rhs[N], app[N];
#pragma unroll
  for (int i = 0; i < 45; i++) Mi[i] = in[(size_t)i * n + t];
#pragma unroll
  for (int i = 0; i < N; i++) { iden[i] = 1.f / Mi[tri(i, i)]; rhs[i] = in[(size_t)(45 + i) * n + t]; dqd[i] = 0.f; app[i] = 0.f; }
  const float hi = 1e3f, sthr = 3.16e-4f;
  unsigned long long ok = 0;
#define ROW(i) { float delta = rhs[i] - dqd[i] * iden[i]; const float sum = app[i] + delta; const float sc = fminf(fmaxf(sum, -hi), hi); \
    delta = (sc == sum) ? delta : sc - app[i]; app[i] = sc; _Pragma("unroll") for (int k_ = 0; k_ < N; k_++) dqd[k_] += Mi[tri(k_, i)] * delta; \
    bad = bad || fabsf(delta) > sthr * iden[i]; }
  for (int it = 0; it < sweeps; it++) {
    bool bad = false;
    if (it & 1) { ROW(0) ROW(1) ROW(2) ROW(3) ROW(4) ROW(5) ROW(6) ROW(7) ROW(8) }
    else { ROW(8) ROW(7) ROW(6) ROW(5) ROW(4) ROW(3) ROW(2) ROW(1) ROW(0) }
    if (!bad) ok |= 1ull << (it & 63);
  }
#pragma unroll
  for (int i = 0; i < N; i++) out[(size_t)i * n + t] = dqd[i] + (float)(ok & 1);
}
int main() {
  const int n = 131072, sweeps = 750;
  std::vector<float> h((size_t)54 * n);
  for (int t = 0; t < n; t++) {
    for (int i = 0; i < N; i++) for (int j = 0; j <= i; j++) h[(size_t)tri(i, j) * n + t] = (i == j ? 2.0f + 0.1f * i : 0.3f / (1 + i + j));
    for (int i = 0; i < N; i++) h[(size_t)(45 + i) * n + t] = 0.01f * (i + 1) * ((t % 7) - 3);
  }
  float *din, *dout;
  cudaMalloc(&din, h.size() * 4); cudaMalloc(&dout, (size_t)N * n * 4);
  cudaMemcpy(din, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int variant = 0; variant < 2; variant++) {
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
      // variant 0: 16 warps/SM (registers allow 4 blocks); variant 1: occupancy forced to 2 blocks = 8 warps/SM by 100 KB of dynamic smem
      if (variant == 0) k<4><<<n / 128, 128>>>(din, dout, n, sweeps); else k<4><<<n / 128, 128, 100 * 1024>>>(din, dout, n, sweeps);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double instr = (double)n * sweeps * 9 * 18;  // ~18 SASS instructions per row
      printf("warps/SM=%d: %.3f ms  (~%.2f IPC/SM at 1.9 GHz assuming 18 instr/row)\n", variant == 0 ? 16 : 8, ms, instr / 32 / (ms * 1e-3) / 1.9e9 / 148);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
