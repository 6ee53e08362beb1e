// Microbenchmark (development aid): the unit-row block of heavy_solve_dela (10 static-owner row steps per sweep) and one contact
// (normal + friction step) as written in xarm_heavy.cuh, for a lone warp, with pieces switched off one at a time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o unit_micro unit_micro.cu && ./unit_micro
#include <cuda_runtime.h>
#include <cstdio>
#define FULL 0xffffffffu
// MODE bits: 1 = Delassus column from shared memory (else registers), 2 = residual ("bad") bookkeeping, 4 = contact rows too,
// 8 = no normal rows, 16 = no friction rows, 32 = friction apply as an FMA chain, 64 = rsqrt.approx.ftz on max(l2, tiny),
// 128 = friction: broadcast the un-scaled (sticking) deltas first, a corrective exchange only when an owner lane's pair leaves the cone
template <int MODE>
__global__ void k(float* out, long long* cyc, int sweeps, float uden, float uhi, float sthr, int nc, float mu_in) {
  __shared__ float A[2][64 * 64];
  const int l = threadIdx.x & 15, g = threadIdx.x >> 4;
  float* Ag = A[g];
  for (int i = threadIdx.x & 15; i < 64 * 64; i += 16) Ag[i] = 1e-3f * ((i * 7) % 13) - 5e-3f;
  __syncwarp();
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.01f * l, mapp = 0.f, urhs = 0.1f + 0.01f * l, cu = urhs;
  float appn = 0.f, app1 = 0.f, app2 = 0.f, cn = 0.2f, ca = 0.01f, cb = -0.02f, dinv0 = 0.7f, di1 = 0.6f, di2 = 0.5f, cfmr0 = 0.f, rhs0 = 0.2f, rhs1 = 0.01f, rhs2 = -0.02f, mu = mu_in;
  const float uden_t = uden;
  if (MODE & 8) appn = 1.f;   // friction rows alone: a normal impulse for the cone to refer to
  float rc[10][4];
#pragma unroll
  for (int k2 = 0; k2 < 10; k2++) { rc[k2][0] = Ag[k2 * 64 + l]; rc[k2][1] = Ag[k2 * 64 + l + 16]; rc[k2][2] = Ag[k2 * 64 + l + 32]; rc[k2][3] = Ag[k2 * 64 + l + 48]; }
  const long long t0 = clock64();
  bool anybad = false;
  for (int it = 0; it < sweeps; it++) {
    bool bad = false;
#pragma unroll
    for (int o = 0; o < 10; o++) {
      const float x_ = fmaf(-s3, uden, cu);
      const float xn_ = fminf(fmaxf(x_, -uhi), uhi);
      const float delta = xn_ - mapp;
      const float d_ = __shfl_sync(FULL, delta, o, 16);
      const bool own_ = l == o;
      mapp = own_ ? xn_ : mapp; cu = mapp + urhs;
      if (MODE & 2) bad = bad || (own_ && fabsf(delta) > sthr * uden_t);
      if (MODE & 1) { const float* c_ = Ag + o * 64 + l; s0 += c_[0] * d_; s1 += c_[16] * d_; s2 += c_[32] * d_; s3 += c_[48] * d_; }
      else { s0 += rc[o][0] * d_; s1 += rc[o][1] * d_; s2 += rc[o][2] * d_; s3 += rc[o][3] * d_; }
    }
    if (MODE & 4) {
if (!(MODE & 8))
#pragma unroll 2
      for (int c = 0; c < nc; c++) {
        const float* c_ = Ag + (10 + 3 * c) * 64 + l;
        const float a0 = c_[0], a1 = c_[16], a2 = c_[32], a3 = c_[48];
        const float x_ = fmaf(-s0, dinv0, cn);
        const float xn_ = fminf(fmaxf(x_, 0.f), 1e10f);
        const float d0 = xn_ - appn;
        const float d_ = __shfl_sync(FULL, d0, c, 16);
        const bool own_ = l == c;
        appn = own_ ? xn_ : appn; cn = fmaf(-appn, cfmr0, appn + rhs0);
        bad = bad || (own_ && fabsf(d0) > sthr * dinv0);
        s0 = fmaf(a0, d_, s0); s1 = fmaf(a1, d_, s1); s2 = fmaf(a2, d_, s2); s3 = fmaf(a3, d_, s3);
      }
      const float lim = mu * appn, lim2 = lim * lim;
if (!(MODE & 16))
#pragma unroll 2
      for (int c = 0; c < nc; c++) {
        const float* c_ = Ag + (10 + 3 * c + 1) * 64 + l;
        const float p0 = c_[0], p1 = c_[16], p2 = c_[32], p3 = c_[48], q0 = c_[64], q1 = c_[80], q2 = c_[96], q3 = c_[112];
        float xa = fmaf(-s1, di1, ca), xb = fmaf(-s2, di2, cb);
        if (MODE & 128) {
          const float l2s = xa * xa + xb * xb;
          const bool own_ = l == c;
          const bool slide = own_ && l2s > lim2;
          const float da = xa - app1, db = xb - app2;
          const float da_ = __shfl_sync(FULL, da, c, 16), db_ = __shfl_sync(FULL, db, c, 16);
          float na = own_ ? xa : app1, nb = own_ ? xb : app2;
          s0 = fmaf(q0, db_, fmaf(p0, da_, s0)); s1 = fmaf(q1, db_, fmaf(p1, da_, s1)); s2 = fmaf(q2, db_, fmaf(p2, da_, s2)); s3 = fmaf(q3, db_, fmaf(p3, da_, s3));
          if (__any_sync(FULL, slide)) {
            float r_; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r_) : "f"(fmaxf(l2s, 1e-30f)));
            const float sc = lim * r_;
            const float xs = slide ? xa * sc : xa, ys = slide ? xb * sc : xb;
            const float ea = slide ? xs - xa : 0.f, eb = slide ? ys - xb : 0.f;
            const float ea_ = __shfl_sync(FULL, ea, c, 16), eb_ = __shfl_sync(FULL, eb, c, 16);
            if (own_) { na = xs; nb = ys; }
            s0 = fmaf(q0, eb_, fmaf(p0, ea_, s0)); s1 = fmaf(q1, eb_, fmaf(p1, ea_, s1)); s2 = fmaf(q2, eb_, fmaf(p2, ea_, s2)); s3 = fmaf(q3, eb_, fmaf(p3, ea_, s3));
          }
          bad = bad || (own_ && (fabsf(na - app1) > sthr * di1 || fabsf(nb - app2) > sthr * di2));
          app1 = na; app2 = nb; ca = app1 + rhs1; cb = app2 + rhs2;
          continue;
        }
        const float l2 = xa * xa + xb * xb;
        float sc;
        if (MODE & 64) { float r_; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r_) : "f"(fmaxf(l2, 1e-30f))); sc = l2 > lim2 ? lim * r_ : 1.f; }
        else sc = l2 > lim2 ? lim * rsqrtf(l2) : 1.f;
        xa *= sc; xb *= sc;
        const float da = xa - app1, db = xb - app2;
        const float da_ = __shfl_sync(FULL, da, c, 16), db_ = __shfl_sync(FULL, db, c, 16);
        const bool own_ = l == c;
        app1 = own_ ? xa : app1; app2 = own_ ? xb : app2; ca = app1 + rhs1; cb = app2 + rhs2;
        bad = bad || (own_ && (fabsf(da) > sthr * di1 || fabsf(db) > sthr * di2));
        if (MODE & 32) { s0 = fmaf(q0, db_, fmaf(p0, da_, s0)); s1 = fmaf(q1, db_, fmaf(p1, da_, s1)); s2 = fmaf(q2, db_, fmaf(p2, da_, s2)); s3 = fmaf(q3, db_, fmaf(p3, da_, s3)); }
        else { s0 += p0 * da_ + q0 * db_; s1 += p1 * da_ + q1 * db_; s2 += p2 * da_ + q2 * db_; s3 += p3 * da_ + q3 * db_; }
      }
    }
    anybad = anybad || bad;
  }
  const long long t1 = clock64();
  out[threadIdx.x] = s0 + s1 + s2 + s3 + mapp + appn + app1 + app2 + (anybad ? 1.f : 0.f);
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE>
static double run(float* out, long long* cyc, int sweeps, int nc, float mu = 0.5f) {
  long long h = 0;
  for (int rep = 0; rep < 2; rep++) { k<MODE><<<1, 32>>>(out, cyc, sweeps, 0.5f, 1e3f, 3e-4f, nc, mu); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); }
  return (double)h / sweeps;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 256); cudaMalloc(&cyc, 64);
  const int sweeps = 2000;
  printf("unit block (10 row steps), cycles per sweep: registers, no residual %.0f | smem columns %.0f | + residual bookkeeping %.0f\n",
         run<0>(out, cyc, sweeps, 0), run<1>(out, cyc, sweeps, 0), run<3>(out, cyc, sweeps, 0));
  const double b = run<3>(out, cyc, sweeps, 0);
  const int nc = 12;
  printf("per contact at nc = 12: normal + friction %.0f | normal only %.0f | friction only %.0f\n", (run<7>(out, cyc, sweeps, nc) - b) / nc,
         (run<7 + 16>(out, cyc, sweeps, nc) - b) / nc, (run<7 + 8>(out, cyc, sweeps, nc) - b) / nc);
  printf("friction only: FMA-chain apply %.0f | approx rsqrt %.0f | both %.0f\n", (run<7 + 8 + 32>(out, cyc, sweeps, nc) - b) / nc,
         (run<7 + 8 + 64>(out, cyc, sweeps, nc) - b) / nc, (run<7 + 8 + 32 + 64>(out, cyc, sweeps, nc) - b) / nc);
  printf("normal + friction, both friction changes: %.0f per contact\n", (run<7 + 32 + 64>(out, cyc, sweeps, nc) - b) / nc);
  printf("friction only, stick-first exchange: cone never bites (mu 1e6) %.0f | cone always bites (mu 1e-6) %.0f   [reference form: %.0f | %.0f]\n",
         (run<7 + 8 + 128>(out, cyc, sweeps, nc, 1e6f) - b) / nc, (run<7 + 8 + 128>(out, cyc, sweeps, nc, 1e-6f) - b) / nc,
         (run<7 + 8 + 32 + 64>(out, cyc, sweeps, nc, 1e6f) - b) / nc, (run<7 + 8 + 32 + 64>(out, cyc, sweeps, nc, 1e-6f) - b) / nc);

  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
