// xarm_math.cuh - small fixed-size float math for the per-env device code (one thread simulates one env,
// a warp steps 32 envs in lock-step; everything here lives in registers after full unrolling).
#pragma once
#include <math.h>
#include <stdint.h>
#ifdef XARM_HOST_SIM
// Host build of the device headers, used ONLY by tests/hostsim (a development aid that lets the kernel logic be
// checked against the oracle on a box without a GPU).  The product library never defines XARM_HOST_SIM.
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __constant__ static const
#define __restrict__
#ifdef XARM_HOST_SIM_DOUBLE
// diagnostic build: the whole kernel logic in double precision (separates float rounding from logic differences)
#define float double
#define sqrtf sqrt
#define sinf sin
#define cosf cos
#define sincosf sincos
#define atan2f atan2
#define asinf asin
#define acosf acos
#define fabsf fabs
#define fminf fmin
#define fmaxf fmax
#define tanhf tanh
#endif
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
#else
#include <cuda_runtime.h>
#endif

#define XD __device__ __forceinline__

struct V3 {
  float x, y, z;
};
XD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
XD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
XD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
XD V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
XD V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
XD V3& operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
XD V3& operator-=(V3& a, V3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
XD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
XD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
XD float norm(V3 a) { return sqrtf(dot(a, a)); }
XD float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
XD void setcomp(V3& a, int i, float v) { if (i == 0) a.x = v; else if (i == 1) a.y = v; else a.z = v; }

// 3x3, row-major
struct M3 {
  float m[9];
};
XD V3 operator*(const M3& A, V3 x) {
  return v3(A.m[0] * x.x + A.m[1] * x.y + A.m[2] * x.z, A.m[3] * x.x + A.m[4] * x.y + A.m[5] * x.z,
            A.m[6] * x.x + A.m[7] * x.y + A.m[8] * x.z);
}
XD V3 tmul(const M3& A, V3 x) {  // A^T x
  return v3(A.m[0] * x.x + A.m[3] * x.y + A.m[6] * x.z, A.m[1] * x.x + A.m[4] * x.y + A.m[7] * x.z,
            A.m[2] * x.x + A.m[5] * x.y + A.m[8] * x.z);
}
XD M3 operator*(const M3& A, const M3& B) {
  M3 C;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return C;
}
XD V3 col(const M3& A, int j) { return v3(A.m[j], A.m[3 + j], A.m[6 + j]); }
XD M3 m3_identity() { M3 R; R.m[0] = 1; R.m[1] = 0; R.m[2] = 0; R.m[3] = 0; R.m[4] = 1; R.m[5] = 0; R.m[6] = 0; R.m[7] = 0; R.m[8] = 1; return R; }
XD M3 m3_axis_angle(V3 a, float th) {
  float s, c;
  sincosf(th, &s, &c);
  float t = 1.f - c;
  M3 R;
  R.m[0] = t * a.x * a.x + c;       R.m[1] = t * a.x * a.y - s * a.z; R.m[2] = t * a.x * a.z + s * a.y;
  R.m[3] = t * a.x * a.y + s * a.z; R.m[4] = t * a.y * a.y + c;       R.m[5] = t * a.y * a.z - s * a.x;
  R.m[6] = t * a.x * a.z - s * a.y; R.m[7] = t * a.y * a.z + s * a.x; R.m[8] = t * a.z * a.z + c;
  return R;
}

// symmetric 3x3: xx xy xz yy yz zz
struct S3 {
  float xx, xy, xz, yy, yz, zz;
};
XD V3 operator*(const S3& A, V3 v) {
  return v3(A.xx * v.x + A.xy * v.y + A.xz * v.z, A.xy * v.x + A.yy * v.y + A.yz * v.z, A.xz * v.x + A.yz * v.y + A.zz * v.z);
}
XD S3 operator+(const S3& a, const S3& b) { S3 r = {a.xx + b.xx, a.xy + b.xy, a.xz + b.xz, a.yy + b.yy, a.yz + b.yz, a.zz + b.zz}; return r; }
// R S R^T for rotation R
XD S3 rotate_sym(const M3& R, const S3& S) {
  M3 T;  // T = R S
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float a = R.m[3 * i], b = R.m[3 * i + 1], c = R.m[3 * i + 2];
    T.m[3 * i] = a * S.xx + b * S.xy + c * S.xz;
    T.m[3 * i + 1] = a * S.xy + b * S.yy + c * S.yz;
    T.m[3 * i + 2] = a * S.xz + b * S.yz + c * S.zz;
  }
  S3 o;
  o.xx = T.m[0] * R.m[0] + T.m[1] * R.m[1] + T.m[2] * R.m[2];
  o.xy = T.m[0] * R.m[3] + T.m[1] * R.m[4] + T.m[2] * R.m[5];
  o.xz = T.m[0] * R.m[6] + T.m[1] * R.m[7] + T.m[2] * R.m[8];
  o.yy = T.m[3] * R.m[3] + T.m[4] * R.m[4] + T.m[5] * R.m[5];
  o.yz = T.m[3] * R.m[6] + T.m[4] * R.m[7] + T.m[5] * R.m[8];
  o.zz = T.m[6] * R.m[6] + T.m[7] * R.m[7] + T.m[8] * R.m[8];
  return o;
}

// quaternion xyzw
struct Q4 {
  float x, y, z, w;
};
XD M3 quat_to_m3(Q4 q) {
  M3 R;
  R.m[0] = 1 - 2 * (q.y * q.y + q.z * q.z); R.m[1] = 2 * (q.x * q.y - q.z * q.w);     R.m[2] = 2 * (q.x * q.z + q.y * q.w);
  R.m[3] = 2 * (q.x * q.y + q.z * q.w);     R.m[4] = 1 - 2 * (q.x * q.x + q.z * q.z); R.m[5] = 2 * (q.y * q.z - q.x * q.w);
  R.m[6] = 2 * (q.x * q.z - q.y * q.w);     R.m[7] = 2 * (q.y * q.z + q.x * q.w);     R.m[8] = 1 - 2 * (q.x * q.x + q.y * q.y);
  return R;
}
XD Q4 m3_to_quat(const M3& R) {
  Q4 q;
  float tr = R.m[0] + R.m[4] + R.m[8];
  if (tr > 0.f) {
    float s = sqrtf(tr + 1.f);
    q.w = 0.5f * s; s = 0.5f / s;
    q.x = (R.m[7] - R.m[5]) * s; q.y = (R.m[2] - R.m[6]) * s; q.z = (R.m[3] - R.m[1]) * s;
  } else if (R.m[0] >= R.m[4] && R.m[0] >= R.m[8]) {
    float s = sqrtf(R.m[0] - R.m[4] - R.m[8] + 1.f);
    q.x = 0.5f * s; s = 0.5f / s;
    q.w = (R.m[7] - R.m[5]) * s; q.y = (R.m[3] + R.m[1]) * s; q.z = (R.m[6] + R.m[2]) * s;
  } else if (R.m[4] >= R.m[8]) {
    float s = sqrtf(R.m[4] - R.m[8] - R.m[0] + 1.f);
    q.y = 0.5f * s; s = 0.5f / s;
    q.w = (R.m[2] - R.m[6]) * s; q.z = (R.m[7] + R.m[5]) * s; q.x = (R.m[1] + R.m[3]) * s;
  } else {
    float s = sqrtf(R.m[8] - R.m[0] - R.m[4] + 1.f);
    q.z = 0.5f * s; s = 0.5f / s;
    q.w = (R.m[3] - R.m[1]) * s; q.x = (R.m[2] + R.m[6]) * s; q.y = (R.m[5] + R.m[7]) * s;
  }
  return q;
}
XD Q4 quat_mul(Q4 a, Q4 b) {
  Q4 o;
  o.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  o.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
  o.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
  o.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  return o;
}

// spatial vectors about the world origin: motion [w; v], force [n; f]
struct SV {
  V3 a, l;  // angular part, linear part
};
XD SV operator+(SV x, SV y) { SV r; r.a = x.a + y.a; r.l = x.l + y.l; return r; }
XD SV operator*(float s, SV x) { SV r; r.a = s * x.a; r.l = s * x.l; return r; }
XD SV& operator+=(SV& x, SV y) { x.a += y.a; x.l += y.l; return x; }
XD float dot(SV x, SV y) { return dot(x.a, y.a) + dot(x.l, y.l); }
XD SV sv_zero() { SV r; r.a = v3(0, 0, 0); r.l = v3(0, 0, 0); return r; }
XD SV motion_cross(SV v, SV m) { SV r; r.a = cross(v.a, m.a); r.l = cross(v.a, m.l) + cross(v.l, m.a); return r; }
XD SV force_cross(SV v, SV f) { SV r; r.a = cross(v.a, f.a) + cross(v.l, f.l); r.l = cross(v.a, f.l); return r; }

// rigid-body spatial inertia about the world origin: mass, first moment h = m c, rotational inertia about the origin
struct SI {
  float m;
  V3 h;
  S3 I;
};
XD SV operator*(const SI& I, SV v) {
  SV f;
  f.a = I.I * v.a + cross(I.h, v.l);
  f.l = I.m * v.l - cross(I.h, v.a);
  return f;
}
XD SI operator+(const SI& a, const SI& b) { SI r; r.m = a.m + b.m; r.h = a.h + b.h; r.I = a.I + b.I; return r; }
XD SI si_make(float m, V3 c, const S3& Ic) {
  SI r;
  r.m = m; r.h = m * c;
  float cc = dot(c, c);
  r.I.xx = Ic.xx + m * (cc - c.x * c.x); r.I.xy = Ic.xy - m * c.x * c.y; r.I.xz = Ic.xz - m * c.x * c.z;
  r.I.yy = Ic.yy + m * (cc - c.y * c.y); r.I.yz = Ic.yz - m * c.y * c.z; r.I.zz = Ic.zz + m * (cc - c.z * c.z);
  return r;
}
