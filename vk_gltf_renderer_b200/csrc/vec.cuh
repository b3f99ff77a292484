// vec.cuh — fp32 vector helpers, RNG and bit utilities for the sm_100a path-tracer kernels.
// RNG: xxhash32 seed + PCG stream, the reference's nvshaders/random.h.slang contract
// (call sites shaders/gltf_pathtrace.slang:560 and ~20 rand() sites; SURVEY.md §8 "RNG contract").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PT_HD __host__ __device__ __forceinline__
#define PT_D __device__ __forceinline__

namespace pt {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kTwoPi = 6.28318530717958647692f;
constexpr float kInvPi = 0.31830988618379067154f;
constexpr float kInfinite = 1e32f;  // INFINITE: miss sentinel (gltf_pathtrace.slang:112)
constexpr float kDirac = -1.0f;     // DIRAC pdf sentinel (gltf_pathtrace.slang:344)

PT_HD float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
PT_HD float3 f3(float s) { return make_float3(s, s, s); }
PT_HD float2 f2(float x, float y) { return make_float2(x, y); }
PT_HD float4 f4(float x, float y, float z, float w) { return make_float4(x, y, z, w); }
PT_HD float4 f4(float3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
PT_HD float3 xyz(float4 v) { return make_float3(v.x, v.y, v.z); }

PT_HD float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
PT_HD float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
PT_HD float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
PT_HD float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
PT_HD float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
PT_HD float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
PT_HD float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
PT_HD float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
PT_HD float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
PT_HD float3& operator-=(float3& a, float3 b) { a = a - b; return a; }
PT_HD float3& operator*=(float3& a, float3 b) { a = a * b; return a; }
PT_HD float3& operator*=(float3& a, float s) { a = a * s; return a; }
PT_HD float3& operator/=(float3& a, float s) { a = a / s; return a; }
PT_HD float2 operator+(float2 a, float2 b) { return f2(a.x + b.x, a.y + b.y); }
PT_HD float2 operator-(float2 a, float2 b) { return f2(a.x - b.x, a.y - b.y); }
PT_HD float2 operator*(float2 a, float s) { return f2(a.x * s, a.y * s); }
PT_HD float4 operator+(float4 a, float4 b) { return f4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
PT_HD float4 operator*(float4 a, float4 b) { return f4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
PT_HD float4 operator*(float4 a, float s) { return f4(a.x * s, a.y * s, a.z * s, a.w * s); }
PT_HD float4 operator/(float4 a, float s) { return f4(a.x / s, a.y / s, a.z / s, a.w / s); }
PT_HD float4& operator+=(float4& a, float4 b) { a = a + b; return a; }
PT_HD float4& operator*=(float4& a, float4 b) { a = a * b; return a; }

PT_HD float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PT_HD float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
PT_HD float3 cross(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
PT_HD float length(float3 a) { return sqrtf(dot(a, a)); }
PT_HD float3 normalize(float3 a)
{
  float inv = 1.0f / sqrtf(dot(a, a));
  return a * inv;
}
PT_HD float2 normalize(float2 a)
{
  float inv = 1.0f / sqrtf(dot(a, a));
  return a * inv;
}
PT_HD float3 vmax(float3 a, float3 b) { return f3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
PT_HD float3 vmin(float3 a, float3 b) { return f3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
PT_HD float  maxc(float3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
PT_HD float  clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
PT_HD float  saturatef(float x) { return clampf(x, 0.0f, 1.0f); }
PT_HD float  lerpf(float a, float b, float t) { return a + (b - a) * t; }
PT_HD float3 lerp3(float3 a, float3 b, float t) { return a + (b - a) * t; }
PT_HD float3 reflect(float3 i, float3 n) { return i - n * (2.0f * dot(n, i)); }
PT_HD float  square(float x) { return x * x; }
PT_HD float  signf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
PT_HD float  smoothstepf(float e0, float e1, float x)
{
  float t = saturatef((x - e0) / (e1 - e0));
  return t * t * (3.0f - 2.0f * t);
}
PT_HD float3 expv(float3 a) { return f3(expf(a.x), expf(a.y), expf(a.z)); }
PT_HD float3 logv(float3 a) { return f3(logf(a.x), logf(a.y), logf(a.z)); }
PT_HD float3 sqrtv(float3 a) { return f3(sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)); }

// explicit-FMA cross/dot for the ray/triangle test: the CPU oracle uses the identical fmaf chain,
// so (t,u,v) agree bit-for-bit between the two (oracle/pt_oracle.cpp crossFma/dotFma)
PT_HD float3 crossFma(float3 a, float3 b)
{
  return f3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
PT_HD float dotFma(float3 a, float3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }

// ---- glm column-major 4x4 helpers: m[c*4+r]; `mul(v,M)` == M_glm*v, `mul(M,v)` == M_glm^T*v ----
struct Mat4
{
  float m[16];
};
PT_HD float4 mul_vM(float4 v, const Mat4& M)
{
  return f4(((M.m[0] * v.x + M.m[4] * v.y) + M.m[8] * v.z) + M.m[12] * v.w, ((M.m[1] * v.x + M.m[5] * v.y) + M.m[9] * v.z) + M.m[13] * v.w,
            ((M.m[2] * v.x + M.m[6] * v.y) + M.m[10] * v.z) + M.m[14] * v.w, ((M.m[3] * v.x + M.m[7] * v.y) + M.m[11] * v.z) + M.m[15] * v.w);
}
PT_HD float4 mul_Mv(const Mat4& M, float4 v)
{
  return f4(((M.m[0] * v.x + M.m[1] * v.y) + M.m[2] * v.z) + M.m[3] * v.w, ((M.m[4] * v.x + M.m[5] * v.y) + M.m[6] * v.z) + M.m[7] * v.w,
            ((M.m[8] * v.x + M.m[9] * v.y) + M.m[10] * v.z) + M.m[11] * v.w, ((M.m[12] * v.x + M.m[13] * v.y) + M.m[14] * v.z) + M.m[15] * v.w);
}
PT_HD float3 xfPoint(const float* m, float3 p)
{
  return f3(((m[0] * p.x + m[4] * p.y) + m[8] * p.z) + m[12], ((m[1] * p.x + m[5] * p.y) + m[9] * p.z) + m[13],
            ((m[2] * p.x + m[6] * p.y) + m[10] * p.z) + m[14]);
}
PT_HD float3 xfVector(const float* m, float3 v)
{
  return f3((m[0] * v.x + m[4] * v.y) + m[8] * v.z, (m[1] * v.x + m[5] * v.y) + m[9] * v.z, (m[2] * v.x + m[6] * v.y) + m[10] * v.z);
}
// normals: mul(worldToObject, float4(n,0)).xyz == W2O^T * n
PT_HD float3 xfNormal(const float* w2o, float3 n)
{
  return f3((w2o[0] * n.x + w2o[1] * n.y) + w2o[2] * n.z, (w2o[4] * n.x + w2o[5] * n.y) + w2o[6] * n.z,
            (w2o[8] * n.x + w2o[9] * n.y) + w2o[10] * n.z);
}

// ---- RNG -------------------------------------------------------------------------------------
PT_HD uint32_t xxhash32(uint32_t px, uint32_t py, uint32_t pz)
{
  const uint32_t P1 = 2246822519u, P2 = 3266489917u, P3 = 668265263u, P4 = 374761393u;
  uint32_t       h32 = pz + P4 + px * P2;
  h32 = P3 * ((h32 << 17) | (h32 >> 15));
  h32 += py * P2;
  h32 = P3 * ((h32 << 17) | (h32 >> 15));
  h32 = P1 * (h32 ^ (h32 >> 15));
  h32 = P2 * (h32 ^ (h32 >> 13));
  return h32 ^ (h32 >> 16);
}
PT_HD uint32_t pcg(uint32_t& state)
{
  uint32_t prev = state * 747796405u + 2891336453u;
  uint32_t word = ((prev >> ((prev >> 28u) + 4u)) ^ prev) * 277803737u;
  state = prev;
  return (word >> 22u) ^ word;
}
PT_D float rnd(uint32_t& seed)
{
  uint32_t r = pcg(seed);
  return __uint_as_float(0x3f800000u | (r >> 9)) - 1.0f;
}

// Waechter-Binder self-intersection offset (pathtrace_functions.h.slang:151-167)
PT_D float3 safeOffsetRay(float3 p, float3 dir)
{
  const float scaleValue = 256.0f;
  const int   sx = (int)(scaleValue * dir.x), sy = (int)(scaleValue * dir.y), sz = (int)(scaleValue * dir.z);
  const float3 op = f3(__int_as_float(__float_as_int(p.x) + ((p.x < 0) ? -sx : sx)), __int_as_float(__float_as_int(p.y) + ((p.y < 0) ? -sy : sy)),
                       __int_as_float(__float_as_int(p.z) + ((p.z < 0) ? -sz : sz)));
  const float origin = 1.0f / 32.0f, floatScale = 1.0f / 65536.0f;
  return f3(fabsf(p.x) < origin ? p.x + floatScale * dir.x : op.x, fabsf(p.y) < origin ? p.y + floatScale * dir.y : op.y,
            fabsf(p.z) < origin ? p.z + floatScale * dir.z : op.z);
}

}  // namespace pt
