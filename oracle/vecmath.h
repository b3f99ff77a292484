// TEST INFRASTRUCTURE ONLY — part of the CPU oracle (see oracle/pt_oracle.cpp header).
// Tiny fp32 vector library with HLSL/Slang-like semantics for the scalar restatement.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

struct float2
{
  float x, y;
};
struct float3
{
  float x, y, z;
};
struct float4
{
  float x, y, z, w;
};

static inline float2 f2(float x, float y) { return {x, y}; }
static inline float3 f3(float x, float y, float z) { return {x, y, z}; }
static inline float3 f3(float s) { return {s, s, s}; }
static inline float4 f4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline float4 f4(float3 v, float w) { return {v.x, v.y, v.z, w}; }
static inline float3 xyz(float4 v) { return {v.x, v.y, v.z}; }

static inline float3 operator+(float3 a, float3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline float3 operator-(float3 a, float3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline float3 operator*(float3 a, float3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline float3 operator/(float3 a, float3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
static inline float3 operator*(float3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline float3 operator*(float s, float3 a) { return {a.x * s, a.y * s, a.z * s}; }
static inline float3 operator/(float3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
static inline float3 operator-(float3 a) { return {-a.x, -a.y, -a.z}; }
static inline float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
static inline float3& operator-=(float3& a, float3 b) { a = a - b; return a; }
static inline float3& operator*=(float3& a, float3 b) { a = a * b; return a; }
static inline float3& operator*=(float3& a, float s) { a = a * s; return a; }
static inline float3& operator/=(float3& a, float s) { a = a / s; return a; }

static inline float2 operator+(float2 a, float2 b) { return {a.x + b.x, a.y + b.y}; }
static inline float2 operator-(float2 a, float2 b) { return {a.x - b.x, a.y - b.y}; }
static inline float2 operator*(float2 a, float s) { return {a.x * s, a.y * s}; }
static inline float2 operator*(float s, float2 a) { return {a.x * s, a.y * s}; }

static inline float4 operator+(float4 a, float4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
static inline float4 operator*(float4 a, float4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
static inline float4 operator*(float4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
static inline float4 operator/(float4 a, float s) { return {a.x / s, a.y / s, a.z / s, a.w / s}; }
static inline float4& operator+=(float4& a, float4 b) { a = a + b; return a; }
static inline float4& operator*=(float4& a, float4 b) { a = a * b; return a; }

static inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
static inline float3 cross(float3 a, float3 b)
{
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline float length(float3 a) { return sqrtf(dot(a, a)); }
static inline float3 normalize(float3 a)
{
  float inv = 1.0f / sqrtf(dot(a, a));
  return a * inv;
}
static inline float2 normalize(float2 a)
{
  float inv = 1.0f / sqrtf(dot(a, a));
  return a * inv;
}
static inline float3 vmax(float3 a, float3 b) { return {fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)}; }
static inline float3 vmin(float3 a, float3 b) { return {fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)}; }
static inline float  maxc(float3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
static inline float  clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline float  saturate(float x) { return clampf(x, 0.0f, 1.0f); }
static inline float  lerpf(float a, float b, float t) { return a + (b - a) * t; }  // HLSL lerp
static inline float3 lerp3(float3 a, float3 b, float t) { return a + (b - a) * t; }
static inline float3 reflect(float3 i, float3 n) { return i - n * (2.0f * dot(n, i)); }
static inline float  square(float x) { return x * x; }
static inline float  signf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
static inline float  smoothstep(float e0, float e1, float x)
{
  float t = saturate((x - e0) / (e1 - e0));
  return t * t * (3.0f - 2.0f * t);
}
static inline float3 expv(float3 a) { return {expf(a.x), expf(a.y), expf(a.z)}; }
static inline float3 logv(float3 a) { return {logf(a.x), logf(a.y), logf(a.z)}; }
static inline float3 sqrtv(float3 a) { return {sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)}; }

static inline uint32_t asuint(float f)
{
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static inline int32_t asint(float f)
{
  int32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static inline float asfloat(uint32_t u)
{
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline float asfloat(int32_t u)
{
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// glm column-major 4x4: m[c*4+r].  Slang `mul(v, M)` == M_glm * v ; `mul(M, v)` == M_glm^T * v
// (reference convention note: SURVEY.md §8; get_hit.h.slang:76,80).
struct mat4
{
  float m[16];
};
static inline float4 mul_vM(float4 v, const mat4& M)  // M_glm * v
{
  float4 r;
  r.x = ((M.m[0] * v.x + M.m[4] * v.y) + M.m[8] * v.z) + M.m[12] * v.w;
  r.y = ((M.m[1] * v.x + M.m[5] * v.y) + M.m[9] * v.z) + M.m[13] * v.w;
  r.z = ((M.m[2] * v.x + M.m[6] * v.y) + M.m[10] * v.z) + M.m[14] * v.w;
  r.w = ((M.m[3] * v.x + M.m[7] * v.y) + M.m[11] * v.z) + M.m[15] * v.w;
  return r;
}
static inline float4 mul_Mv(const mat4& M, float4 v)  // M_glm^T * v
{
  float4 r;
  r.x = ((M.m[0] * v.x + M.m[1] * v.y) + M.m[2] * v.z) + M.m[3] * v.w;
  r.y = ((M.m[4] * v.x + M.m[5] * v.y) + M.m[6] * v.z) + M.m[7] * v.w;
  r.z = ((M.m[8] * v.x + M.m[9] * v.y) + M.m[10] * v.z) + M.m[11] * v.w;
  r.w = ((M.m[12] * v.x + M.m[13] * v.y) + M.m[14] * v.z) + M.m[15] * v.w;
  return r;
}
// point / vector transforms by objectToWorld (mul(float4(p,1), M).xyz and mul(float4(v,0), M).xyz)
static inline float3 xfPoint(const mat4& M, float3 p)
{
  return {((M.m[0] * p.x + M.m[4] * p.y) + M.m[8] * p.z) + M.m[12], ((M.m[1] * p.x + M.m[5] * p.y) + M.m[9] * p.z) + M.m[13],
          ((M.m[2] * p.x + M.m[6] * p.y) + M.m[10] * p.z) + M.m[14]};
}
static inline float3 xfVector(const mat4& M, float3 v)
{
  return {(M.m[0] * v.x + M.m[4] * v.y) + M.m[8] * v.z, (M.m[1] * v.x + M.m[5] * v.y) + M.m[9] * v.z,
          (M.m[2] * v.x + M.m[6] * v.y) + M.m[10] * v.z};
}
// normal transform: mul(worldToObject, float4(n,0)).xyz == W2O_glm^T * n
static inline float3 xfNormal(const mat4& W2O, float3 n)
{
  return {(W2O.m[0] * n.x + W2O.m[1] * n.y) + W2O.m[2] * n.z, (W2O.m[4] * n.x + W2O.m[5] * n.y) + W2O.m[6] * n.z,
          (W2O.m[8] * n.x + W2O.m[9] * n.y) + W2O.m[10] * n.z};
}

// IEEE binary16 round-trip (round-to-nearest-even), for the reference's float16_t VolumeMedium
// fields (pathtrace_functions.h.slang:118-123).
static inline uint16_t f32_to_f16(float f)
{
  uint32_t x = asuint(f);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t exp = (x >> 23) & 0xFF;
  uint32_t man = x & 0x7FFFFFu;
  if(exp == 0xFF)
    return (uint16_t)(sign | 0x7C00u | (man ? 0x200u : 0));
  int e = (int)exp - 127 + 15;
  if(e >= 31)
    return (uint16_t)(sign | 0x7C00u);
  if(e <= 0)
  {
    if(e < -10)
      return (uint16_t)sign;
    man |= 0x800000u;
    int      shift = 14 - e;
    uint32_t half = man >> shift;
    uint32_t rem = man & ((1u << shift) - 1);
    uint32_t mid = 1u << (shift - 1);
    if(rem > mid || (rem == mid && (half & 1)))
      half++;
    return (uint16_t)(sign | half);
  }
  uint32_t half = (uint32_t)(e << 10) | (man >> 13);
  uint32_t rem = man & 0x1FFFu;
  if(rem > 0x1000u || (rem == 0x1000u && (half & 1)))
    half++;
  return (uint16_t)(sign | half);
}
static inline float f16_to_f32(uint16_t h)
{
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1F;
  uint32_t man = h & 0x3FFu;
  if(exp == 0)
  {
    if(man == 0)
      return asfloat(sign);
    float f = (float)man * (1.0f / 16777216.0f);  // 2^-24
    return (sign ? -f : f);
  }
  if(exp == 31)
    return asfloat(sign | 0x7F800000u | (man << 13));
  return asfloat(sign | ((exp - 15 + 127) << 23) | (man << 13));
}
static inline float  roundHalf(float f) { return f16_to_f32(f32_to_f16(f)); }
static inline float3 roundHalf(float3 v) { return {roundHalf(v.x), roundHalf(v.y), roundHalf(v.z)}; }

}  // namespace orc
