// ocl_shim.hpp -- just enough of OpenCL C 1.2 in C++17 to compile the reference's UNMODIFIED kernel sources
// (physics/ocl/kernels/{define,grid,utils,sph,boids,fluids,clouds}.cl under /root/reference) for the CPU.
// TEST INFRASTRUCTURE ONLY (see oracle/ref/README.md). The kernels are included textually by ref_kernels.cpp after a
// purely syntactic rewrite of OpenCL vector literals "(float4)(a, b, c, d)" into constructor calls "float4(a, b, c, d)"
// (in C++ the former parses as a cast of a comma expression). No reference source is stored in this repository.
//
// Built-in semantics follow the OpenCL 1.2 specification; where it leaves precision to the driver the same canonical
// choices as oracle/rtp_oracle.c are made (dot = fma chain, fast_length = sqrtf(dot), pow(x,2|3) by products, exp =
// the canonical exp), so that discrete outputs and element-wise stages compare bit for bit and neighbour sums to
// rounding-sequence accuracy.
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>

typedef unsigned int uint;
typedef unsigned long ulong;
typedef unsigned char uchar;

#define __kernel
#define __global
#define __local
#define __constant const
#define CLK_LOCAL_MEM_FENCE 0

// everything lives in namespace ocl so that the built-ins (exp, pow, floor, fabs, min, max ...) hide, rather than
// collide with, the C library's global declarations when the kernels are included inside the namespace
namespace ocl
{
extern thread_local size_t g_globalId;
extern size_t g_globalSize;
inline size_t get_global_id(int) { return g_globalId; }
inline size_t get_global_size(int) { return g_globalSize; }

// ---------------------------------------------------------------- vector types
struct float3;
struct float3_swizzle // what "v.xyz" names: the same three floats, convertible to / assignable from a float3
{
  float x, y, z;
  inline operator float3() const;
  inline float3_swizzle& operator=(const float3& v);
};
struct float3
{
  union
  {
    struct
    {
      float x, y, z;
    };
    float3_swizzle xyz;
  };
  float3() = default;
  float3(float s) : x(s), y(s), z(s) {}
  float3(float a, float b, float c) : x(a), y(b), z(c) {}
};
inline float3_swizzle::operator float3() const { return float3(x, y, z); }
inline float3_swizzle& float3_swizzle::operator=(const float3& v)
{
  x = v.x, y = v.y, z = v.z;
  return *this;
}
struct float2
{
  float x, y;
  float2() = default;
  float2(float s) : x(s), y(s) {}
  float2(float a, float b) : x(a), y(b) {}
};
struct float4
{
  union
  {
    struct
    {
      float x, y, z, w;
    };
    float3_swizzle xyz;
  };
  float4() : x(0), y(0), z(0), w(0) {}
  float4(float s) : x(s), y(s), z(s), w(s) {}
  float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
};
struct float8
{
  float s[8];
  float8() = default;
  float8(float v)
  {
    for (float& e : s)
      e = v;
  }
  float8& operator=(float v)
  {
    for (float& e : s)
      e = v;
    return *this;
  }
};
struct int3
{
  int x, y, z;
  int3() = default;
  int3(int s) : x(s), y(s), z(s) {}
  int3(int a, int b, int c) : x(a), y(b), z(c) {}
};
struct uint3
{
  uint x, y, z;
  uint3() = default;
  uint3(uint s) : x(s), y(s), z(s) {}
  uint3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
};
struct uint2
{
  uint x, y;
  uint2() = default;
  uint2(uint s) : x(s), y(s) {}
  uint2(uint a, uint b) : x(a), y(b) {}
};

// ---------------------------------------------------------------- float4
inline float4 operator+(float4 a, float4 b) { return float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline float4 operator-(float4 a, float4 b) { return float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline float4 operator*(float4 a, float4 b) { return float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline float4 operator-(float4 a) { return float4(-a.x, -a.y, -a.z, -a.w); }
inline float4 operator*(float4 a, float s) { return float4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline float4 operator*(float s, float4 a) { return float4(s * a.x, s * a.y, s * a.z, s * a.w); }
inline float4 operator*(float4 a, int s) { return a * (float)s; }
inline float4 operator*(int s, float4 a) { return (float)s * a; }
inline float4 operator*(float4 a, uint s) { return a * (float)s; }
inline float4 operator/(float4 a, float s) { return float4(a.x / s, a.y / s, a.z / s, a.w / s); }
inline float4 operator/(float4 a, int s) { return a / (float)s; }
inline float4& operator+=(float4& a, float4 b) { return a = a + b; }
inline float4& operator-=(float4& a, float4 b) { return a = a - b; }
inline float4& operator*=(float4& a, float s) { return a = a * s; }
inline float4& operator/=(float4& a, float s) { return a = a / s; }
inline float4& operator/=(float4& a, int s) { return a = a / (float)s; }

// ---------------------------------------------------------------- float3
inline float3 operator+(float3 a, float3 b) { return float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator-(float3 a, float3 b) { return float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator*(float3 a, float3 b) { return float3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float3 operator-(float3 a, float s) { return float3(a.x - s, a.y - s, a.z - s); }
inline float3 operator*(float3 a, float s) { return float3(a.x * s, a.y * s, a.z * s); }
inline float3 operator*(float s, float3 a) { return float3(s * a.x, s * a.y, s * a.z); }
inline float3 operator/(float3 a, float s) { return float3(a.x / s, a.y / s, a.z / s); }

// ---------------------------------------------------------------- int3 / uint3
inline int3 operator+(int3 a, int3 b) { return int3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline int3 operator%(int3 a, int3 b) { return int3(a.x % b.x, a.y % b.y, a.z % b.z); }
// relational operators on vectors: -1 (all bits set) for true, 0 for false
inline int3 operator<(int3 a, int3 b) { return int3(-(a.x < b.x), -(a.y < b.y), -(a.z < b.z)); }
inline int3 operator>=(int3 a, int3 b) { return int3(-(a.x >= b.x), -(a.y >= b.y), -(a.z >= b.z)); }
inline int any(int3 a) { return (a.x < 0) || (a.y < 0) || (a.z < 0); } // most significant bit set
inline int all(int3 a) { return (a.x < 0) && (a.y < 0) && (a.z < 0); }
inline int3 isequal(float3 a, float3 b) { return int3(-(a.x == b.x), -(a.y == b.y), -(a.z == b.z)); }
inline int isequal(float a, float b) { return a == b; }
inline int3 convert_int3(uint3 a) { return int3((int)a.x, (int)a.y, (int)a.z); }
inline uint3 convert_uint3(float3 a) { return uint3((uint)a.x, (uint)a.y, (uint)a.z); } // truncation

// ---------------------------------------------------------------- built-ins
inline float clamp(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }
inline float3 clamp(float3 x, float3 lo, float3 hi) { return float3(clamp(x.x, lo.x, hi.x), clamp(x.y, lo.y, hi.y), clamp(x.z, lo.z, hi.z)); }
inline float4 clamp(float4 x, float4 lo, float4 hi)
{
  return float4(clamp(x.x, lo.x, hi.x), clamp(x.y, lo.y, hi.y), clamp(x.z, lo.z, hi.z), clamp(x.w, lo.w, hi.w));
}
inline float4 clamp(float4 x, float lo, float hi) { return clamp(x, float4(lo), float4(hi)); }
inline float3 floor(float3 a) { return float3(std::floor(a.x), std::floor(a.y), std::floor(a.z)); }
inline float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
inline float min(float a, float b) { return std::fmin(a, b); }
inline float max(float a, float b) { return std::fmax(a, b); }
inline uint min(uint a, uint b) { return a < b ? a : b; }
inline uint max(uint a, uint b) { return a > b ? a : b; }
inline float fabs(float a) { return std::fabs(a); }

// canonical: fma chain, w is 0 on this path
inline float dot(float4 a, float4 b) { return std::fma(a.z, b.z, std::fma(a.y, b.y, a.x * b.x)); }
inline float dot(float3 a, float3 b) { return std::fma(a.z, b.z, std::fma(a.y, b.y, a.x * b.x)); }
inline float fast_length(float4 a) { return std::sqrt(dot(a, a)); }
inline float length(float4 a) { return std::sqrt(dot(a, a)); }
inline float length(float3 a) { return std::sqrt(dot(a, a)); }
// OpenCL 1.2 s6.12.5: a vector whose elements are all zero is returned unchanged (tested as dot == 0, like the oracle)
inline float4 fast_normalize(float4 a)
{
  const float d = dot(a, a);
  return d == 0.0f ? a : a * (1.0f / std::sqrt(d));
}
inline float4 normalize(float4 a)
{
  const float l = std::sqrt(dot(a, a));
  return l == 0.0f ? float4(0.0f) : a / l;
}
inline float4 cross(float4 a, float4 b) { return float4(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x, 0.0f); }
inline float pow(float x, uint n) { return n == 2 ? x * x : (n == 3 ? x * x * x : std::pow(x, (float)n)); }
inline float pow(float x, int n) { return pow(x, (uint)n); }

// canonical exp (identical to oracle/rtp_oracle.c canon_expf / canon_exp)
inline float exp(float x)
{
  x = std::fmin(std::fmax(x, -87.0f), 88.0f);
  const float n = std::rint(x * 1.44269504f);
  float r = std::fma(n, -0.693145751953125f, x);
  r = std::fma(n, -1.428606765330187e-06f, r);
  float p = 1.98412698e-4f;
  p = std::fma(p, r, 1.38888889e-3f);
  p = std::fma(p, r, 8.33333333e-3f);
  p = std::fma(p, r, 4.16666667e-2f);
  p = std::fma(p, r, 1.66666667e-1f);
  p = std::fma(p, r, 0.5f);
  p = std::fma(p, r, 1.0f);
  p = std::fma(p, r, 1.0f);
  union
  {
    uint32_t u;
    float f;
  } s;
  s.u = (uint32_t)((int)n + 127) << 23;
  return p * s.f;
}
inline double exp(double x)
{
  x = std::fmin(std::fmax(x, -700.0), 700.0);
  const double n = std::rint(x * 1.4426950408889634);
  double r = std::fma(n, -6.93147180369123816490e-01, x);
  r = std::fma(n, -1.90821492927058770002e-10, r);
  double p = 1.0 / 6227020800.0;
  p = std::fma(p, r, 1.0 / 479001600.0);
  p = std::fma(p, r, 1.0 / 39916800.0);
  p = std::fma(p, r, 1.0 / 3628800.0);
  p = std::fma(p, r, 1.0 / 362880.0);
  p = std::fma(p, r, 1.0 / 40320.0);
  p = std::fma(p, r, 1.0 / 5040.0);
  p = std::fma(p, r, 1.0 / 720.0);
  p = std::fma(p, r, 1.0 / 120.0);
  p = std::fma(p, r, 1.0 / 24.0);
  p = std::fma(p, r, 1.0 / 6.0);
  p = std::fma(p, r, 0.5);
  p = std::fma(p, r, 1.0);
  p = std::fma(p, r, 1.0);
  union
  {
    uint64_t u;
    double f;
  } s;
  s.u = (uint64_t)((int64_t)n + 1023) << 52;
  return p * s.f;
}

} // namespace ocl
