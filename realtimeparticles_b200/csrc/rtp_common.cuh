// rtp_common.cuh -- shared device types and helpers of the sm_100a backend.
//
// Arithmetic policy (DESIGN.md "Numerics"):
//  * everything that feeds a DISCRETE decision or an element-wise stage is computed with explicitly rounded,
//    never-contracted IEEE fp32 operations (__fadd_rn / __fmul_rn / __fdiv_rn / __fsqrt_rn / __fmaf_rn) in the
//    reference's expression order, so cell ids, hit tests, predict/boundary/integrate stages are bit-exact with
//    the oracle;
//  * the per-pair terms inside the neighbour sums follow ONE canonical operation sequence (explicit FMAs, IEEE
//    sqrt and reciprocal, listed in DESIGN.md "Canonical arithmetic" and restated independently in
//    oracle/rtp_oracle.c), accumulated in the reference's order -- so every field of every model is bit-exact
//    with the oracle, not merely within tolerance. The compiler is never allowed to contract or reassociate:
//    all of it is written with the _rn intrinsics.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rtp_cuda.h"

namespace rtp
{
typedef uint32_t u32;

// Programmatic dependent launch (sm_90+), compile-time option RTP_USE_PDL: 0 = plain stream order (default: measured
// fastest on the 130k step), 1 = every kernel lets its successor start launching CTAs at once and
// waits for its predecessor's memory (cudaGridDependencySynchronize), 2 = wait only (the successor is released when the
// predecessor's CTAs have exited). All launches go through launchKernel(), also inside a captured CUDA graph.
#ifndef RTP_USE_PDL
#define RTP_USE_PDL 0
#endif
#if RTP_USE_PDL == 1
#define RTP_PDL_PROLOGUE()                     \
  do                                           \
  {                                            \
    cudaTriggerProgrammaticLaunchCompletion(); \
    cudaGridDependencySynchronize();           \
  } while (0)
#elif RTP_USE_PDL == 2
#define RTP_PDL_PROLOGUE() cudaGridDependencySynchronize()
#else
#define RTP_PDL_PROLOGUE() \
  do                       \
  {                        \
  } while (0)
#endif

template <typename... P, typename... A>
inline void launchKernel(void (*kernel)(P...), unsigned grid, unsigned block, cudaStream_t st, A&&... args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(block, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
#if RTP_USE_PDL
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#endif
  cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

#define RTP_FLOAT_EPS 0.00000001f // define.cl:6
#define RTP_ABS_GRAVITY_ACC_Y 9.81f // define.cl:8
#define RTP_FAR_DIST 1000000.0f // define.cl:10
#define RTP_MAX_STEERING 0.5f // boids.cl:10

// exactly rounded, never contracted
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ float frcp(float a) { return __frcp_rn(a); } // == 1.0f / a
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
// OpenCL clamp(x, lo, hi) = fmin(fmax(x, lo), hi)
__device__ __forceinline__ float fclamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// canonical dot product of the path: fma(z,z, fma(y,y, x*x)) (w is always 0) -- same as oracle dotc()
__device__ __forceinline__ float dot3c(float ax, float ay, float az, float bx, float by, float bz)
{
  return ffma(az, bz, ffma(ay, by, fmul(ax, bx)));
}

// The grid and the baked -D constants of a model (Boids.cpp:103-113, Fluids.cpp:104-119, Clouds.cpp:132-147).
struct GridParams
{
  float absW[3]; // ABS_WALL_X/Y/Z
  float cellSize; // GRID_CELL_SIZE_XYZ
  int res[3]; // GRID_RES_X/Y/Z
  u32 numCells; // GRID_NUM_CELLS
  u32 maxPartsInCell; // NUM_MAX_PARTS_IN_CELL
};

struct SphConsts
{
  float h; // EFFECT_RADIUS
  float h2; // h*h (rounded)
  float supportSq; // smallest float x with sqrtf(x) >= h : (sqrtf(sq) < h) <=> (sq < supportSq)
  float epsSq; // largest float x with sqrtf(x) <= FLOAT_EPS : (len <= FLOAT_EPS) <=> (sq <= epsSq)
  float poly6; // POLY6_COEFF
  float spiky; // SPIKY_COEFF
  float spikyK; // SPIKY_COEFF * -3.0f (one rounding)
  float maxVel; // MAX_VEL
  float effectRadiusSq; // EFFECT_RADIUS_SQUARED (boids)
  float nbrRadiusSq; // neighbour-list radius ((1 + margin) h)^2
  float nbrDmaxSq; // a list stays valid while every particle moved less than sqrt(nbrDmaxSq) = 0.45 margin h
};

// grid.cl:14-24 getCell3DIndexFromPos -- bit-exact: clamp, add, IEEE divide, floor, truncate
__device__ __forceinline__ int3 cell3D(const GridParams& g, float x, float y, float z)
{
  const float px = fadd(fclamp(x, -g.absW[0], g.absW[0]), g.absW[0]);
  const float py = fadd(fclamp(y, -g.absW[1], g.absW[1]), g.absW[1]);
  const float pz = fadd(fclamp(z, -g.absW[2], g.absW[2]), g.absW[2]);
  int3 c;
  c.x = (int)(u32)floorf(fdiv(px, g.cellSize));
  c.y = (int)(u32)floorf(fdiv(py, g.cellSize));
  c.z = (int)(u32)floorf(fdiv(pz, g.cellSize));
  return c;
}
// grid.cl:29-38 getCell1DIndexFromPos
__device__ __forceinline__ u32 cell1D(const GridParams& g, float x, float y, float z)
{
  const int3 c = cell3D(g, x, y, z);
  return (u32)c.x * (u32)g.res[2] * (u32)g.res[1] + (u32)c.y * (u32)g.res[2] + (u32)c.z;
}

enum Traversal
{
  TRAV_BOIDS = 0, // out-of-range cells skipped (boids.cl:84-88)
  TRAV_FLUIDS = 1, // modulo wrap (fluids.cl:107)
  TRAV_CLOUDS = 2 // modulo wrap + image shift in x/z, y skipped (clouds.cl:334-347)
};

// Visit the 27 neighbour cells of ci in the reference's order (iX, iY, iZ ascending) and call
// f(start, end, shiftX, shiftZ) with inclusive particle ranges taken from the table. The three z cells of one
// (iX, iY) column are fetched together and, because cell ids are z-fastest and particles are cell-sorted, usually
// form ONE contiguous index run: exactly-adjacent ranges with the same image shift are merged, which never changes
// the candidate sequence (quirks such as the cap or the start=1 cell simply break the merge).
template <int TRAV, typename F>
__device__ __forceinline__ void forEachNeighbourRun(const GridParams& g, const uint2* __restrict__ table, int3 ci, F&& f)
{
  const int RX = g.res[0], RY = g.res[1], RZ = g.res[2];
#pragma unroll 1
  for (int iX = -1; iX <= 1; ++iX)
  {
    int cx = ci.x + iX;
    float sx = 0.0f;
    if (TRAV == TRAV_BOIDS)
    {
      if (cx < 0 || cx >= RX)
        continue;
    }
    else
    {
      if (TRAV == TRAV_CLOUDS)
        sx = (cx >= RX) ? 2.0f * g.absW[0] : ((cx < 0) ? -2.0f * g.absW[0] : 0.0f);
      cx = (cx + RX) % RX;
    }
#pragma unroll 1
    for (int iY = -1; iY <= 1; ++iY)
    {
      int cy = ci.y + iY;
      if (TRAV == TRAV_FLUIDS)
        cy = (cy + RY) % RY;
      else if (cy < 0 || cy >= RY)
        continue;
      const uint2* __restrict__ row = table + (cx * RY + cy) * RZ;
      uint2 se[3];
      float sz[3];
#pragma unroll
      for (int k = 0; k < 3; ++k)
      {
        int cz = ci.z + k - 1;
        sz[k] = 0.0f;
        bool skip = false;
        if (TRAV == TRAV_BOIDS)
        {
          skip = cz < 0 || cz >= RZ;
        }
        else
        {
          if (TRAV == TRAV_CLOUDS)
            sz[k] = (cz >= RZ) ? 2.0f * g.absW[2] : ((cz < 0) ? -2.0f * g.absW[2] : 0.0f);
          cz = (cz + RZ) % RZ;
        }
        se[k] = skip ? make_uint2(1u, 0u) : __ldg(row + cz);
      }
      // merge exactly-adjacent non-empty ranges
      u32 cs = se[0].x, ce = se[0].y;
      float csz = sz[0];
#pragma unroll
      for (int k = 1; k < 3; ++k)
      {
        const bool curEmpty = cs > ce, nxtEmpty = se[k].x > se[k].y;
        if (nxtEmpty)
          continue;
        if (!curEmpty && ce + 1u == se[k].x && csz == sz[k])
        {
          ce = se[k].y;
        }
        else
        {
          if (!curEmpty)
            f(cs, ce, sx, csz);
          cs = se[k].x;
          ce = se[k].y;
          csz = sz[k];
        }
      }
      if (cs <= ce)
        f(cs, ce, sx, csz);
    }
  }
}

// Run body(e, P[e]) for e = start..end in ascending order, four positions fetched per group. (A software pipeline
// across groups was measured 2% slower: its rotating registers cost four MOVs per candidate in an issue-bound loop,
// and the resident warps hide an L1 hit anyway.)
template <typename Body>
__device__ __forceinline__ void forRangeLoad4(const float4* __restrict__ P, u32 start, u32 end, Body&& body)
{
  u32 e = start;
#pragma unroll 1
  for (; e + 3u <= end; e += 4u)
  {
    const float4 a0 = __ldg(P + e), a1 = __ldg(P + e + 1), a2 = __ldg(P + e + 2), a3 = __ldg(P + e + 3);
    body(e, a0);
    body(e + 1, a1);
    body(e + 2, a2);
    body(e + 3, a3);
  }
#pragma unroll 1
  for (; e <= end; ++e)
    body(e, __ldg(P + e));
}

__device__ __forceinline__ float4 ld4(const float4* __restrict__ p, u32 i) { return __ldg(p + i); }

} // namespace rtp
