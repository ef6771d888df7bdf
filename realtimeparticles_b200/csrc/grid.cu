// grid.cu -- reset / cell-table / camera / render-side kernels shared by the three models.
#include <math.h>

#include "kernels.cuh"
#include "sweep.cuh"

namespace rtp
{
constexpr int EW_THREADS = 256;
static inline int ewBlocks(size_t n) { return (int)((n + EW_THREADS - 1) / EW_THREADS); }

// resetCellIDs grid.cl:65-71 + resetCameraDist utils.cl:35-38 (+ identity permutations for the never-moving tail)
__global__ void __launch_bounds__(EW_THREADS) resetIdsKernel(u32* __restrict__ cellID, u32* __restrict__ cameraDist,
    u32* __restrict__ perm, u32* __restrict__ cameraPerm, u32 M, u32 numCells)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i >= M)
    return;
  cellID[i] = numCells * 2u + i;
  cameraDist[i] = (u32)(RTP_FAR_DIST);
  perm[i] = i;
  cameraPerm[i] = i;
}

// adjustEndCell grid.cl:143-152
__global__ void __launch_bounds__(EW_THREADS) adjustEndCellKernel(uint2* __restrict__ table, u32 numCells, u32 cap)
{
  RTP_PDL_PROLOGUE();
  const u32 c = blockIdx.x * EW_THREADS + threadIdx.x;
  if (c >= numCells)
    return;
  const uint2 se = table[c];
  if (se.y > se.x)
    table[c] = make_uint2(se.x, se.x + min(se.y - se.x, cap));
}

// fillCameraDist utils.cl:43-52: key = (uint) max(FAR_DIST - length(pos - cam) * 100, 0)
__global__ void __launch_bounds__(EW_THREADS) fillCameraDistKernel(const float4* __restrict__ pos, float cx, float cy, float cz,
    u32* __restrict__ keys, u32 N)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i >= N)
    return;
  const float4 p = pos[i];
  const float dx = fsub(p.x, cx), dy = fsub(p.y, cy), dz = fsub(p.z, cz);
  const float len = fsqrt(dot3c(dx, dy, dz, dx, dy, dz));
  keys[i] = (u32)(fmaxf(fsub(RTP_FAR_DIST, fmul(len, 100.0f)), 0.0f));
}

// payload gather of the camera sort: {p_pos,p_col,p_vel,p_predPos} (+ 5 float arrays for clouds)
// Boids.cpp:381, Fluids.cpp:468, Clouds.cpp:624. Output goes to the B buffers; the caller copies back.
__global__ void __launch_bounds__(EW_THREADS) cameraGatherKernel(DeviceState s, int model, const float4* __restrict__ pred,
    float4* __restrict__ predOut)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i >= s.N)
    return;
  const u32 j = s.cameraPerm[i];
  s.posB[i] = s.posA[j];
  s.colB[i] = s.col[j];
  s.velB[i] = s.velA[j];
  if (model == RTP_MODEL_BOIDS)
  {
    s.velC[i] = s.acc[j]; // p_acc travels too (Boids.cpp:381); velC is free scratch here
  }
  else
  {
    predOut[i] = pred[j];
    if (model == RTP_MODEL_CLOUDS)
    {
      s.tempB[i] = s.tempA[j];
      s.buoyB[i] = s.buoyA[j];
      s.vaporB[i] = s.vaporA[j];
      s.cloudB[i] = s.cloudA[j];
      s.partIdB[i] = s.partIdA[j];
    }
  }
}

// resetGridDetector / fillGridDetector grid.cl:43-60 (float8 per cell)
__global__ void __launch_bounds__(EW_THREADS) resetGridDetectorKernel(float4* __restrict__ det, u32 n4)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i < n4)
    det[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
__global__ void __launch_bounds__(EW_THREADS) fillGridDetectorKernel(const float4* __restrict__ pos, GridParams g,
    float4* __restrict__ det, u32 N)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i >= N)
    return;
  const float4 p = pos[i];
  const u32 c = cell1D(g, p.x, p.y, p.z);
  if (c < g.numCells)
  {
    det[2 * c] = make_float4(1.f, 1.f, 1.f, 1.f);
    det[2 * c + 1] = make_float4(1.f, 1.f, 1.f, 1.f);
  }
}

// fld_fillFluidColor fluids.cl:458-479
__global__ void __launch_bounds__(EW_THREADS) fillFluidColorKernel(const float* __restrict__ density, float restDensity,
    float4* __restrict__ col, u32 N)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i >= N)
    return;
  const float constraint = fsub(1.0f, fdiv(density[i], restDensity));
  float4 color = make_float4(0.0f, 0.1f, 1.0f, 0.5f);
  if (constraint > 0.0f)
  {
    // (lightBlue - blue) = (0.7, 0.6, 0, 0)
    color.x = fadd(color.x, fdiv(fmul(fsub(0.7f, 0.0f), constraint), 0.35f));
    color.y = fadd(color.y, fdiv(fmul(fsub(0.7f, 0.1f), constraint), 0.35f));
    color.z = fadd(color.z, fdiv(fmul(fsub(1.0f, 1.0f), constraint), 0.35f));
    color.w = fadd(color.w, fdiv(fmul(fsub(0.5f, 0.5f), constraint), 0.35f));
  }
  else if (constraint < 0.0f)
  {
    // (blue - darkBlue) = (0, 0.1, 0.2, 0)
    color.x = fadd(color.x, fdiv(fmul(fsub(0.0f, 0.0f), constraint), 0.35f));
    color.y = fadd(color.y, fdiv(fmul(fsub(0.1f, 0.0f), constraint), 0.35f));
    color.z = fadd(color.z, fdiv(fmul(fsub(1.0f, 0.8f), constraint), 0.35f));
    color.w = fadd(color.w, fdiv(fmul(fsub(0.5f, 0.5f), constraint), 0.35f));
  }
  col[i] = color;
}

// fillColorFloat utils.cl:65-78
__global__ void __launch_bounds__(EW_THREADS) fillColorFloatKernel(const float* __restrict__ q, float minVal, float maxVal,
    float4* __restrict__ col, u32 N)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i >= N)
    return;
  float val = fdiv(fsub(q[i], minVal), fsub(maxVal, minVal));
  val = fmul(val, (val < 0.0f) ? 0.0f : 1.0f);
  val = fmul(val, (1.0f < val) ? 0.0f : 1.0f);
  col[i] = make_float4(val, val, val, val);
}

// self-test of the range-check-free sqrt / reciprocal (sweep.cuh) against the IEEE intrinsics, over every float whose
// bit pattern lies in [lo, hi]
__global__ void __launch_bounds__(EW_THREADS) selftestMathKernel(u32 lo, u32 hi, unsigned long long* __restrict__ bad)
{
  RTP_PDL_PROLOGUE();
  unsigned long long badSqrt = 0, badRcp = 0;
  for (unsigned long long b = (unsigned long long)lo + blockIdx.x * (unsigned long long)EW_THREADS + threadIdx.x; b <= hi;
       b += (unsigned long long)gridDim.x * EW_THREADS)
  {
    const float x = __uint_as_float((u32)b);
    badSqrt += __float_as_uint(sqrtInRange(x)) != __float_as_uint(__fsqrt_rn(x));
    badRcp += __float_as_uint(rcpInRange(x)) != __float_as_uint(__frcp_rn(x));
  }
  if (badSqrt)
    atomicAdd(bad, badSqrt);
  if (badRcp)
    atomicAdd(bad + 1, badRcp);
}
void launchSelftestMath(u32 lo, u32 hi, unsigned long long* bad, cudaStream_t st)
{
  launchKernel(selftestMathKernel, 148 * 8, EW_THREADS, st, lo, hi, bad);
}

// slab decomposition: drop the ghost copies after a step. Keys = "is ghost" (1 bit) -> one stable radix pass gives the
// owned particles first, in their cell-sorted order; then the state is gathered through that permutation.
__global__ void __launch_bounds__(EW_THREADS) ghostFlagKernel(const u32* __restrict__ perm, const float4* __restrict__ pos, u32 nOwned,
    u32* __restrict__ keys, u32 N)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i < N)
    keys[i] = (perm[i] >= nOwned || !isfinite(pos[i].x)) ? 1u : 0u; // (the last sweep left +inf in the rows it skipped)
}
__global__ void __launch_bounds__(EW_THREADS) compactGatherKernel(DeviceState s, const u32* __restrict__ order, u32 n)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i >= n)
    return;
  const u32 j = order[i];
  s.posB[i] = s.posA[j];
  s.velB[i] = s.velA[j];
}
void launchGhostFlags(const DeviceState& s, u32* keysOut, cudaStream_t st)
{
  if (s.N)
    launchKernel(ghostFlagKernel, ewBlocks(s.N), EW_THREADS, st, s.perm, s.posA, s.nOwned, keysOut, s.N);
}
void launchCompactGather(const DeviceState& s, const u32* order, u32 n, cudaStream_t st)
{
  if (n)
    launchKernel(compactGatherKernel, ewBlocks(n), EW_THREADS, st, s, order, n);
}

// ---- slab decomposition: exchange buffers. out[k] = buf[idx[k]] (idx 0xFFFFFFFF = no particle: `fill`), and back.
template <typename T>
__global__ void __launch_bounds__(EW_THREADS) packRowsKernel(const T* __restrict__ buf, const u32* __restrict__ idx, u32 n, T* __restrict__ out, T fill)
{
  RTP_PDL_PROLOGUE();
  const u32 k = blockIdx.x * EW_THREADS + threadIdx.x;
  if (k >= n)
    return;
  const u32 j = idx[k];
  out[k] = j == 0xFFFFFFFFu ? fill : buf[j];
}
template <typename T>
__global__ void __launch_bounds__(EW_THREADS) unpackRowsKernel(T* __restrict__ buf, const u32* __restrict__ idx, u32 n, const T* __restrict__ in)
{
  RTP_PDL_PROLOGUE();
  const u32 k = blockIdx.x * EW_THREADS + threadIdx.x;
  if (k >= n)
    return;
  const u32 j = idx[k];
  if (j != 0xFFFFFFFFu)
    buf[j] = in[k];
}
__global__ void __launch_bounds__(EW_THREADS) inversePermKernel(const u32* __restrict__ perm, u32 n, u32* __restrict__ inv)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i < n)
    inv[perm[i]] = i;
}
// list validity across slabs: a ghost moved by its owner further than the bound invalidates the lists of the next epoch
__global__ void __launch_bounds__(EW_THREADS) ghostDisplacementKernel(const float4* __restrict__ pred, const float4* __restrict__ buildPos,
    const u32* __restrict__ idx, u32 n, float dmaxSq, u32* __restrict__ invalid)
{
  RTP_PDL_PROLOGUE();
  const u32 k = blockIdx.x * EW_THREADS + threadIdx.x;
  if (k >= n)
    return;
  const u32 i = idx[k];
  if (i == 0xFFFFFFFFu)
    return;
  const float4 p = pred[i], b = buildPos[i];
  const float dx = p.x - b.x, dy = p.y - b.y, dz = p.z - b.z;
  if (!(dx * dx + dy * dy + dz * dz <= dmaxSq) && isfinite(p.x))
    *invalid = 1u;
}
// interior rows of a slab (kernels.cuh: rowPhaseBounds) = the sorted rows of the cells [cellLo, cellHi): bounds[0] = first
// row, bounds[1] = one past the last one; bounds[2] = one past the last row that holds a particle (any cell). Pre-set to
// 0xFFFFFFFF / 0 / 0 by the launcher (an empty interior, no particle); runs on the table BEFORE adjustEndCell caps the ends.
// bounds[3] is the sticky error flag of the launches by row phase (sweep.cuh: a grid too small for its rows).
__global__ void __launch_bounds__(EW_THREADS) rowPhaseBoundsKernel(const uint2* __restrict__ table, u32 scanLo, u32 scanHi, u32 cellLo, u32 cellHi,
    u32* __restrict__ bounds)
{
  RTP_PDL_PROLOGUE();
  const u32 c = scanLo + blockIdx.x * EW_THREADS + threadIdx.x;
  u32 lo = 0xFFFFFFFFu, hi = 0u, last = 0u;
  if (c < scanHi)
  {
    const uint2 se = table[c];
    if (se.y >= se.x) // (start, LAST index) of a cell that holds particles
    {
      last = se.y + 1u;
      if (c >= cellLo && c < cellHi)
        lo = se.x, hi = se.y + 1u;
    }
  }
  lo = __reduce_min_sync(0xFFFFFFFFu, lo);
  hi = __reduce_max_sync(0xFFFFFFFFu, hi);
  last = __reduce_max_sync(0xFFFFFFFFu, last);
  if ((threadIdx.x & 31u) == 0u && last != 0u)
  {
    if (hi != 0u)
    {
      atomicMin(bounds, lo);
      atomicMax(bounds + 1, hi);
    }
    atomicMax(bounds + 2, last);
  }
}
// scanLo / scanHi: the cells that can hold a row of this handle at all (the slab's layers and its ghost layers)
void launchRowPhaseBounds(const DeviceState& s, u32 scanLo, u32 scanHi, u32 cellLo, u32 cellHi, u32* bounds, cudaStream_t st)
{
  cudaMemsetAsync(bounds, 0xFF, sizeof(u32), st);
  cudaMemsetAsync(bounds + 1, 0, 2 * sizeof(u32), st);
  if (cellHi > cellLo && scanHi > scanLo)
    launchKernel(rowPhaseBoundsKernel, ewBlocks(scanHi - scanLo), EW_THREADS, st, (const uint2*)s.table, scanLo, scanHi, cellLo, cellHi, bounds);
}

// slab decomposition: which rows hold a particle whose cell x-layer (key / plane, the +x wall index clamped into the last
// layer) lies below cutLo / at or above cutHi -- the leavers of a migration, the face layers of a halo. One pass over
// keys and positions; the caller compacts the two byte masks (order-preserving, deterministic).
__global__ void __launch_bounds__(EW_THREADS) classifyRowsKernel(const u32* __restrict__ keys, const float4* __restrict__ pos, u32 n, u32 plane,
    u32 lastLayer, u32 cutLo, u32 cutHi, unsigned char* __restrict__ below, unsigned char* __restrict__ above)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i >= n)
    return;
  const bool alive = isfinite(pos[i].x);
  const u32 layer = min(keys[i] / plane, lastLayer);
  below[i] = alive && layer < cutLo;
  above[i] = alive && layer >= cutHi;
}
void launchClassifyRows(const DeviceState& s, const GridParams& g, const u32* keys, u32 n, u32 cutLo, u32 cutHi, unsigned char* below,
    unsigned char* above, cudaStream_t st)
{
  if (n)
    launchKernel(classifyRowsKernel, ewBlocks(n), EW_THREADS, st, keys, (const float4*)s.posA, n, (u32)(g.res[1] * g.res[2]), (u32)(g.res[0] - 1), cutLo,
        cutHi, below, above);
}

// rows that hold no particle any more (it migrated): +inf position, zero velocity -- and the prediction / cell id
// RTP_SHARD_PREDICT computes for such a row (the leavers are cleared after the prediction that found them)
__global__ void __launch_bounds__(EW_THREADS) clearRowsKernel(float4* __restrict__ pos, float4* __restrict__ vel, float4* __restrict__ pred,
    u32* __restrict__ keys, GridParams g, const u32* __restrict__ idx, u32 n)
{
  RTP_PDL_PROLOGUE();
  const u32 k = blockIdx.x * EW_THREADS + threadIdx.x;
  if (k >= n)
    return;
  const u32 j = idx[k];
  if (j != 0xFFFFFFFFu)
  {
    pos[j] = make_float4(INFINITY, INFINITY, INFINITY, 0.0f);
    vel[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    pred[j] = make_float4(INFINITY, INFINITY, INFINITY, 0.0f);
    keys[j] = cell1D(g, INFINITY, INFINITY, INFINITY);
  }
}
void launchClearRows(const DeviceState& s, const GridParams& g, u32* keys, const u32* idx, u32 n, cudaStream_t st)
{
  if (n)
    launchKernel(clearRowsKernel, ewBlocks(n), EW_THREADS, st, s.posA, s.velA, s.pred0, keys, g, idx, n);
}
void launchPackRows(const void* buf, int rowBytes, const u32* idx, u32 n, void* out, cudaStream_t st)
{
  if (!n)
    return;
  const float inf = INFINITY;
  if (rowBytes == 16)
    launchKernel(packRowsKernel<float4>, ewBlocks(n), EW_THREADS, st, (const float4*)buf, idx, n, (float4*)out, make_float4(inf, inf, inf, 0.0f));
  else
    launchKernel(packRowsKernel<float>, ewBlocks(n), EW_THREADS, st, (const float*)buf, idx, n, (float*)out, 0.0f);
}
void launchUnpackRows(void* buf, int rowBytes, const u32* idx, u32 n, const void* in, cudaStream_t st)
{
  if (!n)
    return;
  if (rowBytes == 16)
    launchKernel(unpackRowsKernel<float4>, ewBlocks(n), EW_THREADS, st, (float4*)buf, idx, n, (const float4*)in);
  else
    launchKernel(unpackRowsKernel<float>, ewBlocks(n), EW_THREADS, st, (float*)buf, idx, n, (const float*)in);
}
void launchInversePerm(const DeviceState& s, u32* inv, cudaStream_t st)
{
  if (s.N)
    launchKernel(inversePermKernel, ewBlocks(s.N), EW_THREADS, st, s.perm, s.N, inv);
}
void launchGhostDisplacement(const DeviceState& s, const float4* pred, const u32* idx, u32 n, float dmaxSq, u32* invalid, cudaStream_t st)
{
  if (n)
    launchKernel(ghostDisplacementKernel, ewBlocks(n), EW_THREADS, st, pred, (const float4*)s.nbrBuildPos, idx, n, dmaxSq, invalid);
}

void launchResetIds(const DeviceState& s, u32 numCells, cudaStream_t st)
{
  launchKernel(resetIdsKernel, ewBlocks(s.M), EW_THREADS, st, s.cellID, s.cameraDist, s.perm, s.cameraPerm, s.M, numCells);
}
void launchAdjustEndCell(const DeviceState& s, const GridParams& g, cudaStream_t st, u32 cellLo, u32 cellHi)
{
  // (a slab only looks at the cells that can hold one of its rows)
  cellHi = min(cellHi, g.numCells);
  if (cellHi > cellLo)
    launchKernel(adjustEndCellKernel, ewBlocks(cellHi - cellLo), EW_THREADS, st, s.table + cellLo, cellHi - cellLo, g.maxPartsInCell);
}
void launchFillCameraDist(const DeviceState& s, const float cam[3], u32* keysOut, cudaStream_t st)
{
  if (s.N)
    launchKernel(fillCameraDistKernel, ewBlocks(s.N), EW_THREADS, st, s.posA, cam[0], cam[1], cam[2], keysOut, s.N);
}
void launchCameraGather(const DeviceState& s, int model, const float4* pred, float4* predOut, cudaStream_t st)
{
  if (s.N)
    launchKernel(cameraGatherKernel, ewBlocks(s.N), EW_THREADS, st, s, model, pred, predOut);
}
void launchGridDetector(const DeviceState& s, const GridParams& g, cudaStream_t st)
{
  launchKernel(resetGridDetectorKernel, ewBlocks((size_t)g.numCells * 2), EW_THREADS, st, (float4*)s.partDetector, g.numCells * 2);
  if (s.N)
    launchKernel(fillGridDetectorKernel, ewBlocks(s.N), EW_THREADS, st, s.posA, g, (float4*)s.partDetector, s.N);
}
void launchFillFluidColor(const DeviceState& s, float restDensity, cudaStream_t st)
{
  if (s.N)
    launchKernel(fillFluidColorKernel, ewBlocks(s.N), EW_THREADS, st, s.density, restDensity, s.col, s.N);
}
void launchFillColorFloat(const DeviceState& s, const float* quantity, float minVal, float maxVal, cudaStream_t st)
{
  if (s.N)
    launchKernel(fillColorFloatKernel, ewBlocks(s.N), EW_THREADS, st, quantity, minVal, maxVal, s.col, s.N);
}

} // namespace rtp
