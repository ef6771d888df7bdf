// fluids.cu -- Position Based Fluids solver and its clouds extension on the uniform grid.
// Reference: physics/ocl/kernels/{sph,fluids,clouds}.cl; host sequences Fluids::update (physics/ocl/Fluids.cpp:400-471)
// and Clouds::update (physics/ocl/Clouds.cpp:503-627).
//
// Launches per fluids step with I Jacobi iterations (reference: 3 + 5 I + 5 + buffer copies around two sorts):
//   fluidPredictKernel      fld_predictPosition (fluids.cl:62-74) + fillCellIDs on p_predPos (grid.cl:76-86)
//                           + resetStartEndCell (grid.cl:91-96) + the digit histograms of the cell sort
//   [sort passes]           sort.cu
//   fluidGatherKernel       payload permutation of p_pos, p_vel, p_predPos (radixSort.cl:179-190) + first
//                           fld_applyBoundaryCondition (fluids.cl:435-439) + fillStartCell/fillEndCell (grid.cl:101-138)
//   [adjustEndCell]         grid.cu
//   I x densityLambdaKernel fld_computeDensity (:80-126) + fld_computeConstraintFactor (:131-193) in ONE sweep
//   I x correctionKernel    fld_computeConstraintCorrection (:198-247) + fld_correctPosition (:252-258) + the next
//                           iteration's boundary clamp; the last one also does fld_updateVel (:263-273)
//   vorticityKernel         fld_computeVorticity (:278-324)
//   confinementKernel       fld_applyVorticityConfinement (:329-377), writes the copy p_velInViscosity (Fluids.cpp:451)
//   xsphKernel              fld_applyXsphViscosityCorrection (:383-430) + fld_updatePosition (:444-450)
// The clouds variants use the same kernels with the periodic-image traversal (clouds.cl:334-347).
//
// Numerics: bit-exact with the oracle everywhere. Hit tests, element-wise stages and the per-pair terms use the
// canonical operation sequence (rtp_common.cuh, DESIGN.md "Canonical arithmetic"); sums run in the reference's
// order (27 cells, ascending e, one fp32 accumulator per component).
#include <math.h>

#include <algorithm>
#include <cstdlib>

#include "kernels.cuh"
#include "sort.cuh"
#include "sweep.cuh"

namespace rtp
{
#ifndef RTP_NB_THREADS
#define RTP_NB_THREADS 128
#endif
constexpr int NB_THREADS = RTP_NB_THREADS; // neighbour kernels
#ifndef RTP_NB_WARPS_PER_SM
#define RTP_NB_WARPS_PER_SM 28 // resident warps per SM the neighbour kernels are compiled for (28: 72 registers per thread)
#endif
constexpr int NB_MIN_BLOCKS = RTP_NB_WARPS_PER_SM * 32 / RTP_NB_THREADS;
constexpr int EW_THREADS = 256; // element-wise kernels

__device__ __forceinline__ void fillCellTable(const u32* __restrict__ keys, u32 i, u32 N, u32 numCells, uint2* __restrict__ table)
{
  const u32 id = keys[i];
  if (id < numCells)
  {
    if (i > 0 && id != keys[i - 1]) // fillStartCell grid.cl:101-117
      table[id].x = i;
    const u32 next = (i + 1 < N) ? keys[i + 1] : 0xFFFFFFFFu; // fillEndCell grid.cl:122-138
    if (id != next)
      table[id].y = i;
  }
}

// fld_applyBoundaryCondition fluids.cl:435-439
__device__ __forceinline__ float4 fluidBoundary(const GridParams& g, float4 p)
{
  p.x = fclamp(p.x, fadd(-g.absW[0], 0.01f), fsub(g.absW[0], 0.1f));
  p.y = fclamp(p.y, fadd(-g.absW[1], 0.01f), fsub(g.absW[1], 0.1f));
  p.z = fclamp(p.z, fadd(-g.absW[2], 0.01f), fsub(g.absW[2], 0.1f));
  p.w = 0.0f;
  return p;
}
// cld_applyMixedBoundaryConditions clouds.cl:279-299
__device__ __forceinline__ float4 cloudBoundary(const GridParams& g, float4 np)
{
  float4 p = np;
  const float cx = fclamp(np.x, -g.absW[0], g.absW[0]), cy = fclamp(np.y, -g.absW[1], g.absW[1]), cz = fclamp(np.z, -g.absW[2], g.absW[2]);
  if (fabsf(np.x) > g.absW[0]) p.x = fsub(np.x, fmul(2.0f, cx));
  if (fabsf(np.y) > g.absW[1]) p.y = cy;
  if (fabsf(np.z) > g.absW[2]) p.z = fsub(np.z, fmul(2.0f, cz));
  return p;
}

// ------------------------------------------------------------------ element-wise kernels

// HIST: also build the digit histograms of the cell sort on the way and zero its look-back words (saves the sort's own
// histogram kernel and one read of the keys; the caller has zeroed the sort control block before the launch)
template <bool HIST>
__global__ void __launch_bounds__(EW_THREADS) fluidPredictKernel(DeviceState s, GridParams g, float dt, u32* __restrict__ keys,
    int passes, PassDesc desc, u32* __restrict__ sortCtrl, u32* __restrict__ sortStatus, size_t statusWords, int resets)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * EW_THREADS + threadIdx.x;
  if (resets && i < g.numCells)
    s.table[i] = make_uint2(1u, 0u);
  if (resets && i < NBR_EPOCHS && s.nbrInvalid)
  {
    s.nbrInvalid[i] = s.nbrInvalid[NBR_EPOCHS + i] = 0u; // (own particles / ghosts of a slab, sweep.cuh)
    s.stragCount[i] = s.stragCount[NBR_EPOCHS + i] = 0u; // (two queue classes, sweep.cuh)
    s.stragCursor[i] = s.stragCursor[NBR_EPOCHS + i] = 0u;
  }
  const bool valid = i < s.N;
  u32 key = 0u;
  if (valid)
  {
    const float4 p = s.posA[i], v = s.velA[i];
    // newVel = vel + GRAVITY_ACC * dt ; predPos = pos + newVel * dt
    const float nvx = fadd(v.x, fmul(0.0f, dt)), nvy = fadd(v.y, fmul(-RTP_ABS_GRAVITY_ACC_Y, dt)), nvz = fadd(v.z, fmul(0.0f, dt));
    const float4 pr = make_float4(fadd(p.x, fmul(nvx, dt)), fadd(p.y, fmul(nvy, dt)), fadd(p.z, fmul(nvz, dt)), 0.0f);
    s.pred0[i] = pr;
    key = cell1D(g, pr.x, pr.y, pr.z);
    keys[i] = key;
  }
  if (HIST)
  {
    __shared__ u32 sHist[SORT_MAX_PASSES * SORT_RADIX];
    sortStatusClear(sortStatus, statusWords);
    sortHistogramAdd(sHist, key, valid, passes, desc, sortCtrl);
  }
}

__global__ void __launch_bounds__(EW_THREADS) fluidGatherKernel(DeviceState s, GridParams g)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i >= s.N)
    return;
  const u32 j = s.perm[i];
  s.posB[i] = s.posA[j];
  s.velB[i] = s.velA[j];
  const float4 pr = s.pred0[j];
  // (slab decomposition: a "no particle" row keeps its +inf position instead of being clamped into the box)
  s.pred1[i] = (s.nOwned != 0xFFFFFFFFu && !isfinite(pr.x)) ? pr : fluidBoundary(g, pr);
  fillCellTable(s.cellID, i, s.N, g.numCells, s.table);
}

// canonical exp of the path (DESIGN.md "Canonical arithmetic"): 2^n * P(r), n = rint(x log2 e), r = x - n ln2 in two
// fused steps, P = Taylor-Horner (degree 7 float / 13 double), every step an explicit fma -- identical to the oracle's.
__device__ __forceinline__ float canonExpf(float x)
{
  x = fminf(fmaxf(x, -87.0f), 88.0f);
  const float n = rintf(fmul(x, 1.44269504f));
  float r = ffma(n, -0.693145751953125f, x);
  r = ffma(n, -1.428606765330187e-06f, r);
  float p = 1.98412698e-4f;
  p = ffma(p, r, 1.38888889e-3f);
  p = ffma(p, r, 8.33333333e-3f);
  p = ffma(p, r, 4.16666667e-2f);
  p = ffma(p, r, 1.66666667e-1f);
  p = ffma(p, r, 0.5f);
  p = ffma(p, r, 1.0f);
  p = ffma(p, r, 1.0f);
  return fmul(p, __int_as_float(((int)n + 127) << 23));
}
__device__ __forceinline__ double canonExp(double x)
{
  x = fmin(fmax(x, -700.0), 700.0);
  const double n = rint(__dmul_rn(x, 1.4426950408889634));
  double r = __fma_rn(n, -6.93147180369123816490e-01, x);
  r = __fma_rn(n, -1.90821492927058770002e-10, r);
  double p = 1.0 / 6227020800.0;
  p = __fma_rn(p, r, 1.0 / 479001600.0);
  p = __fma_rn(p, r, 1.0 / 39916800.0);
  p = __fma_rn(p, r, 1.0 / 3628800.0);
  p = __fma_rn(p, r, 1.0 / 362880.0);
  p = __fma_rn(p, r, 1.0 / 40320.0);
  p = __fma_rn(p, r, 1.0 / 5040.0);
  p = __fma_rn(p, r, 1.0 / 720.0);
  p = __fma_rn(p, r, 1.0 / 120.0);
  p = __fma_rn(p, r, 1.0 / 24.0);
  p = __fma_rn(p, r, 1.0 / 6.0);
  p = __fma_rn(p, r, 0.5);
  p = __fma_rn(p, r, 1.0);
  p = __fma_rn(p, r, 1.0);
  return __dmul_rn(p, __longlong_as_double(((long long)n + 1023) << 52));
}

// clouds.cl:72-99
__device__ __forceinline__ float environmentTemp(const GridParams& g, float alt) { return fadd(fmul(-3.5f, fadd(alt, g.absW[1])), 293.0f); }
__device__ __forceinline__ float externalHeatSource(const GridParams& g, float alt)
{
  return fclamp(canonExpf(fdiv(-fadd(alt, g.absW[1]), 3.0f)), 0.0f, 1.0f);
}
__device__ __forceinline__ float saturationVaporDensity(float T)
{
  return (float)__ddiv_rn(__dmul_rn(217.0, canonExp(__dsub_rn(19.5, __ddiv_rn(4303.4, __dsub_rn((double)T, 29.5))))), (double)T);
}

// cld_initTemperature clouds.cl:116-122 + cld_initVaporDensity :127-135 over M
__global__ void __launch_bounds__(EW_THREADS) cloudsInitFieldsKernel(DeviceState s, GridParams g, float coeff)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i >= s.M)
    return;
  const float T = environmentTemp(g, s.posA[i].y);
  s.tempA[i] = T;
  s.vaporA[i] = fmul(coeff, saturationVaporDensity(T));
}

// Clouds.cpp:514-541 in one pass: 6 thermodynamics kernels + 3 buffer copies (clouds.cl:159-252), cld_predictPosition
// (:257-273), cld_applyMixedBoundaryConditions (:279-299), fillCellIDs, resetStartEndCell.
__global__ void __launch_bounds__(EW_THREADS) cloudsThermoPredictKernel(DeviceState s, GridParams g, rtp_cloud_params c,
    u32* __restrict__ keys)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i < g.numCells)
    s.table[i] = make_uint2(1u, 0u);
  if (i < NBR_EPOCHS && s.nbrInvalid)
  {
    s.nbrInvalid[i] = s.nbrInvalid[NBR_EPOCHS + i] = 0u; // (own particles / ghosts of a slab, sweep.cuh)
    s.stragCount[i] = s.stragCount[NBR_EPOCHS + i] = 0u; // (two queue classes, sweep.cuh)
    s.stragCursor[i] = s.stragCursor[NBR_EPOCHS + i] = 0u;
  }
  if (i >= s.N)
    return;
  const float4 p = s.posA[i], v = s.velA[i];
  const float dt = c.timeStep;
  // cld_heatFromGround
  float temp = fminf(fadd(s.tempA[i], fmul(fmul(externalHeatSource(g, p.y), c.groundHeatCoeff), dt)), 313.0f);
  // cld_computeBuoyancy (cloud density before the phase transition)
  const float envTemp = environmentTemp(g, p.y);
  const float cloudIn = s.cloudA[i], vaporIn = s.vaporA[i];
  const float buoy = fsub(fdiv(fmul(c.buoyancyCoeff, fsub(temp, envTemp)), envTemp), fmul(fmul(c.gravCoeff, RTP_ABS_GRAVITY_ACC_Y), cloudIn));
  // cld_applyAdiabaticCooling
  const float tempIn = fmaxf(fsub(temp, fmul(fmul(c.adiabaticLapseRate, v.y), dt)), 223.0f);
  // cld_generateCloud
  const float gen = fmul(c.phaseTransitionRate, fsub(vaporIn, saturationVaporDensity(tempIn)));
  // cld_applyPhaseTransition
  s.cloudA[i] = fmaxf(fadd(cloudIn, fmul(gen, dt)), 0.0f);
  s.vaporA[i] = fmaxf(fsub(vaporIn, fmul(gen, dt)), 0.0f);
  // cld_applyLatentHeat
  temp = fadd(tempIn, fmaxf(fmul(fmul(c.latentHeatCoeff, gen), dt), 0.0f));
  s.tempA[i] = temp;
  s.buoyA[i] = buoy;
  s.cloudGen[i] = gen;
  // cld_predictPosition
  const float pvx = fadd(v.x, fmul(0.0f, dt)), pvy = fadd(v.y, fmul(buoy, dt)), pvz = fadd(v.z, fmul(0.0f, dt));
  const float4 tot = make_float4(fmul(pvx, dt), fmul(pvy, dt), fmul(pvz, dt), 0.0f);
  s.totCorrA[i] = tot;
  float4 pr = make_float4(fadd(p.x, tot.x), fadd(p.y, tot.y), fadd(p.z, tot.z), 0.0f);
  pr = cloudBoundary(g, pr);
  s.pred0[i] = pr;
  keys[i] = cell1D(g, pr.x, pr.y, pr.z);
}

__global__ void __launch_bounds__(EW_THREADS) cloudsGatherKernel(DeviceState s, GridParams g)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i >= s.N)
    return;
  const u32 j = s.perm[i];
  s.posB[i] = s.posA[j];
  s.velB[i] = s.velA[j];
  s.pred1[i] = s.pred0[j];
  s.totCorrB[i] = s.totCorrA[j];
  s.tempB[i] = s.tempA[j];
  s.buoyB[i] = s.buoyA[j];
  s.vaporB[i] = s.vaporA[j];
  s.cloudB[i] = s.cloudA[j];
  s.partIdB[i] = s.partIdA[j];
  fillCellTable(s.cellID, i, s.N, g.numCells, s.table);
}

// cld_updatePosition clouds.cl:942-953 + write the sorted state back to the canonical buffers
__global__ void __launch_bounds__(EW_THREADS) cloudsFinishKernel(DeviceState s, GridParams g, rtp_cloud_params c,
    const float4* __restrict__ pred, int copyVel)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * EW_THREADS + threadIdx.x;
  if (i >= s.N)
    return;
  const float4 pp = pred[i];
  float4 p = pp;
  const float a = fadd(pp.y, g.absW[1]);
  p.x = fadd(p.x, fmul(fmul(fmul(fsub(1.0f, canonExpf(fmul(-a, 0.2f))), c.windCoeff), c.timeStep), (float)(c.dim - 2)));
  p.z = fadd(p.z, fmul(fmul(fmul(fsub(1.0f, canonExpf(fmul(-a, 0.3f))), 0.7f), c.windCoeff), c.timeStep));
  s.posA[i] = p;
  if (copyVel)
    s.velA[i] = s.velB[i];
  s.tempA[i] = s.tempB[i];
  s.buoyA[i] = s.buoyB[i];
  s.vaporA[i] = s.vaporB[i];
  s.cloudA[i] = s.cloudB[i];
  s.partIdA[i] = s.partIdB[i];
  s.totCorrA[i] = s.totCorrB[i];
}

// ------------------------------------------------------------------ neighbour kernels (engine: sweep.cuh)

template <int TRAV>
__global__ void __launch_bounds__(NB_THREADS, NB_MIN_BLOCKS) densityLambdaKernel(DeviceState s, GridParams g, SphConsts c, float rho0, float cfm,
    const float4* __restrict__ pred, int nbrMode, int epoch)
{
  RTP_PDL_PROLOGUE();
  const u32 row0 = ctaFirstRow(s);
  if (row0 == NO_ROWS)
    return;
  recordGhostBuildPos(s, row0, pred, nbrMode, epoch);
  producerLoop(s, row0, pred, nbrMode, epoch,
      [&](const u32 i, const bool strag) -> int
      {
        const float4 pi = pred[i];
        float density = 0.f, gx = 0.f, gy = 0.f, gz = 0.f, sumG2 = 0.f;
        const int r = sweepProducer<TRAV>(g, c, s, pred, pi, i, nbrMode, epoch, strag,
            [&](u32, float dx, float dy, float dz, float sq)
            {
              const float cs = spikyCoefOrZero(c, sq);
              return PairTerm<6> { { fmul(c.poly6, poly6nc(c, sq)), cs, dx, dy, dz, fmul(fmul(cs, cs), sq) } };
            },
            [&](const PairTerm<6>& t)
            {
              density = fadd(density, t.v[0]);
              gx = ffma(t.v[2], t.v[1], gx);
              gy = ffma(t.v[3], t.v[1], gy);
              gz = ffma(t.v[4], t.v[1], gz);
              sumG2 = fadd(sumG2, t.v[5]);
            });
        if (r != SWEEP_DONE)
          return r;
        s.density[i] = density;
        // fluids.cl:189-192
        const float densityC = fsub(fdiv(density, rho0), 1.0f);
        float ssg = fadd(sumG2, dot3c(gx, gy, gz, gx, gy, gz));
        ssg = fdiv(ssg, fmul(rho0, rho0));
        s.lambda[i] = fdiv(-densityC, fadd(ssg, cfm));
        return (int)SWEEP_DONE;
      });
}

// The list BUILD of a step (first Jacobi iteration) runs as two kernels (tilebuild.cuh): the block-cooperative filter
// (CTA tiles staged by TMA, candidates within the list radius -> margin mask) and the same sweep as above walking that mask.
template <int TRAV>
__global__ void __launch_bounds__(NB_THREADS, NB_MIN_BLOCKS) marginMaskKernel(DeviceState s, GridParams g, SphConsts c, const float4* __restrict__ P)
{
  RTP_PDL_PROLOGUE();
  const u32 row0 = ctaFirstRow(s);
  if (row0 == NO_ROWS)
    return;
  __shared__ TileSmem sm;
  const u32 i = row0 + threadIdx.x;
  const bool active = i < s.N && !isPassiveRow(s, i, P[i]);
  const float4 pi = active ? P[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  tileFilterToMask<TRAV>(sm, g, c.nbrRadiusSq, s.table, P, pi, active, row0, s.marginMask, s.buildStats);
}

template <int TRAV>
__global__ void __launch_bounds__(NB_THREADS, NB_MIN_BLOCKS) densityLambdaBuildKernel(DeviceState s, GridParams g, SphConsts c, float rho0, float cfm,
    const float4* __restrict__ pred, int epoch)
{
  RTP_PDL_PROLOGUE();
  const u32 row0 = ctaFirstRow(s);
  if (row0 == NO_ROWS)
    return;
  __shared__ MaskWalkSmem sm;
  const u32 i = row0 + threadIdx.x;
  recordGhostBuildPos(s, row0, pred, NBR_BUILD, epoch);
  const bool active = i < s.N && !isPassiveRow(s, i, pred[i]);
  const float4 pi = active ? pred[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  float density = 0.f, gx = 0.f, gy = 0.f, gz = 0.f, sumG2 = 0.f;
  const bool done = sweepProducerFromMask<TRAV>(sm, g, c, s, pred, pi, i, active, epoch,
      [&](u32, float dx, float dy, float dz, float sq)
      {
        const float cs = spikyCoefOrZero(c, sq);
        return PairTerm<6> { { fmul(c.poly6, poly6nc(c, sq)), cs, dx, dy, dz, fmul(fmul(cs, cs), sq) } };
      },
      [&](const PairTerm<6>& t)
      {
        density = fadd(density, t.v[0]);
        gx = ffma(t.v[2], t.v[1], gx);
        gy = ffma(t.v[3], t.v[1], gy);
        gz = ffma(t.v[4], t.v[1], gz);
        sumG2 = fadd(sumG2, t.v[5]);
      });
  if (!done)
    return;
  s.density[i] = density;
  // fluids.cl:189-192
  const float densityC = fsub(fdiv(density, rho0), 1.0f);
  float ssg = fadd(sumG2, dot3c(gx, gy, gz, gx, gy, gz));
  ssg = fdiv(ssg, fmul(rho0, rho0));
  s.lambda[i] = fdiv(-densityC, fadd(ssg, cfm));
}

// ART: artificial pressure compiled in as -1 = off, 4 = exponent 4 (the reference's default), 0 = as the parameters say
template <int TRAV, bool LAST, int ART>
__global__ void __launch_bounds__(NB_THREADS, NB_MIN_BLOCKS) correctionKernel(DeviceState s, GridParams g, SphConsts c, FluidStepParams fp,
    const float4* __restrict__ pred, float4* __restrict__ predOut, int writeCorr, int nbrMode, int epoch)
{
  RTP_PDL_PROLOGUE();
  const u32 row0 = ctaFirstRow(s);
  if (row0 == NO_ROWS)
    return;
  const u32 i = row0 + threadIdx.x;
  if (i >= s.N)
    return;
  const float4 pi = pred[i];
  if (isPassiveRow(s, i, pi))
  {
    // slab decomposition: ghost rows get the owner's value from the caller; "no particle" rows stay what they are
    predOut[i] = pi;
    if (LAST && !fp.f.isVorticityConfEnabled)
    {
      s.posA[i] = make_float4(INFINITY, INFINITY, INFINITY, 0.0f);
      s.velA[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  const float* __restrict__ lambda = s.lambda;
  const float li = lambda[i];
  const bool art = ART == 0 ? fp.f.isArtPressureEnabled != 0 : ART > 0;
  const u32 artExp = ART > 0 ? (u32)ART : fp.f.artPressureExp;
  const float artCoeff = fp.f.artPressureCoeff, invDen = fp.invArtDenom;
  float cx = 0.f, cy = 0.f, cz = 0.f;
  sweepConsumer<TRAV>(g, c, s, pred, pi, i, nbrMode, epoch,
      [&](u32 e, float dx, float dy, float dz, float sq, float cs)
      {
        float sc = fadd(li, __ldg(lambda + e));
        if (art)
        {
          // artPressure fluids.cl:51-57: -k * (W(vec) / W(dq h))^n ; POLY6_COEFF cancels in the ratio; n <= 6
          const float ratio = fmul(poly6nc(c, sq), invDen);
          float pw = ratio;
#pragma unroll
          for (u32 q = 1; q < 6; ++q)
            if (q < artExp)
              pw = fmul(pw, ratio);
          sc = fadd(sc, -fmul(artCoeff, pw));
        }
        const float w = fmul(sc, cs);
        cx = ffma(dx, w, cx);
        cy = ffma(dy, w, cy);
        cz = ffma(dz, w, cz);
      });
  const float rho0 = fp.f.restDensity;
  const float4 corr = make_float4(fdiv(cx, rho0), fdiv(cy, rho0), fdiv(cz, rho0), 0.0f);
  if (writeCorr)
    s.corrPos[i] = corr;
  // fld_correctPosition / cld_correctPosition
  float4 np = make_float4(fadd(pi.x, corr.x), fadd(pi.y, corr.y), fadd(pi.z, corr.z), 0.0f);
  const float idt = fadd(fp.f.timeStep, RTP_FLOAT_EPS);
  if (TRAV == TRAV_FLUIDS)
  {
    if (!LAST)
    {
      np = fluidBoundary(g, np); // next iteration's fld_applyBoundaryCondition (Fluids.cpp:430)
      predOut[i] = np;
    }
    else
    {
      predOut[i] = np;
      // fld_updateVel fluids.cl:263-273
      const float4 p0 = s.posB[i];
      const float4 v = make_float4(fclamp(fdiv(fsub(np.x, p0.x), idt), -c.maxVel, c.maxVel),
          fclamp(fdiv(fsub(np.y, p0.y), idt), -c.maxVel, c.maxVel), fclamp(fdiv(fsub(np.z, p0.z), idt), -c.maxVel, c.maxVel), 0.0f);
      if (fp.f.isVorticityConfEnabled)
      {
        s.velB[i] = v;
      }
      else
      {
        s.velA[i] = v;
        s.posA[i] = np; // fld_updatePosition fluids.cl:444-450
      }
    }
  }
  else
  {
    // clouds: second cld_correctPosition on p_totCorrPos, then the boundary kernel (Clouds.cpp:579-586)
    const float4 t0 = s.totCorrB[i];
    const float4 tot = make_float4(fadd(t0.x, corr.x), fadd(t0.y, corr.y), fadd(t0.z, corr.z), 0.0f);
    s.totCorrB[i] = tot;
    np = cloudBoundary(g, np);
    predOut[i] = np;
    if (LAST)
    {
      // cld_updateVel clouds.cl:958-967
      s.velB[i] = make_float4(fclamp(fdiv(tot.x, idt), -c.maxVel, c.maxVel), fclamp(fdiv(tot.y, idt), -c.maxVel, c.maxVel),
          fclamp(fdiv(tot.z, idt), -c.maxVel, c.maxVel), 0.0f);
    }
  }
  if (nbrMode != NBR_OFF)
  {
    checkListValidity<TRAV>(g, c, s, i, np, epoch + 1);
    // the next producer sweep (epoch + 1) cannot use this particle's margin list: straggler queue (sweep.cuh)
    if (usableMarginList(g, s, cell3D(g, np.x, np.y, np.z), i, NBR_USE, false) == NBR_OVERFLOW)
      pushStraggler(s, epoch + 1, i);
  }
}

template <int TRAV>
__global__ void __launch_bounds__(NB_THREADS, NB_MIN_BLOCKS) vorticityKernel(DeviceState s, GridParams g, SphConsts c, const float4* __restrict__ pred,
    int nbrMode, int epoch)
{
  RTP_PDL_PROLOGUE();
  const u32 row0 = ctaFirstRow(s);
  if (row0 == NO_ROWS)
    return;
  const float4* __restrict__ V = s.velB;
  recordGhostBuildPos(s, row0, pred, nbrMode, epoch);
  producerLoop(s, row0, pred, nbrMode, epoch,
      [&](const u32 i, const bool strag) -> int
      {
        const float4 pi = pred[i];
        const float4 vi = V[i];
        float wx = 0.f, wy = 0.f, wz = 0.f;
        const int r = sweepProducer<TRAV>(g, c, s, pred, pi, i, nbrMode, epoch, strag,
            [&](u32 e, float dx, float dy, float dz, float sq)
            {
              const float cs = spikyCoefOrZero(c, sq);
              const float4 vj = ld4(V, e);
              const float ax = fsub(vj.x, vi.x), ay = fsub(vj.y, vi.y), az = fsub(vj.z, vi.z);
              // cross(dv, vec * c) = cross(dv, vec) * c, cross(a,b).x = fma(a.y, b.z, -(a.z * b.y))
              return PairTerm<4> { { ffma(ay, dz, -fmul(az, dy)), ffma(az, dx, -fmul(ax, dz)), ffma(ax, dy, -fmul(ay, dx)), cs } };
            },
            [&](const PairTerm<4>& t)
            {
              wx = ffma(t.v[0], t.v[3], wx);
              wy = ffma(t.v[1], t.v[3], wy);
              wz = ffma(t.v[2], t.v[3], wz);
            });
        if (r != SWEEP_DONE)
          return r;
        s.vort[i] = make_float4(wx, wy, wz, 0.0f);
        s.vortNorm[i] = fsqrt(dot3c(wx, wy, wz, wx, wy, wz)); // fast_length(vort[e]) of the next sweep
        return (int)SWEEP_DONE;
      });
}

template <int TRAV>
__global__ void __launch_bounds__(NB_THREADS, NB_MIN_BLOCKS) confinementKernel(DeviceState s, GridParams g, SphConsts c, float coeff, float dt,
    const float4* __restrict__ pred, int nbrMode, int epoch)
{
  RTP_PDL_PROLOGUE();
  const u32 row0 = ctaFirstRow(s);
  if (row0 == NO_ROWS)
    return;
  const u32 i = row0 + threadIdx.x;
  if (i >= s.N)
    return;
  const float4 pi = pred[i];
  if (isPassiveRow(s, i, pi))
    return;
  const float* __restrict__ wn = s.vortNorm;
  float nx = 0.f, ny = 0.f, nz = 0.f;
  sweepConsumer<TRAV>(g, c, s, pred, pi, i, nbrMode, epoch,
      [&](u32 e, float dx, float dy, float dz, float, float cs)
      {
        const float w = fmul(__ldg(wn + e), cs);
        nx = ffma(dx, w, nx);
        ny = ffma(dy, w, ny);
        nz = ffma(dz, w, nz);
      });
  // normalize(n) with normalize(0) = 0, then vel += coeff * cross(n, vorticity) * dt   (fluids.cl:376)
  const float l = fsqrt(dot3c(nx, ny, nz, nx, ny, nz));
  if (l == 0.0f)
  {
    nx = ny = nz = 0.0f;
  }
  else
  {
    nx = fdiv(nx, l); ny = fdiv(ny, l); nz = fdiv(nz, l);
  }
  const float4 w = s.vort[i];
  const float crx = fsub(fmul(ny, w.z), fmul(nz, w.y)), cry = fsub(fmul(nz, w.x), fmul(nx, w.z)), crz = fsub(fmul(nx, w.y), fmul(ny, w.x));
  const float4 v = s.velB[i];
  s.velC[i] = make_float4(fadd(v.x, fmul(fmul(crx, coeff), dt)), fadd(v.y, fmul(fmul(cry, coeff), dt)), fadd(v.z, fmul(fmul(crz, coeff), dt)), 0.0f);
}

template <int TRAV>
__global__ void __launch_bounds__(NB_THREADS, NB_MIN_BLOCKS) xsphKernel(DeviceState s, GridParams g, SphConsts c, float coeff, const float4* __restrict__ pred,
    int nbrMode, int epoch)
{
  RTP_PDL_PROLOGUE();
  const u32 row0 = ctaFirstRow(s);
  if (row0 == NO_ROWS)
    return;
  const u32 i = row0 + threadIdx.x;
  if (i >= s.N)
    return;
  const float4 pi = pred[i];
  if (isPassiveRow(s, i, pi))
  {
    // slab decomposition: the step leaves ghost and "no particle" rows marked (the caller compacts the particles to the front)
    s.posA[i] = make_float4(INFINITY, INFINITY, INFINITY, 0.0f);
    s.velA[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float4* __restrict__ V = s.velC;
  const float4 vi = V[i];
  float sx = 0.f, sy = 0.f, sz = 0.f;
  sweepConsumer<TRAV>(g, c, s, pred, pi, i, nbrMode, epoch,
      [&](u32 e, float, float, float, float sq, float)
      {
        const float W = fmul(c.poly6, poly6nc(c, sq));
        const float4 vj = ld4(V, e);
        sx = ffma(fsub(vj.x, vi.x), W, sx);
        sy = ffma(fsub(vj.y, vi.y), W, sy);
        sz = ffma(fsub(vj.z, vi.z), W, sz);
      });
  s.velA[i] = make_float4(fadd(vi.x, fmul(sx, coeff)), fadd(vi.y, fmul(sy, coeff)), fadd(vi.z, fmul(sz, coeff)), 0.0f);
  if (TRAV == TRAV_FLUIDS)
    s.posA[i] = pi; // fld_updatePosition fluids.cl:444-450 (clouds: cloudsFinishKernel)
}

// cld_computeLaplacianTemp clouds.cl:508-569 -- on the SORTED p_pos with the table built from p_predPos (Clouds.cpp:253)
__global__ void __launch_bounds__(NB_THREADS, NB_MIN_BLOCKS) laplacianTempKernel(DeviceState s, GridParams g, SphConsts c, float rho0, int nbrMode)
{
  RTP_PDL_PROLOGUE();
  const float* __restrict__ T = s.tempB;
  producerLoop(s, blockIdx.x * NB_THREADS, s.posB, nbrMode, NBR_EPOCH_TEMP,
      [&](const u32 i, const bool strag) -> int
      {
        const float4 pi = s.posB[i];
        const float Ti = T[i];
        float lap = 0.f;
        const int r = sweepProducer<TRAV_CLOUDS>(g, c, s, s.posB, pi, i, nbrMode, NBR_EPOCH_TEMP, strag,
            [&](u32 e, float, float, float, float sq)
            {
              const float cs = spikyCoefOrZero(c, sq);
              // dot(vec, grad) = c * sq ; x / d = x * (1/d)
              return PairTerm<2> { { fmul(fsub(Ti, __ldg(T + e)), fmul(cs, sq)), rcpInRange(fadd(sq, RTP_FLOAT_EPS)) } };
            },
            [&](const PairTerm<2>& t) { lap = ffma(t.v[0], t.v[1], lap); });
        if (r != SWEEP_DONE)
          return r;
        s.lapTemp[i] = fdiv(lap, rho0);
        return (int)SWEEP_DONE;
      });
}

// the same sweep as the list build of the temperature epoch: walks the margin mask marginMaskKernel wrote (tilebuild.cuh)
__global__ void __launch_bounds__(NB_THREADS, NB_MIN_BLOCKS) laplacianTempBuildKernel(DeviceState s, GridParams g, SphConsts c, float rho0)
{
  RTP_PDL_PROLOGUE();
  __shared__ MaskWalkSmem sm;
  const float* __restrict__ T = s.tempB;
  const u32 i = blockIdx.x * NB_THREADS + threadIdx.x;
  const bool active = i < s.N;
  const float4 pi = active ? s.posB[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float Ti = active ? T[i] : 0.f;
  float lap = 0.f;
  const bool done = sweepProducerFromMask<TRAV_CLOUDS>(sm, g, c, s, s.posB, pi, i, active, NBR_EPOCH_TEMP,
      [&](u32 e, float, float, float, float sq)
      {
        const float cs = spikyCoefOrZero(c, sq);
        return PairTerm<2> { { fmul(fsub(Ti, __ldg(T + e)), fmul(cs, sq)), rcpInRange(fadd(sq, RTP_FLOAT_EPS)) } };
      },
      [&](const PairTerm<2>& t) { lap = ffma(t.v[0], t.v[1], lap); });
  if (done)
    s.lapTemp[i] = fdiv(lap, rho0);
}

// cld_computeConstraintFactorTemp clouds.cl:575-648
__global__ void __launch_bounds__(NB_THREADS, NB_MIN_BLOCKS) lambdaTempKernel(DeviceState s, GridParams g, SphConsts c, float rho0, float cfm, int nbrMode)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * NB_THREADS + threadIdx.x;
  if (i >= s.N)
    return;
  const float4 pi = s.posB[i];
  float sumD = 0.f, sumD2 = 0.f;
  sweepConsumer<TRAV_CLOUDS>(g, c, s, s.posB, pi, i, nbrMode, NBR_EPOCH_TEMP,
      [&](u32, float, float, float, float sq, float cs)
      {
        const float d = fmul(fmul(cs, sq), rcpInRange(ffma(sq, rho0, RTP_FLOAT_EPS)));
        sumD = fadd(sumD, d);
        sumD2 = ffma(d, d, sumD2);
      });
  const float ssg = fadd(sumD2, fmul(sumD, sumD));
  s.lambdaTemp[i] = fdiv(-s.lapTemp[i], fadd(ssg, cfm));
}

// cld_computeConstraintCorrectionTemp clouds.cl:654-722 + cld_correctTemperature :931-937
__global__ void __launch_bounds__(NB_THREADS, NB_MIN_BLOCKS) correctTempKernel(DeviceState s, GridParams g, SphConsts c, float rho0, int nbrMode)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * NB_THREADS + threadIdx.x;
  if (i >= s.N)
    return;
  const float4 pi = s.posB[i];
  const float* __restrict__ L = s.lambdaTemp;
  const float li = L[i];
  float corr = 0.f;
  sweepConsumer<TRAV_CLOUDS>(g, c, s, s.posB, pi, i, nbrMode, NBR_EPOCH_TEMP,
      [&](u32 e, float, float, float, float sq, float cs)
      {
        const float d = fmul(fmul(cs, sq), rcpInRange(ffma(sq, rho0, RTP_FLOAT_EPS)));
        corr = ffma(fadd(li, __ldg(L + e)), d, corr);
      });
  s.corrTemp[i] = corr;
  s.tempB[i] = fadd(s.tempB[i], fmul(0.3f, corr));
}

// ------------------------------------------------------------------ launch wrappers

static inline int ewBlocks(size_t n) { return (int)((n + EW_THREADS - 1) / EW_THREADS); }
static inline int nbBlocks(size_t n) { return (int)((n + NB_THREADS - 1) / NB_THREADS); }
// grid of a PBF neighbour sweep: every row, or the caller's bound for a launch by row phase (kernels.cuh: rowPhase)
static inline int sweepBlocks(const DeviceState& s) { return s.rowPhase ? (int)s.rowPhaseBlocks : nbBlocks(s.N); }

static_assert(EW_THREADS == SORT_THREADS, "fluidPredictKernel<true> builds the sort histograms with SORT_THREADS threads per block");
void launchFluidPredict(const DeviceState& s, const GridParams& g, const FluidStepParams& p, u32* keysOut, const SortPlan* fusedSort,
    u32* sortCtrl, u32* sortStatus, cudaStream_t st, bool resets)
{
  // resets = false: only the prediction and the cell ids of the rows (slab decomposition: the arrivals of a migration)
  const int blocks = ewBlocks(resets ? max(s.N, g.numCells) : s.N);
  if (!blocks)
    return;
  if (fusedSort && fusedSort->n)
    launchKernel(fluidPredictKernel<true>, blocks, EW_THREADS, st, s, g, p.f.timeStep, keysOut, fusedSort->passes, makePassDesc(*fusedSort),
        sortCtrl, sortStatus, sortStatusWords(*fusedSort), resets ? 1 : 0);
  else
    launchKernel(fluidPredictKernel<false>, blocks, EW_THREADS, st, s, g, p.f.timeStep, keysOut, 0, PassDesc {}, (u32*)nullptr, (u32*)nullptr,
        (size_t)0, resets ? 1 : 0);
}
void launchMarginMask(const DeviceState& s, int model, const GridParams& g, const SphConsts& c, const float4* P, cudaStream_t st)
{
  if (!s.N)
    return;
  if (model == RTP_MODEL_CLOUDS)
    launchKernel(marginMaskKernel<TRAV_CLOUDS>, sweepBlocks(s), NB_THREADS, st, s, g, c, P);
  else
    launchKernel(marginMaskKernel<TRAV_FLUIDS>, sweepBlocks(s), NB_THREADS, st, s, g, c, P);
}
void launchFluidGather(const DeviceState& s, const GridParams& g, cudaStream_t st)
{
  if (s.N)
    launchKernel(fluidGatherKernel, ewBlocks(s.N), EW_THREADS, st, s, g);
}
int launchDensityLambda(const DeviceState& s, int model, const GridParams& g, const SphConsts& c, const FluidStepParams& p,
    const float4* pred, int nbrMode, int epoch, cudaStream_t st)
{
  if (!s.N)
    return 0;
  static_assert(NB_THREADS == TB_THREADS && NB_THREADS == SWEEP_BLOCK_ROWS, "tilebuild.cuh is written for the neighbour kernels' block size");
  if (nbrMode == NBR_BUILD && s.tiledBuild)
  {
    launchMarginMask(s, model, g, c, pred, st);
    if (model == RTP_MODEL_CLOUDS)
      launchKernel(densityLambdaBuildKernel<TRAV_CLOUDS>, sweepBlocks(s), NB_THREADS, st, s, g, c, p.f.restDensity, p.f.relaxCFM, pred, epoch);
    else
      launchKernel(densityLambdaBuildKernel<TRAV_FLUIDS>, sweepBlocks(s), NB_THREADS, st, s, g, c, p.f.restDensity, p.f.relaxCFM, pred, epoch);
    return 2; // filter + build walk
  }
  if (model == RTP_MODEL_CLOUDS)
    launchKernel(densityLambdaKernel<TRAV_CLOUDS>, sweepBlocks(s), NB_THREADS, st, s, g, c, p.f.restDensity, p.f.relaxCFM, pred, nbrMode, epoch);
  else
    launchKernel(densityLambdaKernel<TRAV_FLUIDS>, sweepBlocks(s), NB_THREADS, st, s, g, c, p.f.restDensity, p.f.relaxCFM, pred, nbrMode, epoch);
  return 1;
}
template <int TRAV, bool LAST>
static void launchCorrectionArt(const DeviceState& s, const GridParams& g, const SphConsts& c, const FluidStepParams& p, const float4* pred,
    float4* predOut, int wc, int nbrMode, int epoch, cudaStream_t st)
{
  const int nb = sweepBlocks(s);
  if (!p.f.isArtPressureEnabled)
    launchKernel(correctionKernel<TRAV, LAST, -1>, nb, NB_THREADS, st, s, g, c, p, pred, predOut, wc, nbrMode, epoch);
  else if (p.f.artPressureExp == 4u)
    launchKernel(correctionKernel<TRAV, LAST, 4>, nb, NB_THREADS, st, s, g, c, p, pred, predOut, wc, nbrMode, epoch);
  else
    launchKernel(correctionKernel<TRAV, LAST, 0>, nb, NB_THREADS, st, s, g, c, p, pred, predOut, wc, nbrMode, epoch);
}
void launchCorrection(const DeviceState& s, int model, const GridParams& g, const SphConsts& c, const FluidStepParams& p,
    const rtp_cloud_params&, const float4* pred, float4* predOut, bool last, bool writeCorr, int nbrMode, int epoch, cudaStream_t st)
{
  if (!s.N)
    return;
  const int wc = writeCorr ? 1 : 0;
  if (model == RTP_MODEL_CLOUDS)
  {
    if (last)
      launchCorrectionArt<TRAV_CLOUDS, true>(s, g, c, p, pred, predOut, wc, nbrMode, epoch, st);
    else
      launchCorrectionArt<TRAV_CLOUDS, false>(s, g, c, p, pred, predOut, wc, nbrMode, epoch, st);
  }
  else
  {
    if (last)
      launchCorrectionArt<TRAV_FLUIDS, true>(s, g, c, p, pred, predOut, wc, nbrMode, epoch, st);
    else
      launchCorrectionArt<TRAV_FLUIDS, false>(s, g, c, p, pred, predOut, wc, nbrMode, epoch, st);
  }
}
void launchVorticity(const DeviceState& s, int model, const GridParams& g, const SphConsts& c, const float4* pred, int nbrMode,
    int epoch, cudaStream_t st)
{
  if (!s.N)
    return;
  if (model == RTP_MODEL_CLOUDS)
    launchKernel(vorticityKernel<TRAV_CLOUDS>, sweepBlocks(s), NB_THREADS, st, s, g, c, pred, nbrMode, epoch);
  else
    launchKernel(vorticityKernel<TRAV_FLUIDS>, sweepBlocks(s), NB_THREADS, st, s, g, c, pred, nbrMode, epoch);
}
void launchConfinement(const DeviceState& s, int model, const GridParams& g, const SphConsts& c, const FluidStepParams& p,
    const float4* pred, int nbrMode, int epoch, cudaStream_t st)
{
  if (!s.N)
    return;
  if (model == RTP_MODEL_CLOUDS)
    launchKernel(confinementKernel<TRAV_CLOUDS>, sweepBlocks(s), NB_THREADS, st, s, g, c, p.f.vorticityConfCoeff, p.f.timeStep, pred, nbrMode, epoch);
  else
    launchKernel(confinementKernel<TRAV_FLUIDS>, sweepBlocks(s), NB_THREADS, st, s, g, c, p.f.vorticityConfCoeff, p.f.timeStep, pred, nbrMode, epoch);
}
void launchXsph(const DeviceState& s, int model, const GridParams& g, const SphConsts& c, const FluidStepParams& p,
    const rtp_cloud_params&, const float4* pred, int nbrMode, int epoch, cudaStream_t st)
{
  if (!s.N)
    return;
  if (model == RTP_MODEL_CLOUDS)
    launchKernel(xsphKernel<TRAV_CLOUDS>, sweepBlocks(s), NB_THREADS, st, s, g, c, p.f.xsphViscosityCoeff, pred, nbrMode, epoch);
  else
    launchKernel(xsphKernel<TRAV_FLUIDS>, sweepBlocks(s), NB_THREADS, st, s, g, c, p.f.xsphViscosityCoeff, pred, nbrMode, epoch);
}
void launchCloudsInitFields(const DeviceState& s, const GridParams& g, const rtp_cloud_params& cloud, cudaStream_t st)
{
  launchKernel(cloudsInitFieldsKernel, ewBlocks(s.M), EW_THREADS, st, s, g, cloud.initVaporDensityCoeff);
}
void launchCloudsThermoPredict(const DeviceState& s, const GridParams& g, const rtp_cloud_params& cloud, u32* keysOut, cudaStream_t st)
{
  launchKernel(cloudsThermoPredictKernel, ewBlocks(max(s.N, g.numCells)), EW_THREADS, st, s, g, cloud, keysOut);
}
void launchCloudsGather(const DeviceState& s, const GridParams& g, cudaStream_t st)
{
  if (s.N)
    launchKernel(cloudsGatherKernel, ewBlocks(s.N), EW_THREADS, st, s, g);
}
int launchCloudsLaplacianTemp(const DeviceState& s, const GridParams& g, const SphConsts& c, const rtp_cloud_params& cloud, int nbrMode, cudaStream_t st)
{
  if (s.N && nbrMode == NBR_BUILD && s.tiledBuild)
  {
    launchMarginMask(s, RTP_MODEL_CLOUDS, g, c, s.posB, st);
    launchKernel(laplacianTempBuildKernel, nbBlocks(s.N), NB_THREADS, st, s, g, c, cloud.restDensity);
    return 2; // filter + build walk
  }
  if (s.N)
    launchKernel(laplacianTempKernel, nbBlocks(s.N), NB_THREADS, st, s, g, c, cloud.restDensity, nbrMode);
  return s.N ? 1 : 0;
}
void launchCloudsLambdaTemp(const DeviceState& s, const GridParams& g, const SphConsts& c, const rtp_cloud_params& cloud, int nbrMode, cudaStream_t st)
{
  if (s.N)
    launchKernel(lambdaTempKernel, nbBlocks(s.N), NB_THREADS, st, s, g, c, cloud.restDensity, cloud.relaxCFM, nbrMode);
}
void launchCloudsCorrectTemp(const DeviceState& s, const GridParams& g, const SphConsts& c, const rtp_cloud_params& cloud, int nbrMode, cudaStream_t st)
{
  if (s.N)
    launchKernel(correctTempKernel, nbBlocks(s.N), NB_THREADS, st, s, g, c, cloud.restDensity, nbrMode);
}
void launchCloudsFinish(const DeviceState& s, const GridParams& g, const rtp_cloud_params& cloud, const float4* pred, bool copyVel, cudaStream_t st)
{
  if (s.N)
    launchKernel(cloudsFinishKernel, ewBlocks(s.N), EW_THREADS, st, s, g, cloud, pred, copyVel ? 1 : 0);
}

// ---- diagnostics: how the margin / hit lists of the last step would serve a sweep at the final predicted positions
__global__ void __launch_bounds__(EW_THREADS) listStatsKernel(DeviceState s, GridParams g, SphConsts c, const float4* __restrict__ pred,
    unsigned long long* __restrict__ out)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s.N)
    return;
  const float4 pi = pred[i];
  const int3 ci = cell3D(g, pi.x, pi.y, pi.z);
  const u32 raw = s.nbrCount[i];
  const float4 bp = s.nbrBuildPos[i];
  const int3 cb = cell3D(g, bp.x, bp.y, bp.z);
  atomicAdd(out + 0, 1ull);
  if (raw == NBR_OVERFLOW)
    atomicAdd(out + 1, 1ull);
  else
  {
    atomicAdd(out + 4, (unsigned long long)raw);
    atomicMax(out + 5, (unsigned long long)raw);
  }
  if (cb.x != ci.x || cb.y != ci.y || cb.z != ci.z)
  {
    atomicAdd(out + 2, 1ull);
  }
  if (s.hitCount[i] == NBR_OVERFLOW)
    atomicAdd(out + 6, 1ull);
  const float dx = pi.x - bp.x, dy = pi.y - bp.y, dz = pi.z - bp.z;
  if (!(dx * dx + dy * dy + dz * dz <= c.nbrDmaxSq))
    atomicAdd(out + 7, 1ull);
}
void launchListStats(const DeviceState& s, const GridParams& g, const SphConsts& c, const float4* pred, unsigned long long* out, cudaStream_t st)
{
  if (s.N)
    launchKernel(listStatsKernel, ewBlocks(s.N), EW_THREADS, st, s, g, c, pred, out);
}

} // namespace rtp
