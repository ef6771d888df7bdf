// boids.cu -- Reynolds boids on the uniform grid (reference: physics/ocl/kernels/boids.cl, host sequence
// Boids::update physics/ocl/Boids.cpp:323-384).
//
// Step = 3 model launches around the sort:
//   boidsCellIdsKernel : fillCellIDs on p_pos (grid.cl:76-86) + resetStartEndCell (grid.cl:91-96)
//   boidsGatherKernel  : payload permutation of p_pos/p_vel (radixSort.cl:179-190) into the sorted working copies,
//                        pre-normalised neighbour velocities, fillStartCell/fillEndCell (grid.cl:101-138)
//   boidsRulesKernel   : bd_applyBoidsRulesWithGrid3D/2D (boids.cl:46-221) fused with bd_addTargetRule (:226-241),
//                        bd_updateVel (:246-259) and bd_updatePosAndApply{Wall,Periodic}BC (:264-315)
// p_acc and p_col are not permuted by the cell sort: p_acc is fully rewritten every step and p_col is uniform.
// The whole rule evaluation is bit-exact with the oracle: exact hit test, sums in the reference's order, the
// canonical operation sequence of DESIGN.md "Canonical arithmetic".
#include "kernels.cuh"

namespace rtp
{
#ifndef RTP_BD_THREADS
#define RTP_BD_THREADS 128
#endif
constexpr int BD_THREADS = RTP_BD_THREADS;
#ifndef RTP_BD_GROUP
#define RTP_BD_GROUP 8
#endif
constexpr int BD_GROUP = RTP_BD_GROUP; // candidates in flight per thread in the rules sweep

// fast_normalize(v) = v * (1 / sqrt(dot(v, v))), and a zero vector is returned unchanged (OpenCL 1.2 s6.12.5; the test is
// dot == 0, as in the oracle)
__device__ __forceinline__ float4 fastNormalize(const float4 v)
{
  const float d = dot3c(v.x, v.y, v.z, v.x, v.y, v.z);
  if (d == 0.0f)
    return v;
  const float r = fdiv(1.0f, fsqrt(d));
  return make_float4(fmul(v.x, r), fmul(v.y, r), fmul(v.z, r), fmul(v.w, r));
}

__device__ __forceinline__ void fillCellTable(const u32* __restrict__ keys, u32 i, u32 N, u32 numCells, uint2* __restrict__ table)
{
  const u32 id = keys[i];
  if (id < numCells)
  {
    // fillStartCell grid.cl:101-117: sorted index 0 never writes its start (reset value 1 stays)
    if (i > 0 && id != keys[i - 1])
      table[id].x = i;
    // fillEndCell grid.cl:122-138: cellID[N] is a tail key >= 2C (or out of bounds when N == M): never equal
    const u32 next = (i + 1 < N) ? keys[i + 1] : 0xFFFFFFFFu;
    if (id != next)
      table[id].y = i;
  }
}

__global__ void __launch_bounds__(256) boidsCellIdsKernel(const float4* __restrict__ pos, GridParams g, u32* __restrict__ keys,
    uint2* __restrict__ table, u32 N)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * 256 + threadIdx.x;
  if (i < g.numCells)
    table[i] = make_uint2(1u, 0u);
  if (i < N)
  {
    const float4 p = pos[i];
    keys[i] = cell1D(g, p.x, p.y, p.z);
  }
}

__global__ void __launch_bounds__(256) boidsGatherKernel(DeviceState s, GridParams g)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * 256 + threadIdx.x;
  if (i >= s.N)
    return;
  const u32 j = s.perm[i];
  s.posB[i] = s.posA[j];
  const float4 v = s.velA[j];
  s.velB[i] = v;
  // fast_normalize(velocity[e]) (boids.cl:104) depends on e only: do it once per particle, exactly
  s.velC[i] = fastNormalize(v);
  fillCellTable(s.cellID, i, s.N, g.numCells, s.table);
}

template <bool DIM2>
__global__ void __launch_bounds__(BD_THREADS) boidsRulesKernel(DeviceState s, GridParams g, SphConsts c, BoidsStepParams p)
{
  RTP_PDL_PROLOGUE();
  const u32 i = blockIdx.x * BD_THREADS + threadIdx.x;
  if (i >= s.N)
    return;
  const float4* __restrict__ P = s.posB;
  const float4* __restrict__ NV = s.velC;
  const float4 pi = P[i];
  const int3 ci = cell3D(g, pi.x, pi.y, pi.z);

  int count = 0;
  float apx = 0.f, apy = 0.f, apz = 0.f; // averageBoidsPos
  float avx = 0.f, avy = 0.f, avz = 0.f; // averageBoidsVel
  float rpx = 0.f, rpy = 0.f, rpz = 0.f; // repulseHeading

  // One candidate: the reference's test and its nine ordered sums (boids.cl:96-110). nv / r are fetched and computed
  // BEFORE the test by the caller: no load and no reciprocal sits behind the (divergent) hit branch.
  auto accumulate = [&](const float4 pj, const float4 nv, float dx, float dy, float dz, float r)
  {
    apx = fadd(apx, pj.x); apy = fadd(apy, pj.y); apz = fadd(apz, pj.z);
    avx = fadd(avx, nv.x); avy = fadd(avy, nv.y); avz = fadd(avz, nv.z);
    rpx = ffma(dx, r, rpx); rpy = ffma(dy, r, rpy); rpz = ffma(dz, r, rpz); // vec / squaredDist == vec * (1 / squaredDist), one fused op per component
    ++count;
  };
  // The flock collapses into a few cells holding thousands of boids each: the step is the ~10 warps per SM that sweep them,
  // far too few to hide a load behind every hit. Candidates are therefore taken BD_GROUP at a time: positions AND pre-normalised
  // velocities of all four are in flight together, the four tests and reciprocals are independent, and only the short,
  // load-free sums run in order.
  auto visit = [&](u32 start, u32 end, float, float)
  {
    u32 e = start;
#pragma unroll 1
    for (; e + (u32)(BD_GROUP - 1) <= end; e += (u32)BD_GROUP)
    {
      float4 a[BD_GROUP], v[BD_GROUP];
#pragma unroll
      for (int k = 0; k < BD_GROUP; ++k)
      {
        a[k] = __ldg(P + e + k);
        v[k] = __ldg(NV + e + k);
      }
      float dx[BD_GROUP], dy[BD_GROUP], dz[BD_GROUP], r[BD_GROUP];
      bool hit[BD_GROUP];
#pragma unroll
      for (int k = 0; k < BD_GROUP; ++k)
      {
        dx[k] = pi.x - a[k].x, dy[k] = pi.y - a[k].y, dz[k] = pi.z - a[k].z;
        const float sq = dot3c(dx[k], dy[k], dz[k], dx[k], dy[k], dz[k]);
        hit[k] = sq < c.effectRadiusSq && sq > RTP_FLOAT_EPS;
        r[k] = frcp(hit[k] ? sq : 1.0f);
      }
#pragma unroll
      for (int k = 0; k < BD_GROUP; ++k)
        if (hit[k])
          accumulate(a[k], v[k], dx[k], dy[k], dz[k], r[k]);
    }
#pragma unroll 1
    for (; e <= end; ++e)
    {
      const float4 pj = __ldg(P + e), nv = __ldg(NV + e);
      const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
      const float sq = dot3c(dx, dy, dz, dx, dy, dz);
      if (sq < c.effectRadiusSq && sq > RTP_FLOAT_EPS)
        accumulate(pj, nv, dx, dy, dz, frcp(sq));
    }
  };

  if (!DIM2)
  {
    forEachNeighbourRun<TRAV_BOIDS>(g, s.table, ci, visit);
  }
  else
  {
    // boids.cl:170-183: 9 YZ cells, every axis tested against GRID_RES_X, x cell forced to GRID_RES_X / 2
    const int RX = g.res[0], RY = g.res[1];
    for (int iY = -1; iY <= 1; ++iY)
      for (int iZ = -1; iZ <= 1; ++iZ)
      {
        const int cx = ci.x, cy = ci.y + iY, cz = ci.z + iZ;
        if (cx < 0 || cy < 0 || cz < 0 || cx >= RX || cy >= RX || cz >= RX)
          continue;
        const uint2 se = __ldg(&s.table[(RX / 2 * RX + cy) * RY + cz]);
        if (se.x <= se.y)
          visit(se.x, se.y, 0.f, 0.f);
      }
  }

  // The w lane is carried like the reference's float4 arithmetic does (it stays 0 on this path).
  float ax = 0.f, ay = 0.f, az = 0.f, aw = 0.f;
  if (count != 0)
  {
    // boids.cl:115-131
    const float fc = (float)count;
    const float vs = p.rules.velocityScale;
    const float4 ap = fastNormalize(make_float4(fsub(fdiv(apx, fc), pi.x), fsub(fdiv(apy, fc), pi.y), fsub(fdiv(apz, fc), pi.z), 0.0f));
    const float4 av = fastNormalize(make_float4(avx, avy, avz, 0.0f));
    const float4 rp = fastNormalize(make_float4(rpx, rpy, rpz, 0.0f));
    const float al = p.rules.alignmentScale, se = p.rules.separationScale, co = p.rules.cohesionScale;
    ax = fadd(fadd(fmul(fmul(av.x, vs), al), fmul(fmul(rp.x, vs), se)), fmul(fmul(ap.x, vs), co));
    ay = fadd(fadd(fmul(fmul(av.y, vs), al), fmul(fmul(rp.y, vs), se)), fmul(fmul(ap.y, vs), co));
    az = fadd(fadd(fmul(fmul(av.z, vs), al), fmul(fmul(rp.z, vs), se)), fmul(fmul(ap.z, vs), co));
    aw = fadd(fadd(fmul(fmul(av.w, vs), al), fmul(fmul(rp.w, vs), se)), fmul(fmul(ap.w, vs), co));
  }

  // bd_addTargetRule boids.cl:226-241
  if (p.targetActive)
  {
    const float tx = fsub(p.targetPos[0], pi.x), ty = fsub(p.targetPos[1], pi.y), tz = fsub(p.targetPos[2], pi.z);
    const float dist = fsqrt(dot3c(tx, ty, tz, tx, ty, tz));
    if (dist < p.target.targetRadiusEffect)
    {
      const float sgn = (float)p.target.targetSignEffect;
      const float k = fclamp(fdiv(1.3f, dist), 0.0f, 1.4f * RTP_MAX_STEERING);
      ax = fadd(ax, fmul(fmul(tx, sgn), k));
      ay = fadd(ay, fmul(fmul(ty, sgn), k));
      az = fadd(az, fmul(fmul(tz, sgn), k));
      aw = fadd(aw, fmul(fmul(fsub(p.targetPos[3], pi.w), sgn), k));
    }
  }
  s.acc[i] = make_float4(ax, ay, az, aw);

  // bd_updateVel boids.cl:246-259
  const float4 vi = s.velB[i];
  const float maxV = p.rules.velocityScale;
  const float nvx = fadd(vi.x, fmul(ax, p.dt)), nvy = fadd(vi.y, fmul(ay, p.dt)), nvz = fadd(vi.z, fmul(az, p.dt)), nvw = fadd(vi.w, fmul(aw, p.dt));
  const float len = fsqrt(dot3c(nvx, nvy, nvz, nvx, nvy, nvz));
  const float norm = fclamp(len, fmul(0.2f, maxV), maxV);
  const float4 nn = fastNormalize(make_float4(nvx, nvy, nvz, nvw));
  float vx = fmul(nn.x, norm), vy = fmul(nn.y, norm), vz = fmul(nn.z, norm), vw = fmul(nn.w, norm);

  // bd_updatePosAndApplyWallBC boids.cl:264-284 / bd_updatePosAndApplyPeriodicBC :289-315
  const float npx = fadd(pi.x, fmul(vx, p.dt)), npy = fadd(pi.y, fmul(vy, p.dt)), npz = fadd(pi.z, fmul(vz, p.dt));
  float cpx = fclamp(npx, -g.absW[0], g.absW[0]), cpy = fclamp(npy, -g.absW[1], g.absW[1]), cpz = fclamp(npz, -g.absW[2], g.absW[2]);
  if (p.boundary == RTP_BOUNDARY_CYCLIC_WALL)
  {
    if (!(cpx == npx)) cpx = fmul(cpx, -1.0f);
    if (!(cpy == npy)) cpy = fmul(cpy, -1.0f);
    if (!(cpz == npz)) cpz = fmul(cpz, -1.0f);
  }
  else if (!(cpx == npx && cpy == npy && cpz == npz))
  {
    vx = fmul(vx, -0.5f); vy = fmul(vy, -0.5f); vz = fmul(vz, -0.5f); vw = fmul(vw, -0.5f);
  }
  s.posA[i] = make_float4(cpx, cpy, cpz, 0.f); // w: clamp(w, 0, 0) == 0 even for NaN
  s.velA[i] = make_float4(vx, vy, vz, vw);
}

void launchBoidsCellIds(const DeviceState& s, const GridParams& g, u32* keysOut, cudaStream_t st)
{
  const u32 n = max(s.N, g.numCells);
  launchKernel(boidsCellIdsKernel, (n + 255) / 256, 256, st, s.posA, g, keysOut, s.table, s.N);
}
void launchBoidsGather(const DeviceState& s, const GridParams& g, cudaStream_t st)
{
  if (s.N)
    launchKernel(boidsGatherKernel, (s.N + 255) / 256, 256, st, s, g);
}
void launchBoidsRules(const DeviceState& s, const GridParams& g, const SphConsts& c, const BoidsStepParams& p, cudaStream_t st)
{
  if (!s.N)
    return;
  const int blocks = (s.N + BD_THREADS - 1) / BD_THREADS;
  if (p.dim == 2)
    launchKernel(boidsRulesKernel<true>, blocks, BD_THREADS, st, s, g, c, p);
  else
    launchKernel(boidsRulesKernel<false>, blocks, BD_THREADS, st, s, g, c, p);
}

} // namespace rtp
