// tilebuild.cuh -- block-cooperative, shared-memory-staged build of the per-step neighbour lists (sweep.cuh).
//
// The first sweep of a step visits, per particle, the 27 cells around the particle's current cell (fluids.cl:101-123):
// ~460 candidates of which ~110 are within the list radius (1 + margin) h and ~70 inside the support. One thread walking
// its own ragged runs with a gather and two divergent appends per candidate needed ~40 thread instructions per candidate
// at 20 of 32 lanes (round 1: 185 us of a 650 us step). Here a CTA of 128 consecutive cell-sorted particles works as a team:
//
//  WINDOWS. For a particle, the three z cells of one (iX, iY) neighbour column are one contiguous index run of the
//   cell-sorted position array (ids are z-fastest); a z wrap splits it into "pre | main | post" runs, so a particle has
//   27 SLOTS = 9 columns x {pre, main, post}, visited in the reference's order (pre/post are almost always empty). The
//   lanes of a warp that sit in the same (x, y) column have overlapping runs: per slot and lane GROUP (lanes of one
//   column, found by leader election) the warp WINDOW is the union [min start, max end].
//  STAGING. Per neighbour column the windows of the CTA's warps are merged into disjoint index segments; consecutive
//   columns are packed into STAGES that fit the tile (usually three stages of three columns), and a stage is copied into
//   shared memory with 1-D TMA bulk copies (cp.async.bulk, one per segment, completion on an mbarrier). The resident
//   CTAs of an SM are in different phases, so the copies of one run under the filter of the others.
//  FILTER. Every lane tests ALL candidates of its warp's window in lock step: the candidate is one broadcast LDS.128 for
//   the whole warp, the test three FFMAs and an FSETP: a stage's positions are rewritten in place, right after the copy, as
//   (c - origin, |c - origin|^2) with origin = the CTA's first particle, and |c'|^2 - 2 p'.c' is compared with a per-lane
//   threshold r^2 - |p'|^2 raised by four times the rounding bound (a superset of the exact test; the walk that follows
//   applies the exact support test anyway). The result is one bit of a 32-candidate mask word. No divergence, no per-lane addresses. Candidates of the window outside the lane's
//   own run (another lane's cell) are removed per word with a range mask, so a lane keeps exactly the reference's
//   candidate set; words no lane needs (a hole between two families of runs) are skipped.
//  MASKS. The words go straight to global memory (one coalesced 128-byte row per word and warp) as the MARGIN MASK of the
//   step: per warp up to wordCap words per lane plus, per word, the global index of its bit 0 and its slot. A second,
//   per-thread kernel (sweep.cuh sweepProducerFromMask) walks the set bits of a particle's words in order (= the
//   reference's order: slots ascending, index ascending), appends the candidates to the margin list, applies the exact
//   support test and runs the sweep's pair term. (One fused kernel had to walk after every stage, while the stage's tile
//   was still in shared memory: three short walks per particle, each paced by its slowest lane -- 16 of 32 lanes busy.
//   The split walk covers a whole particle at once and needs no tile.)
//  FALLBACK. A warp whose lanes cannot be described this way (a cap or start = 1 quirk breaks a run, more than NGROUP
//   columns in one warp, windows larger than the tile) builds its lists with the per-thread 27-cell traversal. Both
//   paths emit the same candidates in the same order.
#pragma once

#include "kernels.cuh"

namespace rtp
{
constexpr int TB_THREADS = 128; // == the neighbour kernels' block size
constexpr int TB_WARPS = TB_THREADS / 32;
constexpr u32 TILE_STAGE_CAP = 1088; // positions one stage may hold: 128 + 2 x 101 (cell cap) per column, x 3
constexpr u32 TILE_PAD = 32; // the filter reads whole 32-candidate words: over-read room behind the last window
constexpr int NSLOT = 27;
constexpr int NGROUP = 4; // lane groups of a warp (lanes with the same centre column)
constexpr int SEG_MAX = TB_WARPS * NGROUP + 4; // staged index segments of one column
constexpr int WIN_MAX = 6; // non-empty windows of one warp in one column (usually 1)
constexpr u32 TB_INDEX_BITS = 27; // word descriptor = global index of bit 0 | slot << 27
constexpr u32 TB_INDEX_MASK = (1u << TB_INDEX_BITS) - 1u;

struct __align__(16) TileSmem
{
  float4 tile[TILE_STAGE_CAP + TILE_PAD];
  // per warp and neighbour column: its non-empty windows in slot order: (first index, last index, slot | group << 8, -)
  uint4 winList[TB_WARPS][9][WIN_MAX];
  u32 winCount[TB_WARPS][9];
  // per neighbour column: the windows the warps registered, then the directory of its staged segments (sorted, disjoint)
  uint2 ranges[9][SEG_MAX];
  u32 rangeCount[9];
  u32 segStart[9][SEG_MAX], segEnd[9][SEG_MAX], segOff[9][SEG_MAX]; // segOff: offset inside the column's part of the tile
  u32 segCount[9];
  u32 colLen[9], colOff[9]; // positions staged for the column; offset of the column's part inside the tile of its stage
  u32 stageFirst[10]; // stage k covers the columns stageFirst[k] .. stageFirst[k + 1] - 1
  u32 nStages;
  u32 stageHasData;
  u32 overflow;
  float4 origin; // the filter works in coordinates relative to the CTA's first particle (see FILTER)
  unsigned long long bar;
};

// image code of a list entry: bits 28-29 x image, bits 30-31 z image (0: none, 1: +2W, 2: -2W); same encoding as sweep.cuh
__device__ __forceinline__ u32 imageCode(float sx, float sz)
{
  const u32 cx = sx > 0.0f ? 1u : (sx < 0.0f ? 2u : 0u), cz = sz > 0.0f ? 1u : (sz < 0.0f ? 2u : 0u);
  return (cx << 28) | (cz << 30);
}

// ---------------------------------------------------------------- slots

// image shift of a wrapped neighbour cell (clouds.cl:334-347): raw index beyond the grid -> +2W, below -> -2W
__device__ __forceinline__ float wrapShift(int raw, int res, float absW) { return raw >= res ? 2.0f * absW : (raw < 0 ? -2.0f * absW : 0.0f); }

struct LaneShifts
{
  float sx[3]; // by iX + 1
  float sz[3]; // by kind: pre, main, post
};
// (v + R) % R for v in [-R, 2R): the neighbour-cell indices are in [-1, R + 1]
__device__ __forceinline__ int wrapIndex(int v, int R) { return v < 0 ? v + R : (v >= R ? v - R : v); }

template <int TRAV>
__device__ __forceinline__ LaneShifts laneShifts(const GridParams& g, const int3 ci)
{
  LaneShifts r;
#pragma unroll
  for (int k = 0; k < 3; ++k)
  {
    r.sx[k] = TRAV == TRAV_CLOUDS ? wrapShift(ci.x + k - 1, g.res[0], g.absW[0]) : 0.0f;
    r.sz[k] = 0.0f;
  }
  if (TRAV == TRAV_CLOUDS)
  {
    const int RZ = g.res[2];
    const int z0 = wrapIndex(ci.z - 1, RZ), z1 = wrapIndex(ci.z, RZ), z2 = wrapIndex(ci.z + 1, RZ);
    const bool break01 = z1 != z0 + 1, break12 = z2 != z1 + 1;
    r.sz[1] = wrapShift(ci.z, RZ, g.absW[2]);
    r.sz[0] = break01 ? wrapShift(ci.z - 1, RZ, g.absW[2]) : r.sz[1];
    r.sz[2] = break12 ? wrapShift(ci.z + 1, RZ, g.absW[2]) : r.sz[1];
  }
  return r;
}

// The runs of one lane in the neighbour column (iX, iY): rs/re[kind] (kind 0 = pre, 1 = main, 2 = post), inclusive
// global indices, empty = (0xFFFFFFFF, 0). Returns false when the column cannot be described by at most one ascending
// run per kind (a capped or start = 1 cell breaks the contiguity): the warp then takes the fallback path.
template <int TRAV>
__device__ __forceinline__ bool laneColumn(const GridParams& g, const uint2* __restrict__ table, const int3 ci, const int iX, const int iY,
    const bool active, u32 (&rs)[3], u32 (&re)[3])
{
#pragma unroll
  for (int k = 0; k < 3; ++k)
  {
    rs[k] = 0xFFFFFFFFu;
    re[k] = 0u;
  }
  if (!active)
    return true;
  const int RX = g.res[0], RY = g.res[1], RZ = g.res[2];
  int cx = ci.x + iX, cy = ci.y + iY;
  if (TRAV == TRAV_BOIDS)
  {
    if (cx < 0 || cx >= RX)
      return true;
  }
  else
    cx = wrapIndex(cx, RX);
  if (TRAV == TRAV_FLUIDS)
    cy = wrapIndex(cy, RY);
  else if (cy < 0 || cy >= RY)
    return true;
  const uint2* __restrict__ row = table + (cx * RY + cy) * RZ;
  int zm[3];
  uint2 se[3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
  {
    const int zr = ci.z + k - 1;
    bool skip = false;
    if (TRAV == TRAV_BOIDS)
    {
      skip = zr < 0 || zr >= RZ;
      zm[k] = zr;
    }
    else
      zm[k] = wrapIndex(zr, RZ);
    se[k] = skip ? make_uint2(1u, 0u) : __ldg(row + zm[k]);
  }
  const bool break01 = TRAV != TRAV_BOIDS && zm[1] != zm[0] + 1, break12 = TRAV != TRAV_BOIDS && zm[2] != zm[1] + 1;
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 3; ++k)
  {
    if (se[k].x > se[k].y)
      continue;
    const int kind = k == 0 ? (break01 ? 0 : 1) : (k == 2 ? (break12 ? 2 : 1) : 1);
    // (kind is a compile-time function of k and two predicates: select the slot with predicated moves)
#pragma unroll
    for (int q = 0; q < 3; ++q)
      if (q == kind)
      {
        if (rs[q] > re[q])
        {
          rs[q] = se[k].x;
          re[q] = se[k].y;
        }
        else if (re[q] + 1u == se[k].x)
          re[q] = se[k].y;
        else
          ok = false;
      }
  }
  return ok;
}

// ---------------------------------------------------------------- TMA / mbarrier primitives (sm_90+)

__device__ __forceinline__ u32 smemAddr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(unsigned long long* bar, u32 count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(unsigned long long* bar, u32 bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned long long* bar, u32 parity)
{
  u32 ok = 0u;
  do
  {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smemAddr(bar)), "r"(parity)
        : "memory");
  } while (ok == 0u);
}
// 1-D bulk copy global -> shared, bytes a multiple of 16, both addresses 16-byte aligned; completes on the mbarrier
__device__ __forceinline__ void tmaLoad1D(void* dstSmem, const void* srcGlobal, u32 bytes, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dstSmem)),
               "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar))
               : "memory");
}

// ---------------------------------------------------------------- the builder

// The margin mask of one step / position epoch (see MASKS above): word w of lane l of warp W at
// words[(W * wordCap + w) * 32 + l], its descriptor (global index of bit 0 | slot << 27) at desc[W * wordCap + w];
// warpWords[W] = number of words, or MASK_FALLBACK: that warp builds with the per-thread traversal.
constexpr u32 MASK_FALLBACK = 0xFFFFFFFFu;
typedef MarginMaskBuffers MarginMask; // kernels.cuh

// Filter the candidates of the CTA's 128 particles (positions P, particle of this thread: pi) against sqrt(radiusSq) and
// write the margin mask; row0 = first row of the CTA. Every thread of the CTA must call (inactive threads: active = false). stats: optional counters
// { irregular warps, warps over the word capacity, CTAs over the tile capacity }.
template <int TRAV>
__device__ __forceinline__ void tileFilterToMask(TileSmem& sm, const GridParams& g, const float radiusSq, const uint2* __restrict__ table,
    const float4* __restrict__ P, const float4 pi, const bool active, const u32 row0, const MarginMask& mm, u32* __restrict__ stats)
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int3 ci = cell3D(g, pi.x, pi.y, pi.z);
  const LaneShifts sh = laneShifts<TRAV>(g, ci);

  if (tid < 9)
    sm.rangeCount[tid] = 0u;
  if (tid == 0)
  {
    sm.overflow = 0u;
    const float4 o = P[row0];
    sm.origin = (isfinite(o.x) && isfinite(o.y) && isfinite(o.z)) ? o : make_float4(0.f, 0.f, 0.f, 0.f);
    mbarInit(&sm.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // lane groups: the runs of lanes with the same centre column overlap or touch (the lanes are consecutive in the sorted
  // order); lanes in different columns have runs that can be far apart (columns of different height), hence one window
  // per group. Groups are formed by leader election; a warp with more than NGROUP columns (very sparse particles) is
  // irregular.
  bool regular = true;
  int grp = NGROUP;
  {
    unsigned todo = __ballot_sync(0xFFFFFFFFu, active);
#pragma unroll
    for (int gq = 0; gq < NGROUP; ++gq)
    {
      if (todo == 0u)
        break;
      const int leader = __ffs(todo) - 1;
      const int lx = __shfl_sync(0xFFFFFFFFu, ci.x, leader), ly = __shfl_sync(0xFFFFFFFFu, ci.y, leader);
      const bool same = active && grp == NGROUP && ci.x == lx && ci.y == ly;
      if (same)
        grp = gq;
      todo &= ~__ballot_sync(0xFFFFFFFFu, same);
    }
    if (todo != 0u)
      regular = false;
  }
  __syncthreads();

  // ---- pass A: the windows of this warp (27 slots x NGROUP groups; only the non-empty ones are kept), registered with the CTA
  if (lane < 9)
    sm.winCount[warp][lane] = 0u;
  __syncwarp();
#pragma unroll 1
  for (int col = 0; col < 9; ++col)
  {
    u32 rs[3], re[3];
    regular &= laneColumn<TRAV>(g, table, ci, col / 3 - 1, col % 3 - 1, active, rs, re);
#pragma unroll
    for (int kind = 0; kind < 3; ++kind)
    {
      const bool mine = rs[kind] <= re[kind];
      if (!__any_sync(0xFFFFFFFFu, mine))
        continue;
#pragma unroll
      for (int gq = 0; gq < NGROUP; ++gq)
      {
        const bool in = mine && grp == gq;
        if (!__any_sync(0xFFFFFFFFu, in))
          continue;
        const u32 ws = __reduce_min_sync(0xFFFFFFFFu, in ? rs[kind] : 0xFFFFFFFFu), we = __reduce_max_sync(0xFFFFFFFFu, in ? re[kind] : 0u);
        if (lane == 0)
        {
          const u32 kw = sm.winCount[warp][col];
          if (kw < (u32)WIN_MAX)
          {
            sm.winList[warp][col][kw] = make_uint4(ws, we, (u32)(col * 3 + kind) | ((u32)gq << 8), 0u);
            sm.winCount[warp][col] = kw + 1u;
          }
          const u32 k = atomicAdd(&sm.rangeCount[col], 1u);
          if (kw >= (u32)WIN_MAX || k >= (u32)SEG_MAX)
            sm.overflow = 1u;
          else
            sm.ranges[col][k] = make_uint2(ws, we);
        }
      }
    }
  }
  const bool warpRegular = __all_sync(0xFFFFFFFFu, regular);
  if (!warpRegular && lane == 0 && stats)
    atomicAdd(stats + 0, 1u);
  __syncthreads();

  // ---- the tile directory: per column, sort the registered windows and merge overlapping / touching ones
  if (sm.overflow == 0u)
  {
#pragma unroll 1
    for (int col = warp; col < 9; col += TB_WARPS)
    {
      const u32 n = sm.rangeCount[col];
      uint2 r = (u32)lane < n ? sm.ranges[col][lane] : make_uint2(0xFFFFFFFFu, 0u);
      u32 rank = 0u;
      for (u32 j = 0; j < n; ++j)
      {
        const u32 sj = __shfl_sync(0xFFFFFFFFu, r.x, j);
        rank += (sj < r.x || (sj == r.x && j < (u32)lane)) ? 1u : 0u;
      }
      __syncwarp();
      if ((u32)lane < n)
        sm.ranges[col][rank] = r;
      __syncwarp();
      r = (u32)lane < n ? sm.ranges[col][lane] : make_uint2(0xFFFFFFFFu, 0u);
      u32 pm = (u32)lane < n ? r.y : 0u; // inclusive prefix maximum of the ends
#pragma unroll
      for (int off = 1; off < SEG_MAX; off <<= 1)
      {
        const u32 t = __shfl_up_sync(0xFFFFFFFFu, pm, off);
        if (lane >= off)
          pm = max(pm, t);
      }
      const u32 prevMax = __shfl_up_sync(0xFFFFFFFFu, pm, 1);
      const bool head = (u32)lane < n && (lane == 0 || r.x > prevMax + 1u);
      const unsigned hb = __ballot_sync(0xFFFFFFFFu, head);
      const unsigned above = lane == 31 ? 0u : (hb & (0xFFFFFFFFu << (lane + 1)));
      const int lastLane = above ? (__ffs(above) - 2) : (int)n - 1;
      const u32 segEnd = __shfl_sync(0xFFFFFFFFu, pm, lastLane < 0 ? 0 : lastLane);
      const u32 len = head ? segEnd - r.x + 1u : 0u;
      u32 incl = len;
#pragma unroll
      for (int off = 1; off < SEG_MAX; off <<= 1)
      {
        const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
        if (lane >= off)
          incl += t;
      }
      if (head)
      {
        const u32 k = __popc(hb & ((1u << lane) - 1u));
        sm.segStart[col][k] = r.x;
        sm.segEnd[col][k] = segEnd;
        sm.segOff[col][k] = incl - len;
      }
      const u32 total = __shfl_sync(0xFFFFFFFFu, incl, SEG_MAX - 1);
      if (lane == 0)
      {
        sm.segCount[col] = __popc(hb);
        sm.colLen[col] = total;
        if (total > TILE_STAGE_CAP)
          sm.overflow = 1u;
      }
    }
  }
  __syncthreads();
  if (tid == 0 && sm.overflow == 0u)
  {
    // pack consecutive columns into stages that fit the tile
    u32 ns = 0u, fill = 0u;
    sm.stageFirst[0] = 0u;
    for (u32 col = 0; col < 9u; ++col)
    {
      if (fill + sm.colLen[col] > TILE_STAGE_CAP)
      {
        sm.stageFirst[++ns] = col;
        fill = 0u;
      }
      sm.colOff[col] = fill;
      fill += sm.colLen[col];
    }
    sm.stageFirst[++ns] = 9u;
    sm.nStages = ns;
  }
  __syncthreads();
  const bool useTiles = sm.overflow == 0u;
  if (!useTiles && tid == 0 && stats)
    atomicAdd(stats + 2, 1u);
  const u32 gwarp = row0 / 32u + warp; // (by rows, not by blockIdx: launches by row phase map their blocks, sweep.cuh)
  if (!useTiles)
  {
    if (lane == 0)
      mm.warpWords[gwarp] = MASK_FALLBACK;
    return; // (uniform for the CTA: nobody is left behind at a barrier)
  }

  // one thread: one bulk copy per staged segment of the stage's columns behind one mbarrier phase; false: nothing to stage
  auto issueStage = [&](u32 stage) -> bool
  {
    u32 total = 0u;
    for (u32 col = sm.stageFirst[stage]; col < sm.stageFirst[stage + 1]; ++col)
      total += sm.colLen[col];
    if (total == 0u)
      return false;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbarExpectTx(&sm.bar, total * (u32)sizeof(float4));
    for (u32 col = sm.stageFirst[stage]; col < sm.stageFirst[stage + 1]; ++col)
      for (u32 k = 0; k < sm.segCount[col]; ++k)
        tmaLoad1D(&sm.tile[sm.colOff[col] + sm.segOff[col][k]], P + sm.segStart[col][k],
            (sm.segEnd[col][k] - sm.segStart[col][k] + 1u) * (u32)sizeof(float4), &sm.bar);
    return true;
  };
  // tile index of global index `idx` of column `col` (the whole warp asks for the same idx, which lies inside a staged segment)
  auto tileIndex = [&](int col, u32 idx) -> u32
  {
    const bool hit = (u32)lane < sm.segCount[col] && sm.segStart[col][lane] <= idx && idx <= sm.segEnd[col][lane];
    const int k = __ffs(__ballot_sync(0xFFFFFFFFu, hit)) - 1;
    return sm.colOff[col] + sm.segOff[col][k] + (idx - sm.segStart[col][k]);
  };

  // (clouds: the filter uses the shifted centre pi - shift, one rounding away from the canonical (pi - pj) - shift; the
  //  list radius has 12 % of slack and the caller applies the exact support test)
  // The test runs on |c'|^2 - 2 p'.c' < r^2 - |p'|^2 (p', c' relative to the CTA's first particle): its rounding error is
  // bounded by a few ulp of D^2, D = |p'| + 6 r >= the distance of the lane's own candidates from the origin, and the
  // threshold is raised by 2e-6 D^2 (four times that bound): the mask is a superset of the canonical sq < radiusSq.
  const float4 org = sm.origin;
  const float filterSq = radiusSq, sixR = 6.0f * sqrtf(radiusSq);

  u32 wordsOut = 0u;
  const size_t wordBase = (size_t)gwarp * mm.wordCap;
  {
    u32 phase = 0u;
    const u32 nStages = sm.nStages;
#pragma unroll 1
    for (u32 stage = 0; stage < nStages; ++stage)
    {
      __syncthreads(); // every warp is through with the previous stage's tile
      if (tid == 0)
        sm.stageHasData = issueStage(stage) ? 1u : 0u;
      __syncthreads();
      if (sm.stageHasData == 0u)
        continue;
      mbarWait(&sm.bar, phase);
      phase ^= 1u;
      {
        // staged positions -> (c - origin, |c - origin|^2): the filter's squared distance becomes three FFMAs per candidate
        const u32 lastCol = sm.stageFirst[stage + 1] - 1u;
        const u32 staged = sm.colOff[lastCol] + sm.colLen[lastCol];
        for (u32 j = (u32)tid; j < staged; j += (u32)TB_THREADS)
        {
          float4 q = sm.tile[j];
          q.x -= org.x;
          q.y -= org.y;
          q.z -= org.z;
          q.w = fmaf(q.z, q.z, fmaf(q.y, q.y, q.x * q.x));
          sm.tile[j] = q;
        }
        __syncthreads();
      }
      if (!warpRegular)
        continue;
      const float4* tileBuf = sm.tile;
#pragma unroll 1
      for (int col = (int)sm.stageFirst[stage]; col < (int)sm.stageFirst[stage + 1]; ++col)
      {
        const u32 nWin = sm.winCount[warp][col];
        if (nWin == 0u)
          continue;
        const int ix = col / 3;
        u32 rs[3], re[3];
        laneColumn<TRAV>(g, table, ci, ix - 1, col - ix * 3 - 1, active, rs, re);
        const float px = TRAV == TRAV_CLOUDS ? pi.x - (ix == 0 ? sh.sx[0] : (ix == 1 ? sh.sx[1] : sh.sx[2])) : pi.x, py = pi.y;
#pragma unroll 1
        for (u32 kw = 0; kw < nWin; ++kw)
        {
          const uint4 wl = sm.winList[warp][col][kw];
          const int slot = (int)(wl.z & 0xFFu), gq = (int)(wl.z >> 8), kind = slot - col * 3;
          const float pz = TRAV == TRAV_CLOUDS ? pi.z - (kind == 0 ? sh.sz[0] : (kind == 1 ? sh.sz[1] : sh.sz[2])) : pi.z;
          // this lane's own run, if it belongs to the group
          const bool in = grp == gq;
          const u32 s0 = !in ? 0xFFFFFFFFu : (kind == 0 ? rs[0] : (kind == 1 ? rs[1] : rs[2]));
          const u32 e0 = !in ? 0u : (kind == 0 ? re[0] : (kind == 1 ? re[1] : re[2]));
          const u32 tileBase = tileIndex(col, wl.x);
          const u32 nW = ((wl.y - wl.x) >> 5) + 1u;
          const float ppx = px - org.x, ppy = py - org.y, ppz = pz - org.z;
          const float n2 = fmaf(ppz, ppz, fmaf(ppy, ppy, ppx * ppx)), D = sqrtf(n2) + sixR;
          const float thr = (filterSq + 2e-6f * D * D) - n2;
          const float ax = -2.0f * ppx, ay = -2.0f * ppy, az = -2.0f * ppz;
#pragma unroll 1
          for (u32 w = 0; w < nW; ++w)
          {
            const u32 B = wl.x + 32u * w;
            // the candidates of this word that are in the lane's own run [s0, e0]
            u32 rm = 0u;
            if (s0 <= e0 && s0 <= B + 31u && e0 >= B)
            {
              const u32 lo = s0 > B ? s0 - B : 0u, hi = e0 < B + 31u ? e0 - B : 31u;
              rm = (0xFFFFFFFFu >> (31u - hi)) & (0xFFFFFFFFu << lo);
            }
            if (!__any_sync(0xFFFFFFFFu, rm != 0u))
              continue; // nobody's run reaches into this word (a hole between two families of runs)
            const float4* __restrict__ tp = tileBuf + tileBase + 32u * w;
            u32 m = 0u;
#pragma unroll
            for (int k = 0; k < 32; ++k)
            {
              const float4 q = tp[k];
              if (fmaf(ax, q.x, fmaf(ay, q.y, fmaf(az, q.z, q.w))) < thr)
                m |= 1u << k;
            }
            if (wordsOut < mm.wordCap)
            {
              mm.words[(wordBase + wordsOut) * 32u + lane] = m & rm;
              if (lane == 0)
                mm.desc[wordBase + wordsOut] = B | ((u32)slot << TB_INDEX_BITS);
            }
            ++wordsOut;
          }
        }
      }
    }
  }
  const bool fits = wordsOut <= mm.wordCap;
  if (lane == 0)
  {
    mm.warpWords[gwarp] = (warpRegular && fits) ? wordsOut : MASK_FALLBACK;
    if (warpRegular && !fits && stats)
      atomicAdd(stats + 1, 1u);
  }
}

} // namespace rtp
