// sort.cu -- onesweep radix sort kernels (see sort.cuh for the design and the reference contract).
#include "sort.cuh"

namespace rtp
{
__device__ __forceinline__ u32 ldRelaxed(const u32* p)
{
  u32 v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stRelaxed(u32* p, u32 v)
{
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// exclusive prefix over the 256 threads of a CTA (warp shuffles + 8 partials)
__device__ __forceinline__ u32 blockExclusiveScan256(u32 v, u32* sWarp)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u32 incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1)
  {
    const u32 t = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off)
      incl += t;
  }
  if (lane == 31)
    sWarp[warp] = incl;
  __syncthreads();
  u32 base = 0;
#pragma unroll
  for (int w = 0; w < SORT_WARPS; ++w)
    base += (w < warp) ? sWarp[w] : 0u;
  __syncthreads();
  return base + incl - v;
}

// One read of the keys builds the digit histogram of every pass; side job: zero the look-back status words.
__global__ void __launch_bounds__(SORT_THREADS) sortHistogramKernel(const u32* __restrict__ keys, u32 n, int passes,
    PassDesc desc, u32* __restrict__ ctrl, u32* __restrict__ status, size_t statusWords)
{
  RTP_PDL_PROLOGUE();
  __shared__ u32 sHist[SORT_MAX_PASSES * SORT_RADIX];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < passes * SORT_RADIX; i += SORT_THREADS)
    sHist[i] = 0;
  const size_t gtid = (size_t)blockIdx.x * SORT_THREADS + tid, gstride = (size_t)gridDim.x * SORT_THREADS;
  for (size_t i = gtid; i < statusWords; i += gstride)
    status[i] = 0;
  __syncthreads();
  // whole warps iterate together so that __match_any_sync sees a full mask
  const size_t nRound = ((size_t)n + 31) & ~(size_t)31;
  for (size_t i = gtid; i < nRound; i += gstride)
  {
    const bool valid = i < n;
    const u32 k = valid ? __ldg(keys + i) : 0u;
    for (int p = 0; p < passes; ++p)
    {
      const u32 d = (k >> desc.shift[p]) & desc.mask[p];
      const u32 peers = __match_any_sync(0xffffffffu, valid ? d : (SORT_RADIX + lane));
      if (valid && lane == (__ffs(peers) - 1))
        atomicAdd(&sHist[p * SORT_RADIX + d], __popc(peers));
    }
  }
  __syncthreads();
  for (int i = tid; i < passes * SORT_RADIX; i += SORT_THREADS)
    if (sHist[i])
      atomicAdd(&ctrl[i], sHist[i]);
}

// Large tiles (ITEMS = 16) fetch their values only when the tile is reordered (after the look-back): sixteen registers less
// per thread through the ranking and the look-back, four resident CTAs per SM instead of three.
template <int ITEMS>
__global__ void __launch_bounds__(SORT_THREADS, ITEMS > 4 ? 4 : 1) onesweepPassKernel(const u32* __restrict__ kin, const u32* __restrict__ vin,
    u32* __restrict__ kout, u32* __restrict__ vout, u32 n, int shift, u32 mask, const u32* __restrict__ ghist,
    u32* __restrict__ status, u32* __restrict__ ticket)
{
  RTP_PDL_PROLOGUE();
  constexpr int TILE = SORT_THREADS * ITEMS;
  __shared__ u32 sWarpHist[SORT_WARPS][SORT_RADIX];
  __shared__ u32 sKeys[TILE];
  __shared__ u32 sVals[TILE];
  __shared__ u32 sGBase[SORT_RADIX];
  __shared__ u32 sLBase[SORT_RADIX];
  __shared__ u32 sScan[SORT_WARPS];
  __shared__ u32 sTile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0)
    sTile = atomicAdd(ticket, 1u);
  for (int i = tid; i < SORT_WARPS * SORT_RADIX; i += SORT_THREADS)
    (&sWarpHist[0][0])[i] = 0;
  __syncthreads();
  const u32 tile = sTile;
  const u32 tileBase = tile * TILE;

  // warp-striped load: warp w owns [tileBase + w*32*ITEMS, +32*ITEMS), item r of lane l = base + r*32 + l
  constexpr bool LATE_VALS = ITEMS > 4;
  u32 key[ITEMS], val[LATE_VALS ? 1 : ITEMS], rank[ITEMS];
  const u32 warpBase = tileBase + warp * 32 * ITEMS;
#pragma unroll
  for (int r = 0; r < ITEMS; ++r)
  {
    const u32 idx = warpBase + r * 32 + lane;
    const bool valid = idx < n;
    key[r] = valid ? __ldg(kin + idx) : 0xFFFFFFFFu;
    if (!LATE_VALS)
      val[r] = valid ? (vin ? __ldg(vin + idx) : idx) : 0u;
  }

  // stable ranking inside the warp: (round, lane) order == memory order
  const u32 ltMask = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < ITEMS; ++r)
  {
    const bool valid = (warpBase + r * 32 + lane) < n;
    const u32 d = (key[r] >> shift) & mask;
    const u32 peers = __match_any_sync(0xffffffffu, valid ? d : (SORT_RADIX + lane));
    const u32 pre = sWarpHist[warp][d];
    __syncwarp();
    if (valid && lane == (__ffs(peers) - 1))
      sWarpHist[warp][d] = pre + __popc(peers);
    __syncwarp();
    rank[r] = pre + __popc(peers & ltMask);
  }
  __syncthreads();

  // thread d owns digit d: exclusive scan over warps -> warp offsets, tile total
  const int d = tid;
  u32 tot = 0;
#pragma unroll
  for (int w = 0; w < SORT_WARPS; ++w)
  {
    const u32 t = sWarpHist[w][d];
    sWarpHist[w][d] = tot;
    tot += t;
  }
  u32* st = status + (size_t)tile * SORT_RADIX;
  stRelaxed(st + d, tot | (tile == 0 ? SORT_FLAG_PREFIX : SORT_FLAG_AGGREGATE));

  const u32 gscan = blockExclusiveScan256(__ldg(ghist + d), sScan);
  const u32 lscan = blockExclusiveScan256(tot, sScan);

  // decoupled look-back over the previous tiles' words of this digit
  u32 excl = 0;
  if (tile > 0)
  {
    int t = (int)tile - 1;
    while (true)
    {
      const u32 s = ldRelaxed(status + (size_t)t * SORT_RADIX + d);
      const u32 flag = s & SORT_FLAG_MASK;
      if (flag == 0)
        continue;
      excl += s & SORT_VALUE_MASK;
      if (flag == SORT_FLAG_PREFIX)
        break;
      --t;
    }
    stRelaxed(st + d, (excl + tot) | SORT_FLAG_PREFIX);
  }
  sGBase[d] = gscan + excl - lscan;
  sLBase[d] = lscan;
  __syncthreads();

  if (LATE_VALS)
  {
    u32 v[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; ++r)
    {
      const u32 idx = warpBase + r * 32 + lane;
      v[r] = (vin && idx < n) ? __ldg(vin + idx) : idx;
    }
#pragma unroll
    for (int r = 0; r < ITEMS; ++r)
    {
      if ((warpBase + r * 32 + lane) < n)
      {
        const u32 dd = (key[r] >> shift) & mask;
        const u32 lp = sLBase[dd] + sWarpHist[warp][dd] + rank[r];
        sKeys[lp] = key[r];
        sVals[lp] = v[r];
      }
    }
  }
  else
  {
#pragma unroll
    for (int r = 0; r < ITEMS; ++r)
    {
      if ((warpBase + r * 32 + lane) < n)
      {
        const u32 dd = (key[r] >> shift) & mask;
        const u32 lp = sLBase[dd] + sWarpHist[warp][dd] + rank[r];
        sKeys[lp] = key[r];
        sVals[lp] = val[LATE_VALS ? 0 : r];
      }
    }
  }
  __syncthreads();

  const u32 nValid = min((u32)TILE, n - tileBase);
  for (u32 j = tid; j < nValid; j += SORT_THREADS)
  {
    const u32 k = sKeys[j];
    const u32 pos = sGBase[(k >> shift) & mask] + j;
    kout[pos] = k;
    vout[pos] = sVals[j];
  }
}

SortPlan makeSortPlan(u32 n, int keyBits)
{
  SortPlan p;
  p.n = n;
  if (keyBits < 1)
    keyBits = 1;
  if (keyBits > 32)
    keyBits = 32;
  p.passes = (keyBits + SORT_RADIX_BITS - 1) / SORT_RADIX_BITS;
  for (int i = 0; i < p.passes; ++i)
  {
    p.shift[i] = SORT_RADIX_BITS * i;
    const int rem = keyBits - SORT_RADIX_BITS * i;
    p.bits[i] = rem < SORT_RADIX_BITS ? rem : SORT_RADIX_BITS;
  }
  // small inputs: small tiles so that every SM gets one; large inputs: amortise the per-tile 256-bin bookkeeping
  p.itemsPerThread = (n <= (1u << 20)) ? 4 : 16;
  const u32 tile = SORT_THREADS * p.itemsPerThread;
  p.tiles = (n + tile - 1) / tile;
  return p;
}

size_t sortStatusWords(const SortPlan& plan) { return (size_t)plan.passes * plan.tiles * SORT_RADIX; }

PassDesc makePassDesc(const SortPlan& plan)
{
  PassDesc desc;
  for (int i = 0; i < SORT_MAX_PASSES; ++i)
  {
    desc.shift[i] = plan.shift[i];
    desc.mask[i] = plan.bits[i] ? ((1u << plan.bits[i]) - 1u) : 0u;
  }
  return desc;
}

void enqueueSortBegin(const SortPlan& plan, u32* ctrl, cudaStream_t stream)
{
  if (plan.n)
    cudaMemsetAsync(ctrl, 0, SORT_CTRL_WORDS * sizeof(u32), stream); // memset node, not counted as a kernel
}

int enqueueSortPasses(const SortPlan& plan, u32* keys0, u32* vals0, u32* keys1, u32* vals1, u32* ctrl, u32* status,
    cudaStream_t stream)
{
  if (plan.n == 0)
    return 0;
  const PassDesc desc = makePassDesc(plan);
  u32* kbuf[2] = { keys0, keys1 };
  u32* vbuf[2] = { vals0, vals1 };
  int launches = 0;
  for (int p = 0; p < plan.passes; ++p)
  {
    const int src = ((plan.passes - p) % 2 == 0) ? 0 : 1;
    const int dst = 1 - src;
    const u32* vin = (p == 0) ? nullptr : vbuf[src];
    u32* st = status + (size_t)p * plan.tiles * SORT_RADIX;
    u32* ticket = ctrl + SORT_MAX_PASSES * SORT_RADIX + p;
    const u32* gh = ctrl + p * SORT_RADIX;
    if (plan.itemsPerThread == 4)
      launchKernel(onesweepPassKernel<4>, plan.tiles, SORT_THREADS, stream, kbuf[src], vin, kbuf[dst], vbuf[dst], plan.n,
          desc.shift[p], desc.mask[p], gh, st, ticket);
    else
      launchKernel(onesweepPassKernel<16>, plan.tiles, SORT_THREADS, stream, kbuf[src], vin, kbuf[dst], vbuf[dst], plan.n,
          desc.shift[p], desc.mask[p], gh, st, ticket);
    ++launches;
  }
  return launches;
}

int enqueueSort(const SortPlan& plan, u32* keys0, u32* vals0, u32* keys1, u32* vals1, u32* ctrl, u32* status,
    cudaStream_t stream)
{
  if (plan.n == 0)
    return 0;
  enqueueSortBegin(plan, ctrl, stream);
  const PassDesc desc = makePassDesc(plan);
  u32* kbuf[2] = { keys0, keys1 };
  const int first = (plan.passes % 2 == 0) ? 0 : 1;
  int blocks = (int)((plan.n + SORT_THREADS * 8 - 1) / (SORT_THREADS * 8));
  if (blocks > 148 * 8)
    blocks = 148 * 8;
  if (blocks < 1)
    blocks = 1;
  launchKernel(sortHistogramKernel, blocks, SORT_THREADS, stream, kbuf[first], plan.n, plan.passes, desc, ctrl, status,
      sortStatusWords(plan));
  return 1 + enqueueSortPasses(plan, keys0, vals0, keys1, vals1, ctrl, status, stream);
}

} // namespace rtp
