// slab_group.cu -- the x-slab decomposition of the PBF step driven from INSIDE the library (rtp_slab_group_*, include/rtp_cuda.h):
// one host thread, one rtp handle per slab, any assignment of slabs to the GPUs of the box (several slabs may share a GPU).
// New design (the reference is single-device); the Python orchestration realtimeparticles_b200/sharded.py
// (one process per GPU over torch.distributed) runs the same step and the two are bit-identical (tests/test_sharded.py).
//
// Everything is written above the stage API of the C ABI (rtp_shard_stage_rows / pack / unpack / classify / ...), so the
// kernels are the single-GPU ones. What this file adds is the part sharded.py does with torch: the order-preserving
// compaction of the face masks into fixed-capacity index lists, the index maps of a refresh, and the TRANSPORT -- a slab
// PULLS its neighbour's packed rows straight into its own rows with cudaMemcpyAsync over peer memory (NVLink between two
// GPUs, a device copy inside one), ordered by events between the slabs' streams. Nothing returns to the host inside a step:
// rtp_slab_group_step only enqueues; capacities are checked on the device and read back by rtp_slab_group_check.
//
// Row layout of a slab (static, see sharded.py): rows [0, S) the slab's own region -- particles, cell-sorted, at the front,
// "no particle" rows (+inf position) behind them, the last 2 x migrate_cap rows arrival slots --, then one ghost region of
// ghost_cap rows per face. Events: a send buffer is re-packed only after the neighbour has pulled its previous content
// (evPulled), a pull starts only after the neighbour's pack (evPacked).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/rtp_cuda.h"

namespace
{
typedef uint32_t u32;
typedef unsigned long long u64;
constexpr u32 NO_ROW = 0xFFFFFFFFu;
constexpr int CP_THREADS = 256, CP_ITEMS = 16, CP_TILE = CP_THREADS * CP_ITEMS;

// ---- order-preserving compaction of a byte mask: idx[k] = k-th row with mask != 0, padded with NO_ROW up to cap

__device__ __forceinline__ u32 threadItems(const uint8_t* __restrict__ mask, u32 n, u32 first, u32& bits)
{
  // 16 consecutive rows of one thread -> a 16-bit pattern; returns its population count
  bits = 0u;
  if (first + CP_ITEMS <= n && ((size_t)(mask + first) & 15u) == 0u)
  {
    const uint4 v = *reinterpret_cast<const uint4*>(mask + first);
    const u32 w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int b = 0; b < 4; ++b)
        bits |= ((w[q] >> (8 * b)) & 0xFFu) ? (1u << (4 * q + b)) : 0u;
  }
  else
  {
    for (int k = 0; k < CP_ITEMS; ++k)
      if (first + k < n && mask[first + k])
        bits |= 1u << k;
  }
  return __popc(bits);
}

__global__ void __launch_bounds__(CP_THREADS) maskCountKernel(const uint8_t* __restrict__ mask, u32 n, u32* __restrict__ tileCount)
{
  __shared__ u32 sTot;
  if (threadIdx.x == 0)
    sTot = 0u;
  __syncthreads();
  u32 bits;
  u32 c = threadItems(mask, n, blockIdx.x * CP_TILE + threadIdx.x * CP_ITEMS, bits);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
    c += __shfl_down_sync(0xFFFFFFFFu, c, off);
  if ((threadIdx.x & 31) == 0 && c)
    atomicAdd(&sTot, c);
  __syncthreads();
  if (threadIdx.x == 0)
    tileCount[blockIdx.x] = sTot;
}

// one block: exclusive scan of the tile counts in place; total -> *count; capacity flag and running total of the caller
__global__ void __launch_bounds__(1024) maskScanKernel(u32* __restrict__ tileCount, u32 nTiles, u32 cap, u32* __restrict__ count,
    u64* __restrict__ flags, u64 errBit, int addToMigrated)
{
  __shared__ u32 sWarp[32];
  __shared__ u32 sCarry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0)
    sCarry = 0u;
  __syncthreads();
  for (u32 base = 0; base < nTiles; base += 1024u)
  {
    const u32 i = base + threadIdx.x;
    const u32 v = i < nTiles ? tileCount[i] : 0u;
    u32 incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1)
    {
      const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
      if (lane >= off)
        incl += t;
    }
    if (lane == 31)
      sWarp[warp] = incl;
    __syncthreads();
    u32 wbase = 0u;
    for (int w = 0; w < warp; ++w)
      wbase += sWarp[w];
    const u32 carry = sCarry;
    if (i < nTiles)
      tileCount[i] = carry + wbase + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023)
      sCarry = carry + wbase + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0)
  {
    const u32 total = sCarry;
    *count = total;
    if (total > cap)
      atomicOr(flags + 0, errBit);
    if (addToMigrated)
      atomicAdd(flags + 1, (u64)total);
  }
}

__global__ void __launch_bounds__(CP_THREADS) maskScatterKernel(const uint8_t* __restrict__ mask, u32 n, const u32* __restrict__ tileOffset,
    const u32* __restrict__ count, u32 cap, u32* __restrict__ idx)
{
  __shared__ u32 sWarp[CP_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 first = blockIdx.x * CP_TILE + threadIdx.x * CP_ITEMS;
  u32 bits;
  const u32 c = threadItems(mask, n, first, bits);
  u32 incl = c;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1)
  {
    const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
    if (lane >= off)
      incl += t;
  }
  if (lane == 31)
    sWarp[warp] = incl;
  __syncthreads();
  u32 out = tileOffset[blockIdx.x] + incl - c;
  for (int w = 0; w < warp; ++w)
    out += sWarp[w];
  while (bits)
  {
    const u32 k = __ffs(bits) - 1u;
    bits &= bits - 1u;
    if (out < cap)
      idx[out] = first + k;
    ++out;
  }
  // padding behind the last entry
  const u32 total = *count;
  for (u32 j = total + blockIdx.x * CP_THREADS + threadIdx.x; j < cap; j += gridDim.x * CP_THREADS)
    idx[j] = NO_ROW;
}

// the arrival slots [first, first + n) must hold no particle when a step begins
__global__ void arrivalSlotsFreeKernel(const float4* __restrict__ pos, u32 first, u32 n, u64* __restrict__ flags)
{
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && isfinite(pos[first + i].x))
    atomicOr(flags + 0, 1ull);
}

// index lists of the refreshes of a step (both faces in one list, face f at [f * G, (f + 1) * G)):
//   send[f G + k] = sorted row of the k-th halo row of face f (inv[src_f[k]]), recv[f G + k] = sorted row of ghost k of face f
// (NO_ROW for padding and for a face without neighbour)
__global__ void indexMapsKernel(const u32* __restrict__ inv, const u32* __restrict__ srcL, const u32* __restrict__ srcR, u32 G, u32 S,
    int hasL, int hasR, u32* __restrict__ send, u32* __restrict__ recv)
{
  const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= 2u * G)
    return;
  const u32 f = j / G, k = j - f * G;
  const bool has = f == 0u ? hasL != 0 : hasR != 0;
  const u32 src = has ? (f == 0u ? srcL[k] : srcR[k]) : NO_ROW;
  send[j] = src != NO_ROW ? inv[src] : NO_ROW;
  recv[j] = has ? inv[S + j] : NO_ROW;
}

__global__ void fillNoParticleKernel(float4* __restrict__ rows, u32 n)
{
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    rows[i] = make_float4(INFINITY, INFINITY, INFINITY, 0.0f);
}

enum RowBuf
{
  ROWS_W1 = 0, // first field of a refresh, scalar rows (lambda, |vorticity|)
  ROWS_W4 = 1, // first field of a refresh, float4 rows (predicted position, confined velocity)
  ROWS_W4B = 2, // second field of a refresh (velocity, with the last correction's positions)
  ROWS_COUNT = 3
};

struct Slab
{
  rtp_handle* h = nullptr;
  int dev = 0, rank = 0;
  bool hasL = false, hasR = false;
  u32 xlo = 0, xhi = 0;
  u32 cap = 0, S = 0, A0 = 0, G = 0, Mc = 0;
  cudaStream_t st = nullptr, ex = nullptr;
  uint8_t* mask[2] = { nullptr, nullptr };
  u32 *idxMig[2] = { nullptr, nullptr }, *src[2] = { nullptr, nullptr };
  u32 *inv = nullptr, *sendAll = nullptr, *recvAll = nullptr, *tileCount = nullptr, *cnt = nullptr;
  float *migS[2] = { nullptr, nullptr }, *haloS[2] = { nullptr, nullptr };
  float *rowS[ROWS_COUNT] = {}, *rowR[ROWS_COUNT] = {};
  float4* noParticle = nullptr;
  u64* flags = nullptr; // { capacity bits: 1 arrival slots in use, 2 migration message, 4 ghost region ; migrated total }
  cudaEvent_t evMigPacked = nullptr, evHaloPacked = nullptr, evRowPacked[ROWS_COUNT] = {};
  cudaEvent_t evMigPulled[2] = {}, evHaloPulled[2] = {}, evRowPulled[ROWS_COUNT][2] = {};
  std::vector<void*> allocs;
};
} // namespace

struct rtp_slab_group
{
  std::vector<Slab> slabs;
  u32 grid[3] = { 0, 0, 0 }, box[3] = { 0, 0, 0 };
  int jacobi = 3;
  bool vorticity = true, overlap = false;
  u64 steps = 0;
  std::string err;
};

static thread_local std::string g_groupCreateError;

#define SG_CUDA(g, call)                                                                                      \
  do                                                                                                          \
  {                                                                                                           \
    const cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess)                                                                                    \
    {                                                                                                         \
      (g)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                         \
      return RTP_ERR_CUDA;                                                                                    \
    }                                                                                                         \
  } while (0)
#define SG_RTP(g, sl, call)                                                                                   \
  do                                                                                                          \
  {                                                                                                           \
    const int rc_ = (call);                                                                                   \
    if (rc_ != RTP_OK)                                                                                        \
    {                                                                                                         \
      (g)->err = std::string(#call) + ": " + rtp_last_error((sl).h);                                         \
      return rc_;                                                                                             \
    }                                                                                                         \
  } while (0)

template <typename T>
static cudaError_t slabAlloc(Slab& sl, T** p, size_t count)
{
  void* q = nullptr;
  const cudaError_t e = cudaMalloc(&q, (count ? count : 1) * sizeof(T));
  if (e == cudaSuccess)
  {
    sl.allocs.push_back(q);
    *p = (T*)q;
  }
  return e;
}

static inline unsigned blocksFor(u64 n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

// buffer of the stage API as a typed device pointer
template <typename T>
static int shardBuf(rtp_slab_group* g, Slab& sl, int which, T** p)
{
  void* q = nullptr;
  SG_RTP(g, sl, rtp_shard_buffer(sl.h, which, &q, nullptr));
  *p = (T*)q;
  return RTP_OK;
}

// idx[0, cap) = rows of mask (n rows) in ascending order, NO_ROW padded; errBit raised when they do not fit
static int compactMask(rtp_slab_group* g, Slab& sl, const uint8_t* mask, u32 n, u32 cap, u32* idx, u64 errBit, bool migrated)
{
  const u32 tiles = (n + CP_TILE - 1) / CP_TILE;
  maskCountKernel<<<tiles, CP_THREADS, 0, sl.st>>>(mask, n, sl.tileCount);
  maskScanKernel<<<1, 1024, 0, sl.st>>>(sl.tileCount, tiles, cap, sl.cnt, sl.flags, errBit, migrated ? 1 : 0);
  maskScatterKernel<<<tiles, CP_THREADS, 0, sl.st>>>(mask, n, sl.tileCount, sl.cnt, cap, idx);
  SG_CUDA(g, cudaGetLastError());
  return RTP_OK;
}

static int setCounts(rtp_slab_group* g, Slab& sl, u64 owned, u64 local)
{
  SG_RTP(g, sl, rtp_set_nb_particles(sl.h, local));
  SG_RTP(g, sl, rtp_shard_set_owned(sl.h, owned));
  return RTP_OK;
}

extern "C" const char* rtp_slab_group_last_error(const rtp_slab_group* g) { return g ? g->err.c_str() : g_groupCreateError.c_str(); }

extern "C" void rtp_slab_group_destroy(rtp_slab_group* g)
{
  if (!g)
    return;
  for (Slab& sl : g->slabs)
  {
    cudaSetDevice(sl.dev);
    if (sl.h)
      rtp_sync(sl.h);
    for (void* p : sl.allocs)
      cudaFree(p);
    cudaEvent_t evs[] = { sl.evMigPacked, sl.evHaloPacked, sl.evRowPacked[0], sl.evRowPacked[1], sl.evRowPacked[2], sl.evMigPulled[0],
      sl.evMigPulled[1], sl.evHaloPulled[0], sl.evHaloPulled[1], sl.evRowPulled[0][0], sl.evRowPulled[0][1], sl.evRowPulled[1][0],
      sl.evRowPulled[1][1], sl.evRowPulled[2][0], sl.evRowPulled[2][1] };
    for (cudaEvent_t e : evs)
      if (e)
        cudaEventDestroy(e);
    if (sl.h)
      rtp_destroy(sl.h);
  }
  delete g;
}

extern "C" int rtp_slab_group_create(rtp_slab_group** out, int nslabs, const int* dev_ids, uint64_t slab_capacity, const uint32_t box[3],
    const uint32_t grid[3], uint64_t ghost_cap, uint64_t migrate_cap, int overlap)
{
  if (!out || nslabs < 1 || !dev_ids || !box || !grid || slab_capacity == 0 || slab_capacity >= (1ull << 31))
    return RTP_ERR_INVALID;
  *out = nullptr;
  rtp_slab_group* g = new rtp_slab_group();
  auto failCreate = [&](int rc, const std::string& why)
  {
    g_groupCreateError = why.empty() ? g->err : why;
    rtp_slab_group_destroy(g);
    return rc;
  };
  memcpy(g->grid, grid, sizeof g->grid);
  memcpy(g->box, box, sizeof g->box);
  g->slabs.resize(nslabs);
  const u32 G = nslabs > 1 ? (u32)(ghost_cap ? ghost_cap : slab_capacity / 8) : 0u;
  const u32 Mc = (u32)(migrate_cap ? migrate_cap : (G / 4 > 1 ? G / 4 : 1));
  g->overlap = overlap != 0 && nslabs > 1;
  for (int r = 0; r < nslabs; ++r)
  {
    Slab& sl = g->slabs[r];
    sl.rank = r, sl.dev = dev_ids[r];
    sl.hasL = r > 0, sl.hasR = r + 1 < nslabs;
    sl.xlo = (u32)(((u64)grid[0] * r) / nslabs), sl.xhi = (u32)(((u64)grid[0] * (r + 1)) / nslabs);
    if (nslabs > 1 && sl.xhi - sl.xlo < 2u * RTP_SHARD_GHOST_LAYERS)
      return failCreate(RTP_ERR_INVALID, "slab thinner than 2 x ghost layers");
    sl.cap = (u32)slab_capacity, sl.G = G, sl.Mc = Mc;
    if ((u64)2 * G + 2 * Mc >= slab_capacity)
      return failCreate(RTP_ERR_INVALID, "slab capacity too small for the ghost regions and arrival slots");
    sl.S = sl.cap - 2u * G;
    sl.A0 = sl.S - 2u * Mc;
    rtp_config cfg = {};
    cfg.model = RTP_MODEL_FLUIDS, cfg.device = sl.dev, cfg.max_particles = slab_capacity, cfg.nb_particles = 0;
    memcpy(cfg.box, box, sizeof cfg.box);
    memcpy(cfg.grid, grid, sizeof cfg.grid);
    cfg.dim = 3, cfg.max_parts_in_cell = 0;
    if (rtp_create(&cfg, &sl.h) != RTP_OK)
      return failCreate(RTP_ERR_CUDA, std::string("rtp_create: ") + rtp_last_error(nullptr));
    if (cudaSetDevice(sl.dev) != cudaSuccess)
      return failCreate(RTP_ERR_CUDA, "cudaSetDevice");
    void* st = nullptr;
    rtp_get_stream(sl.h, &st);
    sl.st = (cudaStream_t)st;
    cudaError_t e = cudaSuccess;
    auto A = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
    for (int s = 0; s < 2; ++s)
    {
      A(slabAlloc(sl, &sl.mask[s], (size_t)sl.S));
      A(slabAlloc(sl, &sl.idxMig[s], (size_t)Mc));
      A(slabAlloc(sl, &sl.src[s], (size_t)G));
      A(slabAlloc(sl, &sl.migS[s], (size_t)Mc * 8));
      A(slabAlloc(sl, &sl.haloS[s], (size_t)G * 4));
      A(cudaEventCreateWithFlags(&sl.evMigPulled[s], cudaEventDisableTiming));
      A(cudaEventCreateWithFlags(&sl.evHaloPulled[s], cudaEventDisableTiming));
      for (int b = 0; b < ROWS_COUNT; ++b)
        A(cudaEventCreateWithFlags(&sl.evRowPulled[b][s], cudaEventDisableTiming));
    }
    A(slabAlloc(sl, &sl.inv, (size_t)sl.cap));
    A(slabAlloc(sl, &sl.sendAll, (size_t)2 * G));
    A(slabAlloc(sl, &sl.recvAll, (size_t)2 * G));
    A(slabAlloc(sl, &sl.tileCount, (size_t)(sl.cap + CP_TILE - 1) / CP_TILE));
    A(slabAlloc(sl, &sl.cnt, 4));
    A(slabAlloc(sl, &sl.flags, 2));
    A(slabAlloc(sl, &sl.noParticle, (size_t)(G > 2 * Mc ? G : 2 * Mc)));
    for (int b = 0; b < ROWS_COUNT; ++b)
    {
      const size_t w = b == ROWS_W1 ? 1 : 4;
      A(slabAlloc(sl, &sl.rowS[b], (size_t)2 * G * w));
      A(slabAlloc(sl, &sl.rowR[b], (size_t)2 * G * w));
      if (e == cudaSuccess)
        A(cudaMemsetAsync(sl.rowR[b], 0, (size_t)(2 * G * w ? 2 * G * w : 1) * sizeof(float), sl.st));
      A(cudaEventCreateWithFlags(&sl.evRowPacked[b], cudaEventDisableTiming));
    }
    A(cudaEventCreateWithFlags(&sl.evMigPacked, cudaEventDisableTiming));
    A(cudaEventCreateWithFlags(&sl.evHaloPacked, cudaEventDisableTiming));
    if (e != cudaSuccess)
      return failCreate(RTP_ERR_CUDA, std::string("slab buffers: ") + cudaGetErrorString(e));
    cudaMemsetAsync(sl.flags, 0, 2 * sizeof(u64), sl.st);
    const u32 nNo = G > 2 * Mc ? G : 2 * Mc;
    fillNoParticleKernel<<<blocksFor(nNo, 256), 256, 0, sl.st>>>(sl.noParticle, nNo);
    if (g->overlap)
    {
      const u32 plane = grid[1] * grid[2];
      const u32 lo = sl.xlo + (sl.hasL ? RTP_SHARD_GHOST_LAYERS : 0), hi = sl.xhi - (sl.hasR ? RTP_SHARD_GHOST_LAYERS : 0);
      const int sides = (sl.hasL ? 1 : 0) + (sl.hasR ? 1 : 0);
      if (rtp_shard_set_interior(sl.h, lo * plane, (hi > lo ? hi : lo) * plane, (u64)2 * G * sides) != RTP_OK)
        return failCreate(RTP_ERR_CUDA, std::string("rtp_shard_set_interior: ") + rtp_last_error(sl.h));
      void* ex = nullptr;
      if (rtp_shard_exchange_stream(sl.h, &ex) != RTP_OK)
        return failCreate(RTP_ERR_CUDA, std::string("rtp_shard_exchange_stream: ") + rtp_last_error(sl.h));
      sl.ex = (cudaStream_t)ex;
    }
  }
  // neighbours on different GPUs copy over peer memory (NVLink); without peer access the copies are staged by the driver
  for (int r = 0; r + 1 < nslabs; ++r)
  {
    const int a = g->slabs[r].dev, b = g->slabs[r + 1].dev;
    if (a == b)
      continue;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, a, b) == cudaSuccess && can)
    {
      cudaSetDevice(a);
      cudaDeviceEnablePeerAccess(b, 0);
      cudaSetDevice(b);
      cudaDeviceEnablePeerAccess(a, 0);
      (void)cudaGetLastError(); // (already enabled is fine)
    }
  }
  *out = g;
  return RTP_OK;
}

extern "C" int rtp_slab_group_size(const rtp_slab_group* g) { return g ? (int)g->slabs.size() : 0; }

extern "C" int rtp_slab_group_handle(rtp_slab_group* g, int slab, rtp_handle** h)
{
  if (!g || !h || slab < 0 || slab >= (int)g->slabs.size())
    return RTP_ERR_INVALID;
  *h = g->slabs[slab].h;
  return RTP_OK;
}

extern "C" int rtp_slab_group_set_fluid_params(rtp_slab_group* g, const rtp_fluid_params* fluid, int nb_jacobi_iters)
{
  if (!g || !fluid || nb_jacobi_iters < 1)
    return RTP_ERR_INVALID;
  for (Slab& sl : g->slabs)
    SG_RTP(g, sl, rtp_set_fluid_params(sl.h, fluid, nb_jacobi_iters));
  g->jacobi = nb_jacobi_iters;
  g->vorticity = fluid->isVorticityConfEnabled != 0;
  return RTP_OK;
}

// x-layer of an initial position: the cell formula of grid.cl:14-24 along x (any consistent assignment works: the first
// step migrates by predicted position anyway); same arithmetic as sharded.split_initial_state
static inline u32 initialLayer(float x, float w, float hcell, u32 rx)
{
  const float c = x < -w ? -w : (x > w ? w : x);
  const float f = floorf((c + w) / hcell);
  const long long l = (long long)f;
  return l < 0 ? 0u : (l >= (long long)rx ? rx - 1u : (u32)l);
}

extern "C" int rtp_slab_group_upload(rtp_slab_group* g, const float* pos_xyzw, const float* vel_xyzw, uint64_t n)
{
  if (!g || (n && (!pos_xyzw || !vel_xyzw)))
    return RTP_ERR_INVALID;
  const int W = (int)g->slabs.size();
  const float hcell = (float)g->box[0] / (float)g->grid[0], w = (float)(g->box[0] / 2.0);
  std::vector<std::vector<float>> P(W), V(W);
  for (uint64_t i = 0; i < n; ++i)
  {
    const u32 layer = initialLayer(pos_xyzw[4 * i], w, hcell, g->grid[0]);
    int r = 0;
    while (r + 1 < W && layer >= g->slabs[r].xhi)
      ++r;
    P[r].insert(P[r].end(), pos_xyzw + 4 * i, pos_xyzw + 4 * i + 4);
    V[r].insert(V[r].end(), vel_xyzw + 4 * i, vel_xyzw + 4 * i + 4);
  }
  for (int r = 0; r < W; ++r)
  {
    Slab& sl = g->slabs[r];
    const size_t k = P[r].size() / 4;
    if (k > sl.A0)
    {
      g->err = "slab capacity exceeded: more particles than rows before the arrival slots";
      return RTP_ERR_INVALID;
    }
    SG_CUDA(g, cudaSetDevice(sl.dev));
    float4 *pos = nullptr, *vel = nullptr;
    int rc = shardBuf(g, sl, RTP_SHARD_BUF_POS, &pos);
    if (rc == RTP_OK)
      rc = shardBuf(g, sl, RTP_SHARD_BUF_VEL, &vel);
    if (rc != RTP_OK)
      return rc;
    fillNoParticleKernel<<<blocksFor(sl.cap, 256), 256, 0, sl.st>>>(pos, sl.cap); // "no particle" everywhere ...
    SG_CUDA(g, cudaMemsetAsync(vel, 0, (size_t)sl.cap * sizeof(float4), sl.st));
    if (k)
    {
      SG_CUDA(g, cudaMemcpyAsync(pos, P[r].data(), k * sizeof(float4), cudaMemcpyHostToDevice, sl.st)); // ... but the slab's particles
      SG_CUDA(g, cudaMemcpyAsync(vel, V[r].data(), k * sizeof(float4), cudaMemcpyHostToDevice, sl.st));
    }
    SG_CUDA(g, cudaMemsetAsync(sl.flags, 0, 2 * sizeof(u64), sl.st));
    SG_CUDA(g, cudaStreamSynchronize(sl.st)); // (the host vectors go away)
    rc = setCounts(g, sl, sl.S, sl.S);
    if (rc != RTP_OK)
      return rc;
  }
  return RTP_OK;
}

// one neighbour-to-neighbour copy: dst on slab `to` (its stream `st`), src on slab `from`
static cudaError_t pull(void* dst, const void* src, size_t bytes, cudaStream_t st) { return bytes ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st) : cudaSuccess; }

struct StageDesc
{
  int stage, it, last;
  int nFields;
  int field[2]; // rtp_shard_buffer_id
  int rowBuf[2]; // RowBuf
  int checkEpoch; // < 0: none
};

static int enqueueStep(rtp_slab_group* g)
{
  std::vector<Slab>& S = g->slabs;
  const int W = (int)S.size();
  auto nb = [&](int r, int s) -> Slab* { return s == 0 ? (S[r].hasL ? &S[r - 1] : nullptr) : (S[r].hasR ? &S[r + 1] : nullptr); };

  // 1. predict; particles whose predicted cell left the slab migrate: their row is cleared here, the neighbour gets them in
  //    its arrival slots
  for (int r = 0; r < W; ++r)
  {
    Slab& sl = S[r];
    SG_CUDA(g, cudaSetDevice(sl.dev));
    int rc = setCounts(g, sl, sl.S, sl.S);
    if (rc != RTP_OK)
      return rc;
    SG_RTP(g, sl, rtp_shard_stage(sl.h, RTP_SHARD_PREDICT, 0, 0));
    if (W == 1)
      continue;
    SG_RTP(g, sl, rtp_shard_classify(sl.h, sl.S, sl.xlo, sl.xhi, sl.mask[0], sl.mask[1]));
    float4* pos = nullptr;
    rc = shardBuf(g, sl, RTP_SHARD_BUF_POS, &pos);
    if (rc != RTP_OK)
      return rc;
    arrivalSlotsFreeKernel<<<blocksFor(2 * sl.Mc, 256), 256, 0, sl.st>>>(pos, sl.A0, 2 * sl.Mc, sl.flags);
    for (int s = 0; s < 2; ++s)
    {
      Slab* q = nb(r, s);
      if (!q)
        continue;
      SG_CUDA(g, cudaStreamWaitEvent(sl.st, q->evMigPulled[1 - s], 0)); // the neighbour has taken the previous message
      rc = compactMask(g, sl, sl.mask[s], sl.S, sl.Mc, sl.idxMig[s], 2ull, true);
      if (rc != RTP_OK)
        return rc;
      SG_RTP(g, sl, rtp_shard_pack(sl.h, RTP_SHARD_BUF_POS, sl.idxMig[s], sl.Mc, sl.migS[s]));
      SG_RTP(g, sl, rtp_shard_pack(sl.h, RTP_SHARD_BUF_VEL, sl.idxMig[s], sl.Mc, sl.migS[s] + (size_t)sl.Mc * 4));
      SG_RTP(g, sl, rtp_shard_clear_rows(sl.h, sl.idxMig[s], sl.Mc));
    }
    SG_CUDA(g, cudaEventRecord(sl.evMigPacked, sl.st));
  }
  if (W > 1)
    for (int r = 0; r < W; ++r)
    {
      Slab& sl = S[r];
      SG_CUDA(g, cudaSetDevice(sl.dev));
      float4 *pos = nullptr, *vel = nullptr;
      int rc = shardBuf(g, sl, RTP_SHARD_BUF_POS, &pos);
      if (rc == RTP_OK)
        rc = shardBuf(g, sl, RTP_SHARD_BUF_VEL, &vel);
      if (rc != RTP_OK)
        return rc;
      for (int s = 0; s < 2; ++s)
      {
        Slab* q = nb(r, s);
        if (!q)
          continue;
        const size_t a = (size_t)sl.A0 + (size_t)s * sl.Mc, bytes = (size_t)sl.Mc * sizeof(float4);
        SG_CUDA(g, cudaStreamWaitEvent(sl.st, q->evMigPacked, 0));
        SG_CUDA(g, pull(pos + a, q->migS[1 - s], bytes, sl.st)); // (the message is padded with "no particle" rows)
        SG_CUDA(g, pull(vel + a, q->migS[1 - s] + (size_t)sl.Mc * 4, bytes, sl.st));
        SG_CUDA(g, cudaEventRecord(sl.evMigPulled[s], sl.st));
      }
      SG_RTP(g, sl, rtp_shard_stage(sl.h, RTP_SHARD_PREDICT_FROM, (int)sl.A0, 0)); // the arrivals need their prediction and cell id
    }

  // 2. halo of predicted positions into the ghost regions, sort everything by cell
  if (W > 1)
  {
    for (int r = 0; r < W; ++r)
    {
      Slab& sl = S[r];
      SG_CUDA(g, cudaSetDevice(sl.dev));
      SG_RTP(g, sl, rtp_shard_classify(sl.h, sl.S, sl.xlo + RTP_SHARD_GHOST_LAYERS, sl.xhi - RTP_SHARD_GHOST_LAYERS, sl.mask[0], sl.mask[1]));
      for (int s = 0; s < 2; ++s)
      {
        Slab* q = nb(r, s);
        if (!q)
          continue;
        SG_CUDA(g, cudaStreamWaitEvent(sl.st, q->evHaloPulled[1 - s], 0));
        const int rc = compactMask(g, sl, sl.mask[s], sl.S, sl.G, sl.src[s], 4ull, false);
        if (rc != RTP_OK)
          return rc;
        SG_RTP(g, sl, rtp_shard_pack(sl.h, RTP_SHARD_BUF_PRED_IN, sl.src[s], sl.G, sl.haloS[s]));
      }
      SG_CUDA(g, cudaEventRecord(sl.evHaloPacked, sl.st));
    }
    for (int r = 0; r < W; ++r)
    {
      Slab& sl = S[r];
      SG_CUDA(g, cudaSetDevice(sl.dev));
      float4 *pos = nullptr, *vel = nullptr, *predIn = nullptr;
      int rc = shardBuf(g, sl, RTP_SHARD_BUF_POS, &pos);
      if (rc == RTP_OK)
        rc = shardBuf(g, sl, RTP_SHARD_BUF_VEL, &vel);
      if (rc == RTP_OK)
        rc = shardBuf(g, sl, RTP_SHARD_BUF_PRED_IN, &predIn);
      if (rc != RTP_OK)
        return rc;
      const size_t bytes = (size_t)sl.G * sizeof(float4);
      for (int s = 0; s < 2; ++s)
      {
        Slab* q = nb(r, s);
        const size_t off = (size_t)sl.S + (size_t)s * sl.G;
        if (q)
        {
          SG_CUDA(g, cudaStreamWaitEvent(sl.st, q->evHaloPacked, 0));
          SG_CUDA(g, pull(predIn + off, q->haloS[1 - s], bytes, sl.st));
          SG_CUDA(g, cudaEventRecord(sl.evHaloPulled[s], sl.st));
          SG_CUDA(g, pull(pos + off, predIn + off, bytes, sl.st));
        }
        else
        {
          SG_CUDA(g, pull(predIn + off, sl.noParticle, bytes, sl.st));
          SG_CUDA(g, pull(pos + off, sl.noParticle, bytes, sl.st));
        }
      }
      SG_CUDA(g, cudaMemsetAsync(vel + sl.S, 0, 2 * bytes, sl.st));
      rc = setCounts(g, sl, sl.S, (u64)sl.S + 2ull * sl.G);
      if (rc != RTP_OK)
        return rc;
      SG_RTP(g, sl, rtp_shard_stage(sl.h, RTP_SHARD_GHOST_KEYS, 0, 0));
    }
  }
  for (int r = 0; r < W; ++r)
  {
    Slab& sl = S[r];
    SG_CUDA(g, cudaSetDevice(sl.dev));
    SG_RTP(g, sl, rtp_shard_stage(sl.h, RTP_SHARD_SORT, 0, 0));
    if (W > 1)
    {
      SG_RTP(g, sl, rtp_shard_inverse_perm(sl.h, sl.inv));
      indexMapsKernel<<<blocksFor(2 * sl.G, 256), 256, 0, sl.st>>>(sl.inv, sl.src[0], sl.src[1], sl.G, sl.S, sl.hasL ? 1 : 0, sl.hasR ? 1 : 0,
          sl.sendAll, sl.recvAll);
      SG_CUDA(g, cudaGetLastError());
    }
  }

  // 3. the solver stages, each followed by the refresh of what it produced. With overlap, a stage sweeps its interior rows
  //    first -- while the refresh of the previous stage is still travelling -- and the rest once it has arrived.
  std::vector<StageDesc> stages;
  for (int it = 0; it < g->jacobi; ++it)
  {
    const int last = it == g->jacobi - 1;
    stages.push_back({ RTP_SHARD_DENSITY_LAMBDA, it, 0, 1, { RTP_SHARD_BUF_LAMBDA, 0 }, { ROWS_W1, 0 }, -1 });
    stages.push_back({ RTP_SHARD_CORRECTION, it, last, last ? 2 : 1, { RTP_SHARD_BUF_PRED_CUR, RTP_SHARD_BUF_VEL_SORTED }, { ROWS_W4, ROWS_W4B }, it + 1 });
  }
  if (g->vorticity)
  {
    stages.push_back({ RTP_SHARD_VORTICITY, g->jacobi, 0, 1, { RTP_SHARD_BUF_VORT_NORM, 0 }, { ROWS_W1, 0 }, -1 });
    stages.push_back({ RTP_SHARD_CONFINEMENT, g->jacobi, 0, 1, { RTP_SHARD_BUF_VEL_CONFINED, 0 }, { ROWS_W4, 0 }, -1 });
    stages.push_back({ RTP_SHARD_XSPH, g->jacobi, 0, 0, { 0, 0 }, { 0, 0 }, -1 });
  }
  for (const StageDesc& sd : stages)
  {
    for (int r = 0; r < W; ++r)
    {
      Slab& sl = S[r];
      if (g->overlap)
      {
        SG_RTP(g, sl, rtp_shard_stage_rows(sl.h, sd.stage, sd.it, sd.last, RTP_ROWS_INTERIOR));
        SG_RTP(g, sl, rtp_shard_exchange_join(sl.h));
        SG_RTP(g, sl, rtp_shard_stage_rows(sl.h, sd.stage, sd.it, sd.last, RTP_ROWS_BOUNDARY));
      }
      else
        SG_RTP(g, sl, rtp_shard_stage(sl.h, sd.stage, sd.it, sd.last));
    }
    if (W == 1 || sd.nFields == 0)
      continue;
    // all fields a stage produced travel in ONE exchange; both faces are packed / unpacked by one launch per field
    for (int r = 0; r < W; ++r)
    {
      Slab& sl = S[r];
      SG_CUDA(g, cudaSetDevice(sl.dev));
      if (g->overlap)
        SG_RTP(g, sl, rtp_shard_exchange_fork(sl.h));
      cudaStream_t xs = g->overlap ? sl.ex : sl.st;
      for (int k = 0; k < sd.nFields; ++k)
      {
        const int b = sd.rowBuf[k];
        for (int s = 0; s < 2; ++s)
          if (Slab* q = nb(r, s))
            SG_CUDA(g, cudaStreamWaitEvent(xs, q->evRowPulled[b][1 - s], 0));
        SG_RTP(g, sl, rtp_shard_pack(sl.h, sd.field[k], sl.sendAll, 2ull * sl.G, sl.rowS[b]));
        SG_CUDA(g, cudaEventRecord(sl.evRowPacked[b], xs));
      }
    }
    for (int r = 0; r < W; ++r)
    {
      Slab& sl = S[r];
      SG_CUDA(g, cudaSetDevice(sl.dev));
      cudaStream_t xs = g->overlap ? sl.ex : sl.st;
      for (int k = 0; k < sd.nFields; ++k)
      {
        const int b = sd.rowBuf[k];
        const size_t w = b == ROWS_W1 ? 1 : 4, face = (size_t)sl.G * w;
        for (int s = 0; s < 2; ++s)
        {
          Slab* q = nb(r, s);
          if (!q)
            continue;
          // my face s receives what the neighbour packed for its face 1 - s
          SG_CUDA(g, cudaStreamWaitEvent(xs, q->evRowPacked[b], 0));
          SG_CUDA(g, pull(sl.rowR[b] + (size_t)s * face, q->rowS[b] + (size_t)(1 - s) * face, face * sizeof(float), xs));
          SG_CUDA(g, cudaEventRecord(sl.evRowPulled[b][s], xs));
        }
        SG_RTP(g, sl, rtp_shard_unpack(sl.h, sd.field[k], sl.recvAll, 2ull * sl.G, sl.rowR[b]));
      }
      if (sd.checkEpoch >= 0)
      {
        // the owner's kernel checks the particles it moves; its ghosts are checked here
        const u32* gi = sl.hasL ? sl.recvAll : sl.recvAll + sl.G;
        const u64 ng = (u64)sl.G * ((sl.hasL ? 1 : 0) + (sl.hasR ? 1 : 0));
        SG_RTP(g, sl, rtp_shard_check_ghosts(sl.h, gi, ng, sd.checkEpoch));
      }
      if (g->overlap)
        SG_RTP(g, sl, rtp_shard_exchange_done(sl.h));
    }
  }

  // 4. particles to the front of the own region, in cell-sorted order; everything else is "no particle"
  for (int r = 0; r < W; ++r)
  {
    Slab& sl = S[r];
    if (W > 1)
      SG_RTP(g, sl, rtp_shard_stage(sl.h, RTP_SHARD_DROP_GHOSTS, 0, 0));
    const int rc = setCounts(g, sl, sl.S, sl.S);
    if (rc != RTP_OK)
      return rc;
  }
  ++g->steps;
  return RTP_OK;
}

extern "C" int rtp_slab_group_step(rtp_slab_group* g, int nsteps)
{
  if (!g || nsteps < 0)
    return RTP_ERR_INVALID;
  for (int k = 0; k < nsteps; ++k)
  {
    const int rc = enqueueStep(g);
    if (rc != RTP_OK)
      return rc;
  }
  return RTP_OK;
}

extern "C" int rtp_slab_group_sync(rtp_slab_group* g)
{
  if (!g)
    return RTP_ERR_INVALID;
  for (Slab& sl : g->slabs)
  {
    SG_CUDA(g, cudaSetDevice(sl.dev));
    SG_CUDA(g, cudaDeviceSynchronize());
  }
  return RTP_OK;
}

extern "C" int rtp_slab_group_check(rtp_slab_group* g, uint64_t* migrated_total)
{
  if (!g)
    return RTP_ERR_INVALID;
  u64 migrated = 0, capFlags = 0;
  int capSlab = -1, rowSlab = -1;
  for (Slab& sl : g->slabs)
  {
    SG_CUDA(g, cudaSetDevice(sl.dev));
    SG_CUDA(g, cudaDeviceSynchronize());
    u64 f[2] = { 0, 0 };
    SG_CUDA(g, cudaMemcpy(f, sl.flags, sizeof f, cudaMemcpyDeviceToHost));
    u32 rowErr = 0;
    if (g->overlap)
    {
      u32* rb = nullptr;
      const int rc = shardBuf(g, sl, RTP_SHARD_BUF_ROW_BOUNDS, &rb);
      if (rc != RTP_OK)
        return rc;
      SG_CUDA(g, cudaMemcpy(&rowErr, rb + 3, sizeof rowErr, cudaMemcpyDeviceToHost));
    }
    if (f[0] && capSlab < 0)
      capSlab = sl.rank, capFlags = f[0];
    if (rowErr && rowSlab < 0)
      rowSlab = sl.rank;
    migrated += f[1];
  }
  if (capSlab >= 0) // (the cause: an overflowing ghost region also outruns the launches sized from it)
  {
    char buf[200];
    snprintf(buf, sizeof buf, "slab %d: capacity exceeded on the device (flags %llu: 1 = arrival slots in use, 2 = migration message, 4 = ghost region)",
        capSlab, capFlags);
    g->err = buf;
    return RTP_ERR_COMM;
  }
  if (rowSlab >= 0)
  {
    g->err = "slab " + std::to_string(rowSlab) + ": a sweep launched by row phase was sized too small for its rows: results since the last check are invalid";
    return RTP_ERR_COMM;
  }
  if (migrated_total)
    *migrated_total = migrated;
  return RTP_OK;
}

extern "C" int64_t rtp_slab_group_download(rtp_slab_group* g, float* pos_xyzw, float* vel_xyzw, uint64_t capacity_rows, uint64_t* per_slab)
{
  if (!g || !pos_xyzw || !vel_xyzw)
    return RTP_ERR_INVALID;
  const int rc = rtp_slab_group_check(g, nullptr);
  if (rc != RTP_OK)
    return rc;
  uint64_t total = 0;
  std::vector<float> P, V;
  for (Slab& sl : g->slabs)
  {
    SG_CUDA(g, cudaSetDevice(sl.dev));
    float4 *pos = nullptr, *vel = nullptr;
    int rc2 = shardBuf(g, sl, RTP_SHARD_BUF_POS, &pos);
    if (rc2 == RTP_OK)
      rc2 = shardBuf(g, sl, RTP_SHARD_BUF_VEL, &vel);
    if (rc2 != RTP_OK)
      return rc2;
    P.resize((size_t)sl.S * 4);
    V.resize((size_t)sl.S * 4);
    SG_CUDA(g, cudaMemcpy(P.data(), pos, (size_t)sl.S * sizeof(float4), cudaMemcpyDeviceToHost));
    SG_CUDA(g, cudaMemcpy(V.data(), vel, (size_t)sl.S * sizeof(float4), cudaMemcpyDeviceToHost));
    uint64_t n = 0; // the particles sit at the front of the own region (after a step: cell-sorted)
    for (size_t i = 0; i < sl.S; ++i)
      if (isfinite(P[4 * i]))
      {
        if (total + n >= capacity_rows)
        {
          g->err = "rtp_slab_group_download: output capacity too small";
          return RTP_ERR_INVALID;
        }
        memcpy(pos_xyzw + 4 * (total + n), &P[4 * i], 16);
        memcpy(vel_xyzw + 4 * (total + n), &V[4 * i], 16);
        ++n;
      }
    if (per_slab)
      per_slab[sl.rank] = n;
    total += n;
  }
  return (int64_t)total;
}
