// sort.cuh -- hand-written onesweep LSD radix sort (32-bit keys + 32-bit permutation payload) for sm_100a.
//
// Replaces RadixSort::sort's key phase (physics/utils/RadixSort.cpp:122-161; kernels/radixSort.cl:19-174), whose
// contract is: stable ascending sort, emits perm with keysAfter[i] == keysBefore[perm[i]] (RadixSort.hpp:13-27).
//
// Design (one read of the keys for all histograms, then ONE read + ONE write of key/value per 8-bit pass):
//   sortHistogramKernel : global per-pass digit histograms (shared-memory atomics, one global atomic per bin/CTA);
//                         also zeroes the look-back status words of every pass.
//   onesweepPassKernel  : per tile: warp-level multisplit ranking (__match_any_sync), tile digit counts published
//                         as (flag|count) words, decoupled look-back over previous tiles, shared-memory reorder,
//                         coalesced scatter. Tiles are handed out by an atomic ticket so look-back never waits on
//                         a CTA that has not started.
// Only ceil(key_bits/8) passes run: cell ids of a 30^3 grid need 2 passes where the reference always runs 4.
#pragma once

#include "rtp_common.cuh"

namespace rtp
{
constexpr int SORT_RADIX_BITS = 8;
constexpr int SORT_RADIX = 1 << SORT_RADIX_BITS;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_MAX_PASSES = 4;

constexpr u32 SORT_FLAG_AGGREGATE = 1u << 30;
constexpr u32 SORT_FLAG_PREFIX = 2u << 30;
constexpr u32 SORT_FLAG_MASK = 3u << 30;
constexpr u32 SORT_VALUE_MASK = ~SORT_FLAG_MASK;

// control block layout (u32 words): [hist: SORT_MAX_PASSES * 256][tile tickets: SORT_MAX_PASSES]
constexpr size_t SORT_CTRL_WORDS = SORT_MAX_PASSES * SORT_RADIX + SORT_MAX_PASSES;

struct SortPlan
{
  u32 n = 0;
  int passes = 0;
  int itemsPerThread = 4;
  u32 tiles = 0;
  int shift[SORT_MAX_PASSES] = { 0, 0, 0, 0 };
  int bits[SORT_MAX_PASSES] = { 0, 0, 0, 0 };
};

// digit extraction of every pass, as the kernels take it
struct PassDesc
{
  int shift[SORT_MAX_PASSES];
  u32 mask[SORT_MAX_PASSES];
};

SortPlan makeSortPlan(u32 n, int keyBits);
PassDesc makePassDesc(const SortPlan& plan);
size_t sortStatusWords(const SortPlan& plan); // passes * tiles * 256

// Enqueue the whole sort. keys0/vals0 and keys1/vals1 are ping-pong buffers of n entries; the input keys are in
// (passes even ? keys0 : keys1) so that the sorted keys and the permutation always END in keys0 / vals0.
// Returns the number of kernel launches enqueued (the control-block memset node is not counted).
int enqueueSort(const SortPlan& plan, u32* keys0, u32* vals0, u32* keys1, u32* vals1, u32* ctrl, u32* status,
    cudaStream_t stream);

// The same in two halves, for a kernel that produces the keys and builds the histograms on the way
// (fluidPredictKernel<true>): enqueueSortBegin zeroes the control block (memset node) BEFORE that kernel, which must add
// every key's digits to ctrl (sortHistogramAdd) and zero the status words (sortStatusClear); enqueueSortPasses then runs
// the onesweep passes only.
void enqueueSortBegin(const SortPlan& plan, u32* ctrl, cudaStream_t stream);
int enqueueSortPasses(const SortPlan& plan, u32* keys0, u32* vals0, u32* keys1, u32* vals1, u32* ctrl, u32* status,
    cudaStream_t stream);

#ifdef __CUDACC__
// Block-wide (SORT_THREADS threads, all of them must call): add the digits of this thread's key (if valid) to the global
// per-pass histograms in ctrl. sHist: SORT_MAX_PASSES * SORT_RADIX words of shared memory.
__device__ __forceinline__ void sortHistogramAdd(u32* sHist, u32 key, bool valid, int passes, const PassDesc& desc, u32* __restrict__ ctrl)
{
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < passes * SORT_RADIX; i += SORT_THREADS)
    sHist[i] = 0;
  __syncthreads();
  for (int p = 0; p < passes; ++p)
  {
    const u32 d = (key >> desc.shift[p]) & desc.mask[p];
    const u32 peers = __match_any_sync(0xffffffffu, valid ? d : (SORT_RADIX + lane));
    if (valid && lane == (__ffs(peers) - 1))
      atomicAdd(&sHist[p * SORT_RADIX + d], __popc(peers));
  }
  __syncthreads();
  for (int i = tid; i < passes * SORT_RADIX; i += SORT_THREADS)
    if (sHist[i])
      atomicAdd(&ctrl[i], sHist[i]);
}
// Grid-wide: zero the look-back status words of every pass.
__device__ __forceinline__ void sortStatusClear(u32* __restrict__ status, size_t statusWords)
{
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gstride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = gtid; i < statusWords; i += gstride)
    status[i] = 0;
}
#endif

} // namespace rtp
