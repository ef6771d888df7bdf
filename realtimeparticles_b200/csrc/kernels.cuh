// kernels.cuh -- host-side launch wrappers of the model kernels (definitions in grid.cu, boids.cu, fluids.cu).
// Every wrapper enqueues exactly one kernel on the given stream and returns nothing; errors are collected by the
// caller with cudaGetLastError().
#pragma once

#include "rtp_common.cuh"

namespace rtp
{
// margin mask of a step (tilebuild.cuh MarginMask): word w of lane l of warp W at words[(W * wordCap + w) * 32 + l]
struct MarginMaskBuffers
{
  u32* words = nullptr;
  u32* desc = nullptr;
  u32* warpWords = nullptr;
  u32 wordCap = 0;
};

// Buffers of one model instance. "A" buffers hold the canonical state seen by rtp_upload/rtp_download (the
// reference's named buffers); "B" buffers are the cell-sorted working copies the neighbour kernels read. Every
// step ends with the state back in the A buffers, so one captured CUDA graph replays for every step.
struct DeviceState
{
  u32 M = 0, N = 0;
  u32 nOwned = 0xFFFFFFFFu; // slab decomposition: unsorted indices >= nOwned are ghost copies (sweep.cuh validity check)
  // slab decomposition, overlap of the ghost refresh with the sweeps: sorted rows [rowPhaseBounds[0], rowPhaseBounds[1]) are
  // INTERIOR (two cell layers or more from both slab faces: no ghost among their neighbours, not a ghost of anybody).
  // rowPhase 0 = a launch covers every row, 1 = only the CTAs with a row outside the interior, 2 = only the CTAs inside it.
  u32* rowPhaseBounds = nullptr; // { first interior row, one past the last, one past the last row holding a particle, error flag }
  int rowPhase = 0;
  int rowPhaseToEnd = 0; // a BOUNDARY launch also visits the "no particle" rows behind the last particle (the sweep that writes the state back)
  u32 rowPhaseBlocks = 0; // grid of a launch by row phase (an upper bound from the caller's capacities; sweep.cuh maps the blocks)
  // float4[M]
  float4 *posA = nullptr, *posB = nullptr, *velA = nullptr, *velB = nullptr, *velC = nullptr;
  float4 *col = nullptr, *colB = nullptr, *acc = nullptr;
  float4 *pred0 = nullptr, *pred1 = nullptr; // ping-pong predicted positions; predCur points at the official one
  float4 *corrPos = nullptr, *vort = nullptr, *totCorrA = nullptr, *totCorrB = nullptr;
  // float[M]
  float *density = nullptr, *lambda = nullptr, *vortNorm = nullptr;
  float *tempA = nullptr, *tempB = nullptr, *vaporA = nullptr, *vaporB = nullptr, *cloudA = nullptr, *cloudB = nullptr;
  float *buoyA = nullptr, *buoyB = nullptr, *partIdA = nullptr, *partIdB = nullptr, *cloudGen = nullptr;
  float *lapTemp = nullptr, *lambdaTemp = nullptr, *corrTemp = nullptr;
  float* partDetector = nullptr; // float8[C]
  // u32
  u32 *cellID = nullptr, *keysTmp = nullptr, *perm = nullptr, *permTmp = nullptr;
  u32 *cameraDist = nullptr, *cameraPerm = nullptr;
  uint2* table = nullptr; // c_startEndPartID
  u32 *sortCtrl = nullptr, *sortStatus = nullptr;
  // per-step margin lists (sweep.cuh): rows of 4 entries, row r of particle i = ((uint4*)nbrList)[r * nbrStride + i]
  u32 *nbrList = nullptr, *nbrCount = nullptr, *nbrInvalid = nullptr;
  float4* nbrBuildPos = nullptr;
  u32 nbrStride = 0, nbrCap = 0;
  // per-epoch hit lists (sweep.cuh), same row layout
  u32 *hitList = nullptr, *hitCount = nullptr;
  u32 *stragQueue = nullptr, *stragCount = nullptr, *stragCursor = nullptr; // straggler queue of the producer sweeps (sweep.cuh), counters per epoch
  u32 hitCap = 0;
  int tiledBuild = 0; // the list build of a step runs block-cooperatively (tilebuild.cuh)
  MarginMaskBuffers marginMask; // the margin mask of the step (tilebuild.cuh), written by the filter kernel, read by the build sweep
  u32* buildStats = nullptr; // tilebuild.cuh counters: { irregular warps, warps over the word capacity, CTAs over the tile capacity }
};

constexpr u32 SWEEP_BLOCK_ROWS = 128; // rows per thread block of a neighbour sweep (fluids.cu: NB_THREADS, tilebuild.cuh: TB_THREADS)

// how a neighbour sweep treats the per-step neighbour lists
enum NbrMode
{
  NBR_OFF = 0, // plain 27-cell traversal
  NBR_BUILD = 1, // 27-cell traversal, (re)build the lists
  NBR_BUILD_IF_INVALID = 2, // use the lists unless a particle moved too far since the build, then rebuild
  NBR_USE = 3 // use the lists (built or validated earlier in the same position epoch)
};
constexpr int NBR_EPOCHS = 16;
constexpr int NBR_EPOCH_TEMP = 15; // clouds temperature sweeps (on p_pos)

struct BoidsStepParams
{
  rtp_boids_params rules;
  rtp_target_params target;
  float targetPos[4];
  int targetActive;
  int boundary;
  int dim;
  float dt; // 0.1, Boids.cpp:334
};

struct FluidStepParams
{
  rtp_fluid_params f;
  float invArtDenom; // 1 / (h^2 - (artPressureRadius*h)^2)^3  (poly6 coefficient cancels in the ratio)
};

// ---- grid.cu
void launchSelftestMath(u32 lo, u32 hi, unsigned long long* bad, cudaStream_t st);
void launchRowPhaseBounds(const DeviceState& s, u32 scanLo, u32 scanHi, u32 cellLo, u32 cellHi, u32* bounds, cudaStream_t st);
void launchClassifyRows(const DeviceState& s, const GridParams& g, const u32* keys, u32 n, u32 cutLo, u32 cutHi, unsigned char* below,
    unsigned char* above, cudaStream_t st);
void launchClearRows(const DeviceState& s, const GridParams& g, u32* keys, const u32* idx, u32 n, cudaStream_t st);
void launchPackRows(const void* buf, int rowBytes, const u32* idx, u32 n, void* out, cudaStream_t st);
void launchUnpackRows(void* buf, int rowBytes, const u32* idx, u32 n, const void* in, cudaStream_t st);
void launchInversePerm(const DeviceState& s, u32* inv, cudaStream_t st);
void launchGhostDisplacement(const DeviceState& s, const float4* pred, const u32* idx, u32 n, float dmaxSq, u32* invalid, cudaStream_t st);
void launchGhostFlags(const DeviceState& s, u32* keysOut, cudaStream_t st);
void launchCompactGather(const DeviceState& s, const u32* order, u32 n, cudaStream_t st);
void launchResetIds(const DeviceState& s, u32 numCells, cudaStream_t st);
void launchAdjustEndCell(const DeviceState& s, const GridParams& g, cudaStream_t st, u32 cellLo = 0, u32 cellHi = 0xFFFFFFFFu);
void launchFillCameraDist(const DeviceState& s, const float cam[3], u32* keysOut, cudaStream_t st);
void launchCameraGather(const DeviceState& s, int model, const float4* pred, float4* predOut, cudaStream_t st);
void launchGridDetector(const DeviceState& s, const GridParams& g, cudaStream_t st); // reset + fill (2 launches)
void launchFillFluidColor(const DeviceState& s, float restDensity, cudaStream_t st);
void launchFillColorFloat(const DeviceState& s, const float* quantity, float minVal, float maxVal, cudaStream_t st);

// ---- boids.cu
void launchBoidsCellIds(const DeviceState& s, const GridParams& g, u32* keysOut, cudaStream_t st);
void launchBoidsGather(const DeviceState& s, const GridParams& g, cudaStream_t st);
void launchBoidsRules(const DeviceState& s, const GridParams& g, const SphConsts& c, const BoidsStepParams& p, cudaStream_t st);

// ---- fluids.cu (fluids + clouds)
struct SortPlan;
// fusedSort != nullptr: the kernel also builds that sort's histograms (sort.cuh enqueueSortBegin / enqueueSortPasses around it)
void launchFluidPredict(const DeviceState& s, const GridParams& g, const FluidStepParams& p, u32* keysOut, const SortPlan* fusedSort,
    u32* sortCtrl, u32* sortStatus, cudaStream_t st, bool resets = true);
// the block-cooperative filter of a list build: writes the margin mask of positions P (tilebuild.cuh)
void launchMarginMask(const DeviceState& s, int model, const GridParams& g, const SphConsts& c, const float4* P, cudaStream_t st);
void launchFluidGather(const DeviceState& s, const GridParams& g, cudaStream_t st);
// (returns the number of kernels launched: the list build of a step is a filter + a walk)
int launchDensityLambda(const DeviceState& s, int model, const GridParams& g, const SphConsts& c, const FluidStepParams& p,
    const float4* pred, int nbrMode, int epoch, cudaStream_t st);
// last: also integrates velocity (updateVel) and, without vorticity, copies the position (updatePosition)
void launchCorrection(const DeviceState& s, int model, const GridParams& g, const SphConsts& c, const FluidStepParams& p,
    const rtp_cloud_params& cloud, const float4* pred, float4* predOut, bool last, bool writeCorr, int nbrMode, int epoch,
    cudaStream_t st);
void launchVorticity(const DeviceState& s, int model, const GridParams& g, const SphConsts& c, const float4* pred, int nbrMode,
    int epoch, cudaStream_t st);
void launchConfinement(const DeviceState& s, int model, const GridParams& g, const SphConsts& c, const FluidStepParams& p,
    const float4* pred, int nbrMode, int epoch, cudaStream_t st);
void launchXsph(const DeviceState& s, int model, const GridParams& g, const SphConsts& c, const FluidStepParams& p,
    const rtp_cloud_params& cloud, const float4* pred, int nbrMode, int epoch, cudaStream_t st);
void launchCloudsInitFields(const DeviceState& s, const GridParams& g, const rtp_cloud_params& cloud, cudaStream_t st);
void launchCloudsThermoPredict(const DeviceState& s, const GridParams& g, const rtp_cloud_params& cloud, u32* keysOut, cudaStream_t st);
void launchCloudsGather(const DeviceState& s, const GridParams& g, cudaStream_t st);
int launchCloudsLaplacianTemp(const DeviceState& s, const GridParams& g, const SphConsts& c, const rtp_cloud_params& cloud, int nbrMode, cudaStream_t st);
void launchCloudsLambdaTemp(const DeviceState& s, const GridParams& g, const SphConsts& c, const rtp_cloud_params& cloud, int nbrMode, cudaStream_t st);
void launchCloudsCorrectTemp(const DeviceState& s, const GridParams& g, const SphConsts& c, const rtp_cloud_params& cloud, int nbrMode, cudaStream_t st);
void launchCloudsFinish(const DeviceState& s, const GridParams& g, const rtp_cloud_params& cloud, const float4* pred, bool smoothing, cudaStream_t st);

void launchListStats(const DeviceState& s, const GridParams& g, const SphConsts& c, const float4* pred, unsigned long long* out, cudaStream_t st);
} // namespace rtp
