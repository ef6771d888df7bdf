// rtp_api.cu -- the C ABI (include/rtp_cuda.h): handle, named buffers, parameter blocks, step orchestration.
// Replaces CL::Context (physics/ocl/Context.cpp), RadixSort (physics/utils/RadixSort.cpp) and the bodies of
// Boids/Fluids/Clouds::update() (physics/ocl/{Boids,Fluids,Clouds}.cpp). No CPU fallback anywhere.
#include "kernels.cuh"
#include "sort.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <random>
#include <string>
#include <vector>

using namespace rtp;

// cuda_gl_interop.h needs <GL/gl.h>, which a headless build box does not have; the one entry point used here only takes the
// VBO name (GLuint = unsigned int), so it is declared by hand. libcudart resolves the driver's GL interop at run time.
extern "C" cudaError_t cudaGraphicsGLRegisterBuffer(struct cudaGraphicsResource** resource, unsigned int buffer, unsigned int flags);

static thread_local std::string g_createError;

struct StageMark
{
  const char* name;
  cudaEvent_t ev;
};

struct rtp_handle
{
  rtp_config cfg;
  cudaStream_t stream = nullptr;
  DeviceState s;
  GridParams g;
  SphConsts c;
  BoidsStepParams bp;
  FluidStepParams fp;
  rtp_cloud_params cp;
  int jacobi = 2;
  int dispField = RTP_F_CLOUD_DENS;
  float dispMin = 1.0f, dispMax = 15.0f;
  SortPlan cellPlan, camPlan;
  float4* predFinal = nullptr;
  float4* shardCur = nullptr; // slab decomposition: prediction buffer the next stage reads
  // ... overlap of a ghost refresh with the sweeps of the interior rows (rtp_shard_set_interior, rtp_shard_exchange_*)
  u32 interiorCellLo = 0, interiorCellHi = 0, maxBoundaryRows = 0;
  u32* rowBounds = nullptr; // device: sorted rows [rowBounds[0], rowBounds[1]) are interior, [2] = rows with a particle, [3] = error
  volatile u32* rowBoundsHost = nullptr; // pinned copy of [0..2], a step or two old (no synchronisation): sizes the grids
  cudaStream_t exchStream = nullptr;
  cudaEvent_t exchFork = nullptr, exchDone = nullptr;
  bool exchActive = false, exchPending = false;
  // the BOUNDARY launches of the sweeps run on a stream of their own (high priority: the exchange waits for them), so that
  // the two launches of a stage share the GPU instead of each paying its tail: bndPre = "the compute stream has reached
  // this stage", bndDone = "the boundary launch of the last stage is through"
  cudaStream_t bndStream = nullptr;
  cudaEvent_t bndPre = nullptr, bndDone = nullptr;
  bool bndPending = false;
  cudaStream_t shardStream() const { return exchActive ? exchStream : stream; } // pack / unpack / check_ghosts
  float nbrMargin = 0.12f; // RTP_NBR_MARGIN (0.10 / 0.12 / 0.15 measured: equal with a cold L2, 4 / 2 / 0 % faster L2-resident)
  bool nbrEnabled = true; // RTP_NBR_LISTS=0 disables the lists (plain 27-cell traversal in every sweep)
  std::vector<void*> allocs;
  std::string err;
  // graph cache for rtp_step_n
  cudaGraphExec_t graphExec = nullptr;
  unsigned graphFlags = 0;
  float graphCam[3] = { 0, 0, 0 };
  bool graphValid = false;
  float4* graphPredFinal = nullptr; // p_predPos buffer at the start of the captured step
  int lastLaunches = 0;
  // OpenGL interop: VBOs the render engine owns, borrowed by name (Model.hpp:67-70). While registered, the VBO IS the
  // canonical buffer of the field: every step maps it and the kernels write straight into it (no copy).
  static const int GL_SLOTS = 3; // p_pos, p_col, c_partDetector
  cudaGraphicsResource* glRes[GL_SLOTS] = { nullptr, nullptr, nullptr };
  void* glOwn[GL_SLOTS] = { nullptr, nullptr, nullptr }; // the library's own allocation, back in place after unregister
  bool glMapped = false;
  // profiling
  bool profiling = false;
  std::vector<StageMark> marks;
  std::vector<std::string> stageNames;
  std::vector<float> stageMs;
};

#define CUDA_TRY(h, expr)                                                                      \
  do                                                                                           \
  {                                                                                            \
    cudaError_t e_ = (expr);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
    {                                                                                          \
      (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e_);                           \
      return RTP_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

static int fail(rtp_handle* h, int code, const char* msg)
{
  if (h)
    h->err = msg;
  return code;
}

// utils/Utils.cpp:24-29 FloatToStr(val, 10) + the OpenCL compiler parsing the -D literal
extern "C" float rtp_baked_constant(float v)
{
  char buf[128];
  snprintf(buf, sizeof buf, "%.10f", (double)v);
  return strtof(buf, nullptr);
}

static void computeConstants(rtp_handle* h)
{
  const rtp_config& cfg = h->cfg;
  GridParams& g = h->g;
  SphConsts& c = h->c;
  for (int k = 0; k < 3; ++k)
  {
    g.absW[k] = rtp_baked_constant((float)cfg.box[k] / 2.0f);
    g.res[k] = (int)cfg.grid[k];
  }
  g.cellSize = rtp_baked_constant((float)cfg.box[0] / (float)cfg.grid[0]);
  g.numCells = cfg.grid[0] * cfg.grid[1] * cfg.grid[2];
  g.maxPartsInCell = cfg.max_parts_in_cell ? cfg.max_parts_in_cell : (cfg.model == RTP_MODEL_BOIDS ? 3000u : 100u);

  // Fluids.cpp:104-119
  const float effectRadius = (float)cfg.box[0] / (float)cfg.grid[0];
  const float PI_F = 3.1415927f;
  c.h = rtp_baked_constant(effectRadius);
  c.h2 = c.h * c.h;
  c.poly6 = rtp_baked_constant(315.0f / (64.0f * PI_F * powf(effectRadius, 9.f)));
  c.spiky = rtp_baked_constant(15.0f / (PI_F * powf(effectRadius, 6.f)));
  c.spikyK = c.spiky * -3.0f;
  c.maxVel = rtp_baked_constant(30.0f);
  c.effectRadiusSq = rtp_baked_constant(1.0f * (float)cfg.box[0] * (float)cfg.box[0] / (float)((size_t)cfg.grid[0] * cfg.grid[0]));
  // (sqrtf(sq) < h) <=> (sq < supportSq): sqrtf is correctly rounded and monotonic
  float x = c.h * c.h;
  while (sqrtf(x) >= c.h)
    x = nextafterf(x, 0.0f);
  while (sqrtf(x) < c.h)
    x = nextafterf(x, INFINITY);
  c.supportSq = x;
  // (sqrtf(sq) <= FLOAT_EPS) <=> (sq <= epsSq)
  x = RTP_FLOAT_EPS * RTP_FLOAT_EPS;
  while (sqrtf(x) > RTP_FLOAT_EPS)
    x = nextafterf(x, 0.0f);
  while (sqrtf(nextafterf(x, INFINITY)) <= RTP_FLOAT_EPS)
    x = nextafterf(x, INFINITY);
  c.epsSq = x;
  // neighbour lists: radius (1 + margin) h, validity bound 0.45 margin h (fluids.cu "Neighbour lists")
  const float lr = (1.0f + h->nbrMargin) * c.h, dm = 0.45f * h->nbrMargin * c.h;
  c.nbrRadiusSq = lr * lr;
  c.nbrDmaxSq = dm * dm;
}

static void updateDerivedFluidParams(rtp_handle* h)
{
  // poly6L(artPressureRadius * EFFECT_RADIUS) without its coefficient (fluids.cl:56)
  const float len = h->fp.f.artPressureRadius * h->c.h;
  const float t = h->c.h * h->c.h - len * len;
  const float den = (len < h->c.h) ? (t * t) * t : 0.0f;
  h->fp.invArtDenom = 1.0f / den;
}

template <typename T>
static cudaError_t devAlloc(rtp_handle* h, T** p, size_t count)
{
  void* q = nullptr;
  const size_t bytes = (count ? count : 1) * sizeof(T);
  cudaError_t e = cudaMalloc(&q, bytes);
  if (e != cudaSuccess)
    return e;
  e = cudaMemsetAsync(q, 0, bytes, h->stream);
  h->allocs.push_back(q);
  *p = (T*)q;
  return e;
}

static int cellKeyBits(const rtp_config& cfg)
{
  // largest key of an active particle: index RES is reachable on every axis (position exactly on the +wall)
  const uint64_t rx = cfg.grid[0], ry = cfg.grid[1], rz = cfg.grid[2];
  uint64_t maxKey = rx * ry * rz + ry * rz + rz;
  int bits = 0;
  while (maxKey)
  {
    ++bits;
    maxKey >>= 1;
  }
  return bits < 1 ? 1 : bits;
}

static void invalidateGraph(rtp_handle* h)
{
  h->graphValid = false;
}

static void makePlans(rtp_handle* h)
{
  h->cellPlan = makeSortPlan(h->s.N, cellKeyBits(h->cfg));
  h->camPlan = makeSortPlan(h->s.N, 20); // keys <= FAR_DIST = 1e6 < 2^20
}

extern "C" int rtp_abi_version(void) { return RTP_ABI_VERSION; }

extern "C" int rtp_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess)
    return 0;
  return n;
}

extern "C" const char* rtp_last_error(const rtp_handle* h) { return h ? h->err.c_str() : g_createError.c_str(); }

extern "C" void rtp_destroy(rtp_handle* h)
{
  if (!h)
    return;
  cudaSetDevice(h->cfg.device);
  if (h->stream)
    cudaStreamSynchronize(h->stream);
  if (h->graphExec)
    cudaGraphExecDestroy(h->graphExec);
  for (int k = 0; k < rtp_handle::GL_SLOTS; ++k)
    if (h->glRes[k])
    {
      cudaGraphicsUnregisterResource(h->glRes[k]);
      h->allocs.push_back(h->glOwn[k]); // (the field pointer is the VBO's: release the library's own buffer too)
      h->glRes[k] = nullptr;
    }
  for (auto& m : h->marks)
    cudaEventDestroy(m.ev);
  if (h->exchStream)
  {
    cudaStreamSynchronize(h->exchStream);
    cudaStreamDestroy(h->exchStream);
    cudaEventDestroy(h->exchFork);
    cudaEventDestroy(h->exchDone);
    cudaStreamSynchronize(h->bndStream);
    cudaStreamDestroy(h->bndStream);
    cudaEventDestroy(h->bndPre);
    cudaEventDestroy(h->bndDone);
  }
  if (h->rowBoundsHost)
    cudaFreeHost((void*)h->rowBoundsHost);
  for (void* p : h->allocs)
    cudaFree(p);
  if (h->stream)
    cudaStreamDestroy(h->stream);
  delete h;
}

extern "C" int rtp_create(const rtp_config* cfg, rtp_handle** out)
{
  if (!cfg || !out)
  {
    g_createError = "rtp_create: null argument";
    return RTP_ERR_INVALID;
  }
  *out = nullptr;
  if (cfg->model < RTP_MODEL_BOIDS || cfg->model > RTP_MODEL_CLOUDS || cfg->max_particles == 0
      || cfg->nb_particles > cfg->max_particles || cfg->max_particles >= (1ull << 30))
  {
    g_createError = "rtp_create: invalid model or particle counts";
    return RTP_ERR_INVALID;
  }
  for (int k = 0; k < 3; ++k)
    if (cfg->box[k] == 0 || cfg->grid[k] == 0)
    {
      g_createError = "rtp_create: box and grid must be non-zero";
      return RTP_ERR_INVALID;
    }
  if ((uint64_t)cfg->grid[0] * cfg->grid[1] * cfg->grid[2] >= (1ull << 30))
  {
    g_createError = "rtp_create: grid too large";
    return RTP_ERR_INVALID;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0 || cfg->device < 0 || cfg->device >= ndev)
  {
    g_createError = std::string("rtp_create: no usable CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "bad ordinal")
        + "); this backend has no CPU fallback";
    return RTP_ERR_CUDA;
  }
  rtp_handle* h = new rtp_handle();
  h->cfg = *cfg;
  h->cfg.dim = cfg->dim == 2 ? 2 : 3;
#define CREATE_TRY(expr)                                                                 \
  do                                                                                     \
  {                                                                                      \
    cudaError_t e2_ = (expr);                                                            \
    if (e2_ != cudaSuccess)                                                              \
    {                                                                                    \
      g_createError = std::string(#expr) + ": " + cudaGetErrorString(e2_);               \
      rtp_destroy(h);                                                                    \
      return RTP_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)
  CREATE_TRY(cudaSetDevice(cfg->device));
  CREATE_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  if (const char* e = getenv("RTP_NBR_MARGIN"))
    h->nbrMargin = fminf(fmaxf((float)atof(e), 0.01f), 1.0f);
  if (const char* e = getenv("RTP_NBR_LISTS"))
    h->nbrEnabled = atoi(e) != 0;
  computeConstants(h);

  // defaults: Boids.hpp:14-26, Fluids.hpp:17-32, Clouds.hpp:15-45 (+ Clouds.cpp:308-311)
  h->bp.rules = rtp_boids_params { 0.5f, 1.6f, 1.6f, 1.45f };
  h->bp.target = rtp_target_params { 2.0f, 1 };
  h->bp.targetPos[0] = h->bp.targetPos[1] = h->bp.targetPos[2] = h->bp.targetPos[3] = 0.0f;
  h->bp.targetActive = 0;
  h->bp.boundary = RTP_BOUNDARY_BOUNCING_WALL;
  h->bp.dim = (int)h->cfg.dim;
  h->bp.dt = 0.1f;
  h->fp.f = rtp_fluid_params { 450.0f, 600.0f, 0.010f, h->cfg.dim, 1, 0.006f, 0.001f, 4, 1, 0.0004f, 0.0001f };
  updateDerivedFluidParams(h);
  h->cp = rtp_cloud_params { h->cfg.dim, 0.01f, 450.0f, 10.0f, 0.10f, 0.0005f, 5.0f, 0.3485f, 0.07f, 1, 600.0f, 0.75f, 1.0f };

  DeviceState& s = h->s;
  const size_t M = cfg->max_particles;
  const size_t C = h->g.numCells;
  s.M = (u32)M;
  s.N = (u32)cfg->nb_particles;
  const int model = cfg->model;
  const bool fluidLike = model != RTP_MODEL_BOIDS;
  CREATE_TRY(devAlloc(h, &s.posA, M));
  CREATE_TRY(devAlloc(h, &s.posB, M));
  CREATE_TRY(devAlloc(h, &s.velA, M));
  CREATE_TRY(devAlloc(h, &s.velB, M));
  CREATE_TRY(devAlloc(h, &s.velC, M));
  CREATE_TRY(devAlloc(h, &s.col, M));
  CREATE_TRY(devAlloc(h, &s.colB, M));
  if (!fluidLike)
    CREATE_TRY(devAlloc(h, &s.acc, M));
  if (fluidLike)
  {
    CREATE_TRY(devAlloc(h, &s.pred0, M));
    CREATE_TRY(devAlloc(h, &s.pred1, M));
    CREATE_TRY(devAlloc(h, &s.corrPos, M));
    CREATE_TRY(devAlloc(h, &s.vort, M));
    CREATE_TRY(devAlloc(h, &s.density, M));
    CREATE_TRY(devAlloc(h, &s.lambda, M));
    CREATE_TRY(devAlloc(h, &s.vortNorm, M));
    h->predFinal = s.pred0;
    // list entries keep 28 bits for the particle index (sweep.cuh)
    if (M > (1u << 28))
      h->nbrEnabled = false;
    // The lists are an optimisation ((nbrCap + hitCap) * 4 B = 1.66 kB per particle of max_particles): when they cannot be
    // allocated the sweeps run the plain 27-cell traversal, which is bit-identical.
#define LIST_TRY(expr)                       \
  do                                         \
  {                                          \
    if (listsOk && (expr) != cudaSuccess)    \
    {                                        \
      (void)cudaGetLastError();              \
      listsOk = false;                       \
    }                                        \
  } while (0)
    if (h->nbrEnabled)
    {
      bool listsOk = true;
      const size_t firstListAlloc = h->allocs.size();
      u32 cap = 256;
      if (const char* e = getenv("RTP_NBR_CAP"))
        cap = (u32)atoi(e);
      cap = (cap + 3u) & ~3u;
      s.nbrCap = cap;
      s.nbrStride = (u32)M;
      LIST_TRY(devAlloc(h, &s.nbrList, (size_t)cap * M));
      LIST_TRY(devAlloc(h, &s.nbrCount, M));
      LIST_TRY(devAlloc(h, &s.nbrBuildPos, M));
      LIST_TRY(devAlloc(h, &s.nbrInvalid, (size_t)2 * NBR_EPOCHS));
      u32 hcap = 160;
      if (const char* e = getenv("RTP_HIT_CAP"))
        hcap = (u32)atoi(e);
      hcap = (hcap + 3u) & ~3u;
      s.hitCap = hcap;
      LIST_TRY(devAlloc(h, &s.hitList, (size_t)hcap * M));
      LIST_TRY(devAlloc(h, &s.hitCount, M));
      LIST_TRY(devAlloc(h, &s.stragQueue, M));
      LIST_TRY(devAlloc(h, &s.stragCount, (size_t)2 * NBR_EPOCHS));
      LIST_TRY(devAlloc(h, &s.stragCursor, (size_t)2 * NBR_EPOCHS));
      LIST_TRY(devAlloc(h, &s.buildStats, (size_t)4));
      // block-cooperative list build (tilebuild.cuh): word descriptors keep 27 bits for the index, the slot logic needs
      // >= 4 cells per axis; RTP_TILED_BUILD=0 selects the per-thread build (bit-identical lists, slower)
      s.tiledBuild = (M <= (1u << 27) && cfg->grid[0] >= 4 && cfg->grid[1] >= 4 && cfg->grid[2] >= 4) ? 1 : 0;
      if (const char* e = getenv("RTP_TILED_BUILD"))
        s.tiledBuild = s.tiledBuild && atoi(e) != 0;
      if (s.tiledBuild)
      {
        // margin mask of the step: ~30 words per warp-lane; a warp that straddles two dense columns needs more
        u32 wcap = 96;
        if (const char* e = getenv("RTP_MASK_WORDS"))
          wcap = (u32)atoi(e) < 1 ? 1u : (u32)atoi(e);
        const size_t warps = ((M + 127) / 128) * 4;
        s.marginMask.wordCap = wcap;
        LIST_TRY(devAlloc(h, &s.marginMask.words, warps * wcap * 32));
        LIST_TRY(devAlloc(h, &s.marginMask.desc, warps * wcap));
        LIST_TRY(devAlloc(h, &s.marginMask.warpWords, warps));
      }
      if (getenv("RTP_TEST_FAIL_LIST_ALLOC")) // (tests: exercise the out-of-memory path)
        listsOk = false;
      if (!listsOk)
      {
        for (size_t k = firstListAlloc; k < h->allocs.size(); ++k)
          cudaFree(h->allocs[k]);
        h->allocs.resize(firstListAlloc);
        s.nbrList = s.nbrCount = s.nbrInvalid = s.hitList = s.hitCount = s.stragQueue = s.stragCount = s.stragCursor = s.buildStats = nullptr;
        s.nbrBuildPos = nullptr;
        s.tiledBuild = 0;
        s.marginMask = MarginMaskBuffers {};
        h->nbrEnabled = false;
      }
    }
#undef LIST_TRY
  }
  if (model == RTP_MODEL_CLOUDS)
  {
    CREATE_TRY(devAlloc(h, &s.totCorrA, M));
    CREATE_TRY(devAlloc(h, &s.totCorrB, M));
    float** fl[] = { &s.tempA, &s.tempB, &s.vaporA, &s.vaporB, &s.cloudA, &s.cloudB, &s.buoyA, &s.buoyB, &s.partIdA, &s.partIdB,
      &s.cloudGen, &s.lapTemp, &s.lambdaTemp, &s.corrTemp };
    for (float** p : fl)
      CREATE_TRY(devAlloc(h, p, M));
  }
  CREATE_TRY(devAlloc(h, &s.partDetector, C * 8));
  CREATE_TRY(devAlloc(h, &s.cellID, M));
  CREATE_TRY(devAlloc(h, &s.keysTmp, M));
  CREATE_TRY(devAlloc(h, &s.perm, M));
  CREATE_TRY(devAlloc(h, &s.permTmp, M));
  CREATE_TRY(devAlloc(h, &s.cameraDist, M));
  CREATE_TRY(devAlloc(h, &s.cameraPerm, M));
  CREATE_TRY(devAlloc(h, &s.table, C));
  CREATE_TRY(devAlloc(h, &s.sortCtrl, SORT_CTRL_WORDS));
  // status words sized for the worst plan over any N <= M
  {
    const SortPlan worstA = makeSortPlan((u32)M, 32);
    size_t words = (size_t)SORT_MAX_PASSES * ((M + SORT_THREADS * 4 - 1) / (SORT_THREADS * 4)) * SORT_RADIX;
    (void)worstA;
    CREATE_TRY(devAlloc(h, &s.sortStatus, words));
  }
  makePlans(h);
  launchResetIds(s, h->g.numCells, h->stream);
  CREATE_TRY(cudaGetLastError());
  CREATE_TRY(cudaStreamSynchronize(h->stream));
#undef CREATE_TRY
  *out = h;
  return RTP_OK;
}

// ------------------------------------------------------------------ OpenGL interop
// Replaces the cl::BufferGL wrapping of the four shared VBOs (Context.cpp:517-541; created by render/Engine.cpp:70-87,
// :300-304) and acquireGLBuffers / releaseGLBuffers around every frame (Context.cpp:710-750).

static int glSlot(int field) { return field == RTP_F_POS ? 0 : (field == RTP_F_COL ? 1 : (field == RTP_F_PART_DETECTOR ? 2 : -1)); }
static void** glFieldPtr(rtp_handle* h, int slot)
{
  return slot == 0 ? (void**)&h->s.posA : (slot == 1 ? (void**)&h->s.col : (void**)&h->s.partDetector);
}
static size_t glFieldBytes(rtp_handle* h, int slot) { return slot == 2 ? 32 * (size_t)h->g.numCells : 16 * (size_t)h->s.M; }
static bool glAny(const rtp_handle* h) { return h->glRes[0] || h->glRes[1] || h->glRes[2]; }

// map every registered VBO on the handle's stream and point the field at it; glUnmap gives it back to OpenGL
static int glMap(rtp_handle* h)
{
  if (!glAny(h) || h->glMapped)
    return RTP_OK;
  cudaGraphicsResource* res[rtp_handle::GL_SLOTS];
  int n = 0;
  for (int k = 0; k < rtp_handle::GL_SLOTS; ++k)
    if (h->glRes[k])
      res[n++] = h->glRes[k];
  CUDA_TRY(h, cudaGraphicsMapResources(n, res, h->stream));
  h->glMapped = true;
  for (int k = 0; k < rtp_handle::GL_SLOTS; ++k)
    if (h->glRes[k])
    {
      void* p = nullptr;
      size_t bytes = 0;
      CUDA_TRY(h, cudaGraphicsResourceGetMappedPointer(&p, &bytes, h->glRes[k]));
      if (bytes < glFieldBytes(h, k))
        return fail(h, RTP_ERR_INVALID, "registered VBO is smaller than the field");
      *glFieldPtr(h, k) = p;
    }
  return RTP_OK;
}
static int glUnmap(rtp_handle* h)
{
  if (!h->glMapped)
    return RTP_OK;
  cudaGraphicsResource* res[rtp_handle::GL_SLOTS];
  int n = 0;
  for (int k = 0; k < rtp_handle::GL_SLOTS; ++k)
    if (h->glRes[k])
    {
      res[n++] = h->glRes[k];
      *glFieldPtr(h, k) = nullptr; // nobody may touch the VBO while OpenGL has it
    }
  h->glMapped = false;
  CUDA_TRY(h, cudaGraphicsUnmapResources(n, res, h->stream));
  return RTP_OK;
}

// ------------------------------------------------------------------ buffers

static int fieldInfo(rtp_handle* h, int field, void** ptr, size_t* bytes)
{
  DeviceState& s = h->s;
  const size_t M = s.M, C = h->g.numCells;
  void* p = nullptr;
  size_t b = 0;
  switch (field)
  {
  case RTP_F_POS: p = s.posA; b = 16 * M; break;
  case RTP_F_COL: p = s.col; b = 16 * M; break;
  case RTP_F_VEL: p = s.velA; b = 16 * M; break;
  case RTP_F_ACC: p = s.acc; b = 16 * M; break;
  case RTP_F_PRED_POS: p = h->predFinal; b = 16 * M; break;
  case RTP_F_CORR_POS: p = s.corrPos; b = 16 * M; break;
  case RTP_F_VORT: p = s.vort; b = 16 * M; break;
  case RTP_F_TOT_CORR_POS: p = s.totCorrA; b = 16 * M; break;
  case RTP_F_DENSITY: p = s.density; b = 4 * M; break;
  case RTP_F_CONST_FACTOR: p = s.lambda; b = 4 * M; break;
  case RTP_F_TEMP: p = s.tempA; b = 4 * M; break;
  case RTP_F_VAPOR_DENS: p = s.vaporA; b = 4 * M; break;
  case RTP_F_CLOUD_DENS: p = s.cloudA; b = 4 * M; break;
  case RTP_F_BUOYANCY: p = s.buoyA; b = 4 * M; break;
  case RTP_F_CLOUD_GEN: p = s.cloudGen; b = 4 * M; break;
  case RTP_F_PART_ID: p = s.partIdA; b = 4 * M; break;
  case RTP_F_LAPLACIAN_TEMP: p = s.lapTemp; b = 4 * M; break;
  case RTP_F_CONST_FACTOR_TEMP: p = s.lambdaTemp; b = 4 * M; break;
  case RTP_F_CORR_TEMP: p = s.corrTemp; b = 4 * M; break;
  case RTP_F_CELL_ID: p = s.cellID; b = 4 * M; break;
  case RTP_F_CAMERA_DIST: p = s.cameraDist; b = 4 * M; break;
  case RTP_F_START_END_CELL: p = s.table; b = 8 * C; break;
  case RTP_F_PERM: p = s.perm; b = 4 * M; break;
  case RTP_F_CAMERA_PERM: p = s.cameraPerm; b = 4 * M; break;
  case RTP_F_PART_DETECTOR: p = s.partDetector; b = 32 * C; break;
  default: return fail(h, RTP_ERR_INVALID, "unknown field id");
  }
  if (!p && !(glSlot(field) >= 0 && h->glRes[glSlot(field)])) // (a field that lives in a VBO has no pointer while unmapped)
    return fail(h, RTP_ERR_STATE, "field does not exist for this model");
  *ptr = p;
  *bytes = b;
  return RTP_OK;
}

extern "C" int rtp_field_bytes(const rtp_handle* h, int field, size_t* bytes)
{
  if (!h || !bytes)
    return RTP_ERR_INVALID;
  void* p;
  return fieldInfo(const_cast<rtp_handle*>(h), field, &p, bytes);
}

static int uploadField(rtp_handle* h, int field, const void* host, size_t bytes, bool blocking)
{
  if (!h || !host)
    return RTP_ERR_INVALID;
  void* p;
  size_t b;
  const int rc = fieldInfo(h, field, &p, &b);
  if (rc != RTP_OK)
    return rc;
  if (bytes != b)
    return fail(h, RTP_ERR_INVALID, "rtp_upload: size mismatch");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  const bool shared = glSlot(field) >= 0 && h->glRes[glSlot(field)];
  if (shared)
  {
    const int mrc = glMap(h);
    if (mrc != RTP_OK)
      return mrc;
    p = *glFieldPtr(h, glSlot(field));
  }
  CUDA_TRY(h, cudaMemcpyAsync(p, host, bytes, cudaMemcpyHostToDevice, h->stream));
  if (shared)
  {
    const int urc = glUnmap(h);
    if (urc != RTP_OK)
      return urc;
  }
  if (blocking)
    CUDA_TRY(h, cudaStreamSynchronize(h->stream)); // blocking, like the reference's CL_TRUE writes (Context.cpp:368)
  return RTP_OK;
}

extern "C" int rtp_upload(rtp_handle* h, int field, const void* host, size_t bytes) { return uploadField(h, field, host, bytes, true); }
extern "C" int rtp_upload_async(rtp_handle* h, int field, const void* host, size_t bytes) { return uploadField(h, field, host, bytes, false); }

static int downloadField(rtp_handle* h, int field, void* host, size_t bytes, bool blocking)
{
  if (!h || !host)
    return RTP_ERR_INVALID;
  void* p;
  size_t b;
  const int rc = fieldInfo(h, field, &p, &b);
  if (rc != RTP_OK)
    return rc;
  if (bytes != b)
    return fail(h, RTP_ERR_INVALID, "rtp_download: size mismatch");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  const bool shared = glSlot(field) >= 0 && h->glRes[glSlot(field)];
  if (shared)
  {
    const int mrc = glMap(h);
    if (mrc != RTP_OK)
      return mrc;
    p = *glFieldPtr(h, glSlot(field));
  }
  CUDA_TRY(h, cudaMemcpyAsync(host, p, bytes, cudaMemcpyDeviceToHost, h->stream));
  if (shared)
  {
    const int urc = glUnmap(h);
    if (urc != RTP_OK)
      return urc;
  }
  if (blocking)
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return RTP_OK;
}

extern "C" int rtp_download(rtp_handle* h, int field, void* host, size_t bytes) { return downloadField(h, field, host, bytes, true); }
extern "C" int rtp_download_async(rtp_handle* h, int field, void* host, size_t bytes) { return downloadField(h, field, host, bytes, false); }

extern "C" int rtp_device_ptr(rtp_handle* h, int field, void** dptr)
{
  if (!h || !dptr)
    return RTP_ERR_INVALID;
  size_t b;
  return fieldInfo(h, field, dptr, &b);
}

// ------------------------------------------------------------------ parameters

extern "C" int rtp_set_boids_params(rtp_handle* h, const rtp_boids_params* rules, const rtp_target_params* target,
    const float target_pos[4], int target_active)
{
  if (!h)
    return RTP_ERR_INVALID;
  if (h->cfg.model != RTP_MODEL_BOIDS)
    return fail(h, RTP_ERR_STATE, "not a boids model");
  if (rules)
    h->bp.rules = *rules;
  if (target)
    h->bp.target = *target;
  if (target_pos)
    memcpy(h->bp.targetPos, target_pos, sizeof h->bp.targetPos);
  h->bp.targetActive = target_active ? 1 : 0;
  invalidateGraph(h);
  return RTP_OK;
}

extern "C" int rtp_set_fluid_params(rtp_handle* h, const rtp_fluid_params* fluid, int nb_jacobi_iters)
{
  if (!h)
    return RTP_ERR_INVALID;
  if (h->cfg.model == RTP_MODEL_BOIDS)
    return fail(h, RTP_ERR_STATE, "not a fluids/clouds model");
  if (fluid)
  {
    // the ranges of the reference's UI (Fluids.cpp:52-74): the exponent is an unrolled product of 1..6 factors, and a
    // radius factor >= 1 puts the reference distance outside the kernel support (division by poly6 = 0)
    if (fluid->isArtPressureEnabled
        && (fluid->artPressureExp < 1u || fluid->artPressureExp > 6u || !(fluid->artPressureRadius > 0.0f && fluid->artPressureRadius < 1.0f)))
      return fail(h, RTP_ERR_INVALID, "artificial pressure: exponent must be in [1, 6] and the radius factor in (0, 1)");
    h->fp.f = *fluid;
  }
  if (nb_jacobi_iters > 0)
    h->jacobi = nb_jacobi_iters;
  updateDerivedFluidParams(h);
  invalidateGraph(h);
  return RTP_OK;
}

extern "C" int rtp_set_cloud_params(rtp_handle* h, const rtp_cloud_params* cloud)
{
  if (!h)
    return RTP_ERR_INVALID;
  if (h->cfg.model != RTP_MODEL_CLOUDS)
    return fail(h, RTP_ERR_STATE, "not a clouds model");
  if (cloud)
    h->cp = *cloud;
  invalidateGraph(h);
  return RTP_OK;
}

extern "C" int rtp_set_boundary(rtp_handle* h, int boundary)
{
  if (!h)
    return RTP_ERR_INVALID;
  if (boundary != RTP_BOUNDARY_BOUNCING_WALL && boundary != RTP_BOUNDARY_CYCLIC_WALL)
    return fail(h, RTP_ERR_INVALID, "unknown boundary");
  h->bp.boundary = boundary;
  invalidateGraph(h);
  return RTP_OK;
}

extern "C" int rtp_set_nb_particles(rtp_handle* h, uint64_t n)
{
  if (!h)
    return RTP_ERR_INVALID;
  if (n > h->s.M)
    return fail(h, RTP_ERR_INVALID, "nb_particles > max_particles");
  h->s.N = (u32)n;
  h->cfg.nb_particles = n;
  makePlans(h);
  invalidateGraph(h);
  return RTP_OK;
}

extern "C" int rtp_set_dimension(rtp_handle* h, int dim)
{
  if (!h)
    return RTP_ERR_INVALID;
  h->cfg.dim = dim == 2 ? 2 : 3;
  h->bp.dim = (int)h->cfg.dim;
  h->fp.f.dim = h->cfg.dim;
  h->cp.dim = h->cfg.dim;
  invalidateGraph(h);
  return RTP_OK;
}

extern "C" int rtp_set_displayed_quantity(rtp_handle* h, int field, float min_val, float max_val)
{
  if (!h)
    return RTP_ERR_INVALID;
  void* p;
  size_t b;
  const int rc = fieldInfo(h, field, &p, &b);
  if (rc != RTP_OK)
    return rc;
  if (b != 4 * (size_t)h->s.M)
    return fail(h, RTP_ERR_INVALID, "displayed quantity must be a float[M] field");
  h->dispField = field;
  h->dispMin = min_val;
  h->dispMax = max_val;
  invalidateGraph(h);
  return RTP_OK;
}

extern "C" int rtp_reset_ids(rtp_handle* h)
{
  if (!h)
    return RTP_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  launchResetIds(h->s, h->g.numCells, h->stream);
  CUDA_TRY(h, cudaGetLastError());
  return RTP_OK;
}

extern "C" int rtp_init_clouds_fields(rtp_handle* h)
{
  if (!h)
    return RTP_ERR_INVALID;
  if (h->cfg.model != RTP_MODEL_CLOUDS)
    return fail(h, RTP_ERR_STATE, "not a clouds model");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  launchCloudsInitFields(h->s, h->g, h->cp, h->stream);
  CUDA_TRY(h, cudaGetLastError());
  return RTP_OK;
}

// ------------------------------------------------------------------ OpenGL interop: registration

extern "C" int rtp_register_gl(rtp_handle* h, int field, unsigned int vbo)
{
  if (!h)
    return RTP_ERR_INVALID;
  const int slot = glSlot(field);
  if (slot < 0)
    return fail(h, RTP_ERR_INVALID, "rtp_register_gl: only p_pos, p_col and c_partDetector are shared with OpenGL");
  if (h->glRes[slot])
    return fail(h, RTP_ERR_STATE, "rtp_register_gl: field already registered");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  cudaGraphicsResource* res = nullptr;
  const cudaError_t e = cudaGraphicsGLRegisterBuffer(&res, vbo, cudaGraphicsRegisterFlagsNone);
  if (e != cudaSuccess)
  {
    (void)cudaGetLastError();
    h->err = std::string("cudaGraphicsGLRegisterBuffer: ") + cudaGetErrorString(e) + " (is an OpenGL context current on this thread?)";
    return RTP_ERR_CUDA;
  }
  // the current contents of the field move into the VBO; from now on the VBO is the field
  void* own = *glFieldPtr(h, slot);
  h->glRes[slot] = res;
  h->glOwn[slot] = own;
  int rc = glMap(h);
  if (rc == RTP_OK)
  {
    cudaMemcpyAsync(*glFieldPtr(h, slot), own, glFieldBytes(h, slot), cudaMemcpyDeviceToDevice, h->stream);
    rc = glUnmap(h);
  }
  if (rc != RTP_OK)
  {
    cudaGraphicsUnregisterResource(res);
    h->glRes[slot] = nullptr;
    *glFieldPtr(h, slot) = own;
    return rc;
  }
  invalidateGraph(h);
  return RTP_OK;
}

extern "C" int rtp_unregister_gl(rtp_handle* h, int field)
{
  if (!h)
    return RTP_ERR_INVALID;
  const int slot = glSlot(field);
  if (slot < 0 || !h->glRes[slot])
    return fail(h, RTP_ERR_STATE, "rtp_unregister_gl: field is not registered");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  // the VBO's contents come back into the library's own buffer
  int rc = glMap(h);
  if (rc == RTP_OK)
  {
    cudaMemcpyAsync(h->glOwn[slot], *glFieldPtr(h, slot), glFieldBytes(h, slot), cudaMemcpyDeviceToDevice, h->stream);
    rc = glUnmap(h);
  }
  cudaStreamSynchronize(h->stream);
  cudaGraphicsUnregisterResource(h->glRes[slot]);
  h->glRes[slot] = nullptr;
  *glFieldPtr(h, slot) = h->glOwn[slot];
  h->glOwn[slot] = nullptr;
  invalidateGraph(h);
  return rc;
}

// ------------------------------------------------------------------ the step

struct StepRecorder
{
  rtp_handle* h;
  bool on;
  size_t used = 0;
  void mark(const char* name)
  {
    if (!on)
      return;
    if (used == h->marks.size())
    {
      StageMark m { name, nullptr };
      cudaEventCreate(&m.ev);
      h->marks.push_back(m);
    }
    h->marks[used].name = name;
    cudaEventRecord(h->marks[used].ev, h->stream);
    ++used;
  }
};

static int enqueueCameraSort(rtp_handle* h, const float cam[3])
{
  DeviceState& s = h->s;
  cudaStream_t st = h->stream;
  const int model = h->cfg.model;
  int launches = 0;
  if (!s.N)
    return 0;
  u32* keysIn = (h->camPlan.passes % 2 == 0) ? s.cameraDist : s.keysTmp;
  launchFillCameraDist(s, cam, keysIn, st);
  ++launches;
  launches += enqueueSort(h->camPlan, s.cameraDist, s.cameraPerm, s.keysTmp, s.permTmp, s.sortCtrl, s.sortStatus, st);
  float4* predScratch = (h->predFinal == s.pred0) ? s.pred1 : s.pred0;
  launchCameraGather(s, model, h->predFinal, predScratch, st);
  ++launches;
  const size_t n4 = (size_t)s.N * sizeof(float4), n1 = (size_t)s.N * sizeof(float);
  // copy nodes are not counted as kernel launches
  cudaMemcpyAsync(s.posA, s.posB, n4, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(s.col, s.colB, n4, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(s.velA, s.velB, n4, cudaMemcpyDeviceToDevice, st);
  if (model == RTP_MODEL_BOIDS)
  {
    cudaMemcpyAsync(s.acc, s.velC, n4, cudaMemcpyDeviceToDevice, st);
  }
  else
  {
    cudaMemcpyAsync(h->predFinal, predScratch, n4, cudaMemcpyDeviceToDevice, st);
    if (model == RTP_MODEL_CLOUDS)
    {
      cudaMemcpyAsync(s.tempA, s.tempB, n1, cudaMemcpyDeviceToDevice, st);
      cudaMemcpyAsync(s.buoyA, s.buoyB, n1, cudaMemcpyDeviceToDevice, st);
      cudaMemcpyAsync(s.vaporA, s.vaporB, n1, cudaMemcpyDeviceToDevice, st);
      cudaMemcpyAsync(s.cloudA, s.cloudB, n1, cudaMemcpyDeviceToDevice, st);
      cudaMemcpyAsync(s.partIdA, s.partIdB, n1, cudaMemcpyDeviceToDevice, st);
    }
  }
  return launches;
}

// Enqueue one update() on the handle's stream; returns the number of launches (kernels + memset/memcpy nodes).
static int enqueueStep(rtp_handle* h, unsigned flags, const float cam[3], bool profile)
{
  DeviceState& s = h->s;
  const GridParams& g = h->g;
  const SphConsts& c = h->c;
  cudaStream_t st = h->stream;
  const int model = h->cfg.model;
  const bool debug = (flags & RTP_STEP_DEBUG_FIELDS) != 0;
  int launches = 0;
  StepRecorder rec { h, profile };
  rec.mark("begin");

  if ((flags & RTP_STEP_PHYSICS) && !s.N)
  {
    // no particle: the step still runs resetStartEndCell over the cells (Fluids.cpp:419)
    launchBoidsCellIds(s, g, s.keysTmp, st);
    ++launches;
  }
  if ((flags & RTP_STEP_PHYSICS) && s.N)
  {
    u32* keysIn = (h->cellPlan.passes % 2 == 0) ? s.cellID : s.keysTmp;
    if (model == RTP_MODEL_BOIDS)
    {
      launchBoidsCellIds(s, g, keysIn, st);
      ++launches;
      rec.mark("fillCellIDs+resetStartEndCell");
      launches += enqueueSort(h->cellPlan, s.cellID, s.perm, s.keysTmp, s.permTmp, s.sortCtrl, s.sortStatus, st);
      rec.mark("radixSort(onesweep)");
      launchBoidsGather(s, g, st);
      rec.mark("gather+cellTable");
      launchAdjustEndCell(s, g, st);
      launches += 2;
      rec.mark("adjustEndCell");
      launchBoidsRules(s, g, c, h->bp, st);
      ++launches;
      rec.mark("boidsRules+updateVel+updatePos");
    }
    else
    {
      const bool clouds = model == RTP_MODEL_CLOUDS;
      const bool lists = s.nbrList != nullptr && h->jacobi + 1 < NBR_EPOCH_TEMP;
      if (clouds)
      {
        launchCloudsThermoPredict(s, g, h->cp, keysIn, st);
        rec.mark("thermo+predict+boundary+fillCellIDs");
        launches += 1 + enqueueSort(h->cellPlan, s.cellID, s.perm, s.keysTmp, s.permTmp, s.sortCtrl, s.sortStatus, st);
      }
      else
      {
        // the predict kernel produces the keys: it builds the sort's digit histograms on the way
        enqueueSortBegin(h->cellPlan, s.sortCtrl, st);
        launchFluidPredict(s, g, h->fp, keysIn, &h->cellPlan, s.sortCtrl, s.sortStatus, st);
        rec.mark("predictPosition+fillCellIDs");
        launches += 1 + enqueueSortPasses(h->cellPlan, s.cellID, s.perm, s.keysTmp, s.permTmp, s.sortCtrl, s.sortStatus, st);
      }
      rec.mark("radixSort(onesweep)");
      if (clouds)
        launchCloudsGather(s, g, st);
      else
        launchFluidGather(s, g, st);
      rec.mark("gather+cellTable");
      launchAdjustEndCell(s, g, st);
      launches += 2;
      rec.mark("adjustEndCell");
      if (clouds && h->cp.isTempSmoothingEnabled)
      {
        launches += launchCloudsLaplacianTemp(s, g, c, h->cp, lists ? NBR_BUILD : NBR_OFF, st);
        rec.mark("laplacianTemp");
        launchCloudsLambdaTemp(s, g, c, h->cp, lists ? NBR_USE : NBR_OFF, st);
        rec.mark("lambdaTemp");
        launchCloudsCorrectTemp(s, g, c, h->cp, lists ? NBR_USE : NBR_OFF, st);
        rec.mark("correctTemp");
        launches += 2;
      }
      float4* cur = s.pred1;
      float4* nxt = s.pred0;
      for (int it = 0; it < h->jacobi; ++it)
      {
        const bool last = it == h->jacobi - 1;
        launches += launchDensityLambda(s, model, g, c, h->fp, cur, !lists ? NBR_OFF : (it == 0 ? NBR_BUILD : NBR_BUILD_IF_INVALID), it, st);
        rec.mark("densityLambda");
        launchCorrection(s, model, g, c, h->fp, h->cp, cur, nxt, last, debug, lists ? NBR_USE : NBR_OFF, it, st);
        rec.mark("correction");
        launches += 1;
        float4* t = cur;
        cur = nxt;
        nxt = t;
      }
      h->predFinal = cur;
      if (h->fp.f.isVorticityConfEnabled)
      {
        launchVorticity(s, model, g, c, cur, lists ? NBR_BUILD_IF_INVALID : NBR_OFF, h->jacobi, st);
        rec.mark("vorticity");
        launchConfinement(s, model, g, c, h->fp, cur, lists ? NBR_USE : NBR_OFF, h->jacobi, st);
        rec.mark("confinement");
        launchXsph(s, model, g, c, h->fp, h->cp, cur, lists ? NBR_USE : NBR_OFF, h->jacobi, st);
        rec.mark("xsph");
        launches += 3;
      }
      if (clouds)
      {
        launchCloudsFinish(s, g, h->cp, cur, !h->fp.f.isVorticityConfEnabled, st);
        ++launches;
        rec.mark("updatePosition");
      }
    }
    if (flags & RTP_STEP_RENDER_AUX)
    {
      launchGridDetector(s, g, st);
      launches += 2;
      if (model == RTP_MODEL_FLUIDS)
      {
        launchFillFluidColor(s, h->fp.f.restDensity, st);
        ++launches;
      }
    }
  }
  if ((flags & RTP_STEP_RENDER_AUX) && model == RTP_MODEL_CLOUDS && s.N)
  {
    void* q;
    size_t b;
    if (fieldInfo(h, h->dispField, &q, &b) == RTP_OK)
    {
      launchFillColorFloat(s, (const float*)q, h->dispMin, h->dispMax, st);
      ++launches;
    }
  }
  if (flags & RTP_STEP_RENDER_AUX)
    rec.mark("renderAux");
  if (flags & RTP_STEP_CAMERA_SORT)
  {
    launches += enqueueCameraSort(h, cam);
    rec.mark("cameraSort");
  }
  if (profile)
  {
    h->stageNames.clear();
    h->stageMs.clear();
    cudaStreamSynchronize(st);
    for (size_t i = 1; i < rec.used; ++i)
    {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, h->marks[i - 1].ev, h->marks[i].ev);
      h->stageNames.push_back(h->marks[i].name);
      h->stageMs.push_back(ms);
    }
  }
  return launches;
}

static const float kDefaultCam[3] = { 32.0f, -1.2f, 0.0f }; // render/Camera.cpp:11

extern "C" int rtp_step(rtp_handle* h, unsigned flags, const float camera_pos[3])
{
  if (!h)
    return RTP_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  // acquireGLBuffers / releaseGLBuffers of the reference (Context.cpp:710-750): the kernels of the step write the mapped
  // VBOs directly
  int rc = glMap(h);
  if (rc != RTP_OK)
    return rc;
  h->lastLaunches = enqueueStep(h, flags, camera_pos ? camera_pos : kDefaultCam, h->profiling);
  rc = glUnmap(h);
  if (rc != RTP_OK)
    return rc;
  CUDA_TRY(h, cudaGetLastError());
  return RTP_OK;
}

extern "C" int rtp_step_n(rtp_handle* h, unsigned flags, const float camera_pos[3], int n)
{
  if (!h || n < 0)
    return RTP_ERR_INVALID;
  if (n == 0)
    return RTP_OK;
  const float* cam = camera_pos ? camera_pos : kDefaultCam;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  if (glAny(h))
  {
    // a mapped VBO may come back at another address: no graph replay, one map / step / unmap per frame
    for (int i = 0; i < n; ++i)
    {
      const int rc = rtp_step(h, flags, cam);
      if (rc != RTP_OK)
        return rc;
    }
    return RTP_OK;
  }
  // the captured graph bakes in the flags, the camera and the buffer p_predPos currently lives in (the camera sort reads it)
  if (!h->graphValid || h->graphFlags != flags || memcmp(h->graphCam, cam, sizeof h->graphCam) != 0 || h->graphPredFinal != h->predFinal)
  {
    if (h->graphExec)
    {
      cudaGraphExecDestroy(h->graphExec);
      h->graphExec = nullptr;
    }
    h->graphValid = false;
    cudaGraph_t graph = nullptr;
    float4* const predBefore = h->predFinal;
    CUDA_TRY(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    h->lastLaunches = enqueueStep(h, flags, cam, false);
    cudaError_t e = cudaStreamEndCapture(h->stream, &graph); // always ends the capture, also after a failed launch
    if (e == cudaSuccess)
      e = cudaGraphInstantiate(&h->graphExec, graph, 0);
    if (graph)
      cudaGraphDestroy(graph);
    if (e != cudaSuccess)
    {
      (void)cudaGetLastError();
      h->graphExec = nullptr;
      h->predFinal = predBefore;
      h->err = std::string("rtp_step_n: graph capture failed: ") + cudaGetErrorString(e);
      return RTP_ERR_CUDA;
    }
    h->graphFlags = flags;
    memcpy(h->graphCam, cam, sizeof h->graphCam);
    h->graphPredFinal = predBefore;
    h->graphValid = true;
  }
  for (int i = 0; i < n; ++i)
    CUDA_TRY(h, cudaGraphLaunch(h->graphExec, h->stream));
  return RTP_OK;
}

extern "C" int rtp_get_stream(rtp_handle* h, void** stream)
{
  if (!h || !stream)
    return RTP_ERR_INVALID;
  *stream = (void*)h->stream;
  return RTP_OK;
}

extern "C" int rtp_sync(rtp_handle* h)
{
  if (!h)
    return RTP_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  CUDA_TRY(h, cudaGetLastError());
  return RTP_OK;
}

// ------------------------------------------------------------------ slab decomposition (stage-wise step)

extern "C" int rtp_shard_set_owned(rtp_handle* h, uint64_t n_owned)
{
  if (!h)
    return RTP_ERR_INVALID;
  h->s.nOwned = n_owned >= 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)n_owned;
  invalidateGraph(h);
  return RTP_OK;
}

extern "C" float rtp_shard_list_dmax_sq(const rtp_handle* h) { return h ? h->c.nbrDmaxSq : 0.0f; }

static int ensureExchangeStream(rtp_handle* h);

extern "C" int rtp_shard_stage(rtp_handle* h, int stage, int iter, int last) { return rtp_shard_stage_rows(h, stage, iter, last, RTP_ROWS_ALL); }

extern "C" int rtp_shard_stage_rows(rtp_handle* h, int stage, int iter, int last, int rows)
{
  if (!h || rows < RTP_ROWS_ALL || rows > RTP_ROWS_INTERIOR)
    return RTP_ERR_INVALID;
  if (h->cfg.model != RTP_MODEL_FLUIDS)
    return fail(h, RTP_ERR_STATE, "slab decomposition is implemented for the fluids model");
  if (rows != RTP_ROWS_ALL && (stage < RTP_SHARD_DENSITY_LAMBDA || stage > RTP_SHARD_XSPH))
    return fail(h, RTP_ERR_INVALID, "only the neighbour sweeps run by row phase");
  if (rows != RTP_ROWS_ALL && !h->rowBounds)
    return fail(h, RTP_ERR_STATE, "rtp_shard_set_interior first");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  // a sweep by row phase: INTERIOR first, then BOUNDARY -- the call that completes the stage moves on the ping-pong buffer
  DeviceState sPhase = h->s;
  sPhase.rowPhase = rows;
  // grids: every block could be interior; the boundary blocks hold at most maxBoundaryRows rows + one straddling block per side
  const u32 allBlocks = (h->s.N + SWEEP_BLOCK_ROWS - 1) / SWEEP_BLOCK_ROWS;
  // (the sweep that writes p_pos / p_vel back writes the "no particle" rows too: they held particles in the unsorted layout)
  sPhase.rowPhaseToEnd = rows == RTP_ROWS_BOUNDARY
      && (stage == RTP_SHARD_XSPH || (stage == RTP_SHARD_CORRECTION && last && !h->fp.f.isVorticityConfEnabled));
  sPhase.rowPhaseBlocks = rows == RTP_ROWS_BOUNDARY && !sPhase.rowPhaseToEnd
      ? min(allBlocks, (h->maxBoundaryRows + SWEEP_BLOCK_ROWS - 1) / SWEEP_BLOCK_ROWS + 3u) : allBlocks;
  if (rows != RTP_ROWS_ALL && h->rowBoundsHost)
  {
    // tighter: the bounds of a recent step (copied to the host without synchronisation, a step or two old) plus a margin the
    // populations cannot outrun in that time; a launch that turns out too small raises rowBounds[3] (sweep.cuh), which the
    // caller reads with its capacity flags
    const u32 b0 = h->rowBoundsHost[0], b1 = h->rowBoundsHost[1], nr = h->rowBoundsHost[2];
    const u32 margin = 64u; // blocks = 8192 rows per side
    if (nr != 0u && nr <= h->s.N && b0 < b1 && b1 <= nr)
    {
      const u32 i0 = (b0 + SWEEP_BLOCK_ROWS - 1) / SWEEP_BLOCK_ROWS, i1 = b1 / SWEEP_BLOCK_ROWS, nb = (nr + SWEEP_BLOCK_ROWS - 1) / SWEEP_BLOCK_ROWS;
      const u32 end = sPhase.rowPhaseToEnd ? allBlocks : nb; // (to the end: the rows without a particle too)
      const u32 want = rows == RTP_ROWS_INTERIOR ? (i1 > i0 ? i1 - i0 : 0u) + margin : i0 + (end > i1 ? end - i1 : 0u) + 2u * margin;
      sPhase.rowPhaseBlocks = min(sPhase.rowPhaseBlocks, want);
    }
  }
  sPhase.rowPhaseBounds = rows != RTP_ROWS_ALL ? h->rowBounds : nullptr; // (a step runs all its sweeps one way or the other:
                                                                          //  the straggler queues are split by row class)
  const bool completes = rows != RTP_ROWS_INTERIOR;
  DeviceState& s = stage >= RTP_SHARD_DENSITY_LAMBDA && stage <= RTP_SHARD_XSPH ? sPhase : h->s;
  cudaStream_t st = h->stream;
  if (rows == RTP_ROWS_BOUNDARY)
  {
    const int rc = ensureExchangeStream(h);
    if (rc != RTP_OK)
      return rc;
    // after everything the compute stream held when this stage began (the INTERIOR call recorded it), beside its interior launch
    st = h->bndStream;
    CUDA_TRY(h, cudaStreamWaitEvent(st, h->bndPre, 0));
  }
  else
  {
    if (h->bndPending) // the previous stage is complete once its boundary launch is
    {
      CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->bndDone, 0));
      h->bndPending = false;
    }
    if (rows == RTP_ROWS_INTERIOR)
    {
      const int rc = ensureExchangeStream(h);
      if (rc != RTP_OK)
        return rc;
      CUDA_TRY(h, cudaEventRecord(h->bndPre, h->stream));
    }
  }
  const GridParams& g = h->g;
  const SphConsts& c = h->c;
  const bool lists = s.nbrList != nullptr && h->jacobi + 1 < NBR_EPOCH_TEMP;
  u32* keysIn = (h->cellPlan.passes % 2 == 0) ? s.cellID : s.keysTmp;
  // iteration it reads shardCur and writes the other prediction buffer
  switch (stage)
  {
  case RTP_SHARD_PREDICT:
  {
    DeviceState own = s;
    own.N = min(s.nOwned, s.N);
    launchFluidPredict(own, g, h->fp, keysIn, nullptr, nullptr, nullptr, st);
    break;
  }
  case RTP_SHARD_PREDICT_FROM:
  {
    // prediction and cell ids of the owned rows [iter, n_owned) only (the arrival slots of a migration), nothing reset
    const u32 end = min(s.nOwned, s.N), first = min((u32)max(iter, 0), end);
    DeviceState own = s;
    own.N = end - first;
    own.posA = s.posA + first, own.velA = s.velA + first, own.pred0 = s.pred0 + first;
    launchFluidPredict(own, g, h->fp, keysIn + first, nullptr, nullptr, nullptr, st, false);
    break;
  }
  case RTP_SHARD_GHOST_KEYS:
  {
    const u32 first = min(s.nOwned, s.N);
    if (s.N > first)
    {
      DeviceState gh = s;
      gh.N = s.N - first;
      gh.posA = s.pred0 + first; // cell ids of the ghosts come from their predicted positions
      launchBoidsCellIds(gh, g, keysIn + first, st);
    }
    break;
  }
  case RTP_SHARD_SORT:
    enqueueSort(h->cellPlan, s.cellID, s.perm, s.keysTmp, s.permTmp, s.sortCtrl, s.sortStatus, st);
    launchFluidGather(s, g, st);
    if (h->rowBounds)
    {
      // the slab's rows live in the interior layers, two face layers and two ghost layers per side
      const u32 plane = (u32)(g.res[1] * g.res[2]), reach = 2u * RTP_SHARD_GHOST_LAYERS * plane;
      const u32 scanLo = h->interiorCellLo > reach ? h->interiorCellLo - reach : 0u, scanHi = min(g.numCells, h->interiorCellHi + reach);
      launchRowPhaseBounds(s, scanLo, scanHi, h->interiorCellLo, h->interiorCellHi, h->rowBounds, st);
      cudaMemcpyAsync((void*)h->rowBoundsHost, h->rowBounds, 3 * sizeof(u32), cudaMemcpyDeviceToHost, st); // read whenever it has landed
      launchAdjustEndCell(s, g, st, scanLo, scanHi);
    }
    else
      launchAdjustEndCell(s, g, st);
    h->shardCur = s.pred1;
    h->predFinal = s.pred1;
    break;
  case RTP_SHARD_DENSITY_LAMBDA:
    launchDensityLambda(s, RTP_MODEL_FLUIDS, g, c, h->fp, h->shardCur, !lists ? NBR_OFF : (iter == 0 ? NBR_BUILD : NBR_BUILD_IF_INVALID), iter, st);
    break;
  case RTP_SHARD_CORRECTION:
  {
    float4* nxt = (h->shardCur == s.pred1) ? s.pred0 : s.pred1;
    launchCorrection(s, RTP_MODEL_FLUIDS, g, c, h->fp, h->cp, h->shardCur, nxt, last != 0, false, lists ? NBR_USE : NBR_OFF, iter, st);
    if (completes)
    {
      h->shardCur = nxt;
      h->predFinal = nxt;
    }
    break;
  }
  case RTP_SHARD_VORTICITY:
    launchVorticity(s, RTP_MODEL_FLUIDS, g, c, h->shardCur, lists ? NBR_BUILD_IF_INVALID : NBR_OFF, iter, st);
    break;
  case RTP_SHARD_CONFINEMENT:
    launchConfinement(s, RTP_MODEL_FLUIDS, g, c, h->fp, h->shardCur, lists ? NBR_USE : NBR_OFF, iter, st);
    break;
  case RTP_SHARD_XSPH:
    launchXsph(s, RTP_MODEL_FLUIDS, g, c, h->fp, h->cp, h->shardCur, lists ? NBR_USE : NBR_OFF, iter, st);
    break;
  case RTP_SHARD_DROP_GHOSTS:
  {
    // owned particles (unsorted index < nOwned) first, cell-sorted order preserved: stable 1-bit partition
    const u32 nOwn = min(s.nOwned, s.N);
    if (s.N > nOwn)
    {
      const SortPlan plan = makeSortPlan(s.N, 1);
      // 1 pass (odd): input keys in the second buffer, result in the first; cameraDist/cameraPerm are free scratch here
      launchGhostFlags(s, s.keysTmp, st);
      enqueueSort(plan, s.cameraDist, s.cameraPerm, s.keysTmp, s.permTmp, s.sortCtrl, s.sortStatus, st);
      launchCompactGather(s, s.cameraPerm, nOwn, st);
      cudaMemcpyAsync(s.posA, s.posB, (size_t)nOwn * sizeof(float4), cudaMemcpyDeviceToDevice, st);
      cudaMemcpyAsync(s.velA, s.velB, (size_t)nOwn * sizeof(float4), cudaMemcpyDeviceToDevice, st);
    }
    break;
  }
  default: return fail(h, RTP_ERR_INVALID, "unknown shard stage");
  }
  CUDA_TRY(h, cudaGetLastError());
  if (rows == RTP_ROWS_BOUNDARY)
  {
    CUDA_TRY(h, cudaEventRecord(h->bndDone, h->bndStream));
    h->bndPending = true;
  }
  return RTP_OK;
}

extern "C" int rtp_shard_buffer(rtp_handle* h, int which, void** dptr, size_t* bytes)
{
  if (!h || !dptr)
    return RTP_ERR_INVALID;
  DeviceState& s = h->s;
  const size_t M = s.M;
  void* p = nullptr;
  size_t b = 0;
  switch (which)
  {
  case RTP_SHARD_BUF_KEYS_IN: p = (h->cellPlan.passes % 2 == 0) ? s.cellID : s.keysTmp; b = 4 * M; break;
  case RTP_SHARD_BUF_PRED_IN: p = s.pred0; b = 16 * M; break;
  case RTP_SHARD_BUF_PRED_CUR: p = h->shardCur ? h->shardCur : s.pred1; b = 16 * M; break;
  case RTP_SHARD_BUF_LAMBDA: p = s.lambda; b = 4 * M; break;
  case RTP_SHARD_BUF_VEL_SORTED: p = s.velB; b = 16 * M; break;
  case RTP_SHARD_BUF_VORT_NORM: p = s.vortNorm; b = 4 * M; break;
  case RTP_SHARD_BUF_VEL_CONFINED: p = s.velC; b = 16 * M; break;
  case RTP_SHARD_BUF_POS: p = s.posA; b = 16 * M; break;
  case RTP_SHARD_BUF_VEL: p = s.velA; b = 16 * M; break;
  case RTP_SHARD_BUF_LIST_BUILD_POS: p = s.nbrBuildPos; b = 16 * M; break;
  case RTP_SHARD_BUF_ROW_BOUNDS: p = h->rowBounds; b = 4 * sizeof(u32); break;
  case RTP_SHARD_BUF_LIST_INVALID: p = s.nbrInvalid; b = 4 * (size_t)2 * NBR_EPOCHS; break;
  default: return fail(h, RTP_ERR_INVALID, "unknown shard buffer");
  }
  if (!p)
    return fail(h, RTP_ERR_STATE, "buffer does not exist for this model / configuration");
  *dptr = p;
  if (bytes)
    *bytes = b;
  return RTP_OK;
}

static int shardRows(rtp_handle* h, int buffer, void** p, int* rowBytes)
{
  size_t bytes = 0;
  const int rc = rtp_shard_buffer(h, buffer, p, &bytes);
  if (rc != RTP_OK)
    return rc;
  *rowBytes = (buffer == RTP_SHARD_BUF_LAMBDA || buffer == RTP_SHARD_BUF_VORT_NORM || buffer == RTP_SHARD_BUF_KEYS_IN) ? 4 : 16;
  return RTP_OK;
}

extern "C" int rtp_shard_pack(rtp_handle* h, int buffer, const uint32_t* d_idx, uint64_t n, void* d_out)
{
  if (!h || (n && (!d_idx || !d_out)) || n > h->s.M)
    return RTP_ERR_INVALID;
  void* p;
  int rb;
  const int rc = shardRows(h, buffer, &p, &rb);
  if (rc != RTP_OK)
    return rc;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  launchPackRows(p, rb, d_idx, (u32)n, d_out, h->shardStream());
  CUDA_TRY(h, cudaGetLastError());
  return RTP_OK;
}

extern "C" int rtp_shard_unpack(rtp_handle* h, int buffer, const uint32_t* d_idx, uint64_t n, const void* d_in)
{
  if (!h || (n && (!d_idx || !d_in)) || n > h->s.M)
    return RTP_ERR_INVALID;
  void* p;
  int rb;
  const int rc = shardRows(h, buffer, &p, &rb);
  if (rc != RTP_OK)
    return rc;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  launchUnpackRows(p, rb, d_idx, (u32)n, d_in, h->shardStream());
  CUDA_TRY(h, cudaGetLastError());
  return RTP_OK;
}

extern "C" int rtp_shard_classify(rtp_handle* h, uint64_t n, uint32_t layer_below, uint32_t layer_from, uint8_t* d_below, uint8_t* d_above)
{
  if (!h || n > h->s.M || (n && (!d_below || !d_above)))
    return RTP_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  const u32* keysIn = (h->cellPlan.passes % 2 == 0) ? h->s.cellID : h->s.keysTmp; // RTP_SHARD_BUF_KEYS_IN
  launchClassifyRows(h->s, h->g, keysIn, (u32)n, layer_below, layer_from, d_below, d_above, h->stream);
  CUDA_TRY(h, cudaGetLastError());
  return RTP_OK;
}

extern "C" int rtp_shard_clear_rows(rtp_handle* h, const uint32_t* d_idx, uint64_t n)
{
  if (!h || (n && !d_idx) || n > h->s.M)
    return RTP_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  launchClearRows(h->s, h->g, (h->cellPlan.passes % 2 == 0) ? h->s.cellID : h->s.keysTmp, d_idx, (u32)n, h->stream);
  CUDA_TRY(h, cudaGetLastError());
  return RTP_OK;
}

extern "C" int rtp_shard_inverse_perm(rtp_handle* h, uint32_t* d_inv)
{
  if (!h || !d_inv)
    return RTP_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  launchInversePerm(h->s, d_inv, h->stream);
  CUDA_TRY(h, cudaGetLastError());
  return RTP_OK;
}

extern "C" int rtp_shard_check_ghosts(rtp_handle* h, const uint32_t* d_sorted_idx, uint64_t n, int next_epoch)
{
  if (!h || (n && !d_sorted_idx) || next_epoch < 0 || next_epoch >= NBR_EPOCHS)
    return RTP_ERR_INVALID;
  if (!h->s.nbrBuildPos || !n)
    return RTP_OK; // lists off: nothing to invalidate
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  launchGhostDisplacement(h->s, h->shardCur ? h->shardCur : h->s.pred1, d_sorted_idx, (u32)n, h->c.nbrDmaxSq, h->s.nbrInvalid + NBR_EPOCHS + next_epoch, h->shardStream());
  CUDA_TRY(h, cudaGetLastError());
  return RTP_OK;
}

// ---- overlap of a ghost refresh with the sweeps of the interior rows
extern "C" int rtp_shard_set_interior(rtp_handle* h, uint32_t cell_lo, uint32_t cell_hi, uint64_t max_boundary_rows)
{
  if (!h || cell_lo > cell_hi || cell_hi > h->g.numCells)
    return RTP_ERR_INVALID;
  h->maxBoundaryRows = (u32)(max_boundary_rows < h->s.M ? max_boundary_rows : h->s.M);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  if (!h->rowBounds)
  {
    const int rc = devAlloc(h, &h->rowBounds, (size_t)4);
    if (rc == RTP_OK)
    {
      void* hp = nullptr;
      CUDA_TRY(h, cudaHostAlloc(&hp, 4 * sizeof(u32), cudaHostAllocDefault));
      memset(hp, 0, 4 * sizeof(u32));
      h->rowBoundsHost = (volatile u32*)hp;
    }
    if (rc != RTP_OK)
      return rc;
  }
  h->interiorCellLo = cell_lo, h->interiorCellHi = cell_hi;
  launchRowPhaseBounds(h->s, 0, 0, 0, 0, h->rowBounds, h->stream); // empty interior, no rows until the next RTP_SHARD_SORT
  return RTP_OK;
}

static int ensureExchangeStream(rtp_handle* h)
{
  if (h->exchStream)
    return RTP_OK;
  int lo = 0, hi = 0;
  CUDA_TRY(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CUDA_TRY(h, cudaStreamCreateWithPriority(&h->exchStream, cudaStreamNonBlocking, hi)); // small kernels: ahead of the sweeps' CTAs
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->exchFork, cudaEventDisableTiming));
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->exchDone, cudaEventDisableTiming));
  CUDA_TRY(h, cudaStreamCreateWithPriority(&h->bndStream, cudaStreamNonBlocking, hi));
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->bndPre, cudaEventDisableTiming));
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->bndDone, cudaEventDisableTiming));
  return RTP_OK;
}

extern "C" int rtp_shard_exchange_stream(rtp_handle* h, void** stream)
{
  if (!h || !stream)
    return RTP_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  const int rc = ensureExchangeStream(h);
  *stream = (void*)h->exchStream;
  return rc;
}

extern "C" int rtp_shard_exchange_fork(rtp_handle* h)
{
  if (!h)
    return RTP_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  const int rc = ensureExchangeStream(h);
  if (rc != RTP_OK)
    return rc;
  if (h->exchActive)
    return fail(h, RTP_ERR_STATE, "an exchange is already open");
  // the rows that travel are boundary rows: produced by the boundary launch of the stage when the sweeps run by row phase
  CUDA_TRY(h, cudaEventRecord(h->exchFork, h->bndPending ? h->bndStream : h->stream));
  CUDA_TRY(h, cudaStreamWaitEvent(h->exchStream, h->exchFork, 0));
  h->exchActive = true;
  return RTP_OK;
}

extern "C" int rtp_shard_exchange_done(rtp_handle* h)
{
  if (!h)
    return RTP_ERR_INVALID;
  if (!h->exchActive)
    return fail(h, RTP_ERR_STATE, "no exchange is open");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  CUDA_TRY(h, cudaEventRecord(h->exchDone, h->exchStream));
  h->exchActive = false;
  h->exchPending = true;
  return RTP_OK;
}

extern "C" int rtp_shard_exchange_join(rtp_handle* h)
{
  if (!h)
    return RTP_ERR_INVALID;
  if (h->exchActive)
    return fail(h, RTP_ERR_STATE, "the exchange is still open (rtp_shard_exchange_done first)");
  if (!h->exchPending)
    return RTP_OK;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  // (only the boundary rows read ghosts: with sweeps by row phase, their stream waits)
  CUDA_TRY(h, cudaStreamWaitEvent(h->rowBounds ? h->bndStream : h->stream, h->exchDone, 0));
  h->exchPending = false;
  return RTP_OK;
}

// ------------------------------------------------------------------ stand-alone sort

extern "C" int rtp_sort_keys(rtp_handle* h, const uint32_t* d_keys_in, uint32_t* d_keys_out, uint32_t* d_perm_out, uint64_t n,
    int key_bits)
{
  if (!h || !d_keys_in || !d_keys_out || !d_perm_out || n >= (1ull << 30))
    return RTP_ERR_INVALID;
  if (n == 0)
    return RTP_OK;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  const SortPlan plan = makeSortPlan((u32)n, key_bits);
  u32 *k1 = nullptr, *v1 = nullptr, *ctrl = nullptr, *status = nullptr;
  cudaError_t e = cudaMalloc(&k1, n * 4);
  if (e == cudaSuccess)
    e = cudaMalloc(&v1, n * 4);
  if (e == cudaSuccess)
    e = cudaMalloc(&ctrl, SORT_CTRL_WORDS * 4);
  if (e == cudaSuccess)
    e = cudaMalloc(&status, sortStatusWords(plan) * 4);
  if (e == cudaSuccess)
  {
    u32* start = (plan.passes % 2 == 0) ? d_keys_out : k1;
    e = cudaMemcpyAsync(start, d_keys_in, n * 4, cudaMemcpyDeviceToDevice, h->stream);
  }
  if (e == cudaSuccess)
  {
    enqueueSort(plan, d_keys_out, d_perm_out, k1, v1, ctrl, status, h->stream);
    e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess)
      e = cudaGetLastError();
  }
  cudaFree(k1); // (cudaFree(nullptr) is a no-op: every path releases what it got)
  cudaFree(v1);
  cudaFree(ctrl);
  cudaFree(status);
  CUDA_TRY(h, e);
  return RTP_OK;
}

extern "C" int rtp_sort_keys_host(rtp_handle* h, const uint32_t* keys_in, uint32_t* keys_out, uint32_t* perm_out, uint64_t n,
    int key_bits)
{
  if (!h || !keys_in || !keys_out || !perm_out || n >= (1ull << 30))
    return RTP_ERR_INVALID;
  if (n == 0)
    return RTP_OK;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  u32 *din = nullptr, *dk = nullptr, *dp = nullptr;
  int rc = RTP_OK;
  if (cudaMalloc(&din, n * 4) != cudaSuccess || cudaMalloc(&dk, n * 4) != cudaSuccess || cudaMalloc(&dp, n * 4) != cudaSuccess
      || cudaMemcpyAsync(din, keys_in, n * 4, cudaMemcpyHostToDevice, h->stream) != cudaSuccess)
  {
    (void)cudaGetLastError();
    rc = fail(h, RTP_ERR_CUDA, "rtp_sort_keys_host: device scratch allocation or upload failed");
  }
  if (rc == RTP_OK)
    rc = rtp_sort_keys(h, din, dk, dp, n, key_bits);
  if (rc == RTP_OK)
  {
    if (cudaMemcpyAsync(keys_out, dk, n * 4, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess
        || cudaMemcpyAsync(perm_out, dp, n * 4, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess
        || cudaStreamSynchronize(h->stream) != cudaSuccess)
      rc = fail(h, RTP_ERR_CUDA, "rtp_sort_keys_host: copy back failed");
  }
  cudaFree(din);
  cudaFree(dk);
  cudaFree(dp);
  return rc;
}

extern "C" int rtp_selftest_math(rtp_handle* h, float lo, float hi, uint64_t* sqrt_mismatches, uint64_t* rcp_mismatches)
{
  if (!h || !(lo > 0.0f) || !(hi >= lo))
    return RTP_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  unsigned long long* d = nullptr;
  CUDA_TRY(h, cudaMalloc(&d, 16));
  cudaMemsetAsync(d, 0, 16, h->stream);
  u32 lb, hb;
  memcpy(&lb, &lo, 4);
  memcpy(&hb, &hi, 4);
  launchSelftestMath(lb, hb, d, h->stream);
  unsigned long long r[2] = { 0, 0 };
  cudaMemcpyAsync(r, d, 16, cudaMemcpyDeviceToHost, h->stream);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  cudaFree(d);
  CUDA_TRY(h, e);
  if (sqrt_mismatches)
    *sqrt_mismatches = r[0];
  if (rcp_mismatches)
    *rcp_mismatches = r[1];
  return RTP_OK;
}

// ------------------------------------------------------------------ profiling

extern "C" int rtp_enable_profiling(rtp_handle* h, int enable)
{
  if (!h)
    return RTP_ERR_INVALID;
  h->profiling = enable != 0;
  return RTP_OK;
}

extern "C" int rtp_get_stage_times(rtp_handle* h, const char** names, float* ms, int cap)
{
  if (!h)
    return RTP_ERR_INVALID;
  const int n = (int)h->stageMs.size();
  for (int i = 0; i < n && i < cap; ++i)
  {
    if (names)
      names[i] = h->stageNames[i].c_str();
    if (ms)
      ms[i] = h->stageMs[i];
  }
  return n;
}

extern "C" int rtp_list_stats(rtp_handle* h, unsigned long long out[8])
{
  if (!h || !out)
    return RTP_ERR_INVALID;
  for (int k = 0; k < 8; ++k)
    out[k] = 0ull;
  if (!h->nbrEnabled || !h->s.nbrCount)
    return RTP_OK;
  unsigned long long* d = nullptr;
  if (cudaMalloc(&d, 64) != cudaSuccess)
    return RTP_ERR_CUDA;
  cudaMemsetAsync(d, 0, 64, h->stream);
  launchListStats(h->s, h->g, h->c, h->predFinal, d, h->stream);
  cudaMemcpyAsync(out, d, 64, cudaMemcpyDeviceToHost, h->stream);
  const cudaError_t e = cudaStreamSynchronize(h->stream);
  cudaFree(d);
  return e == cudaSuccess ? RTP_OK : RTP_ERR_CUDA;
}

extern "C" int rtp_last_launch_count(const rtp_handle* h) { return h ? h->lastLaunches : 0; }

// ------------------------------------------------------------------ initial conditions (host)

// uniform lattice: point (ix,iy,iz) = start + i * (end - start) / res, x-major order  (utils/Geometry.cpp:198-227)
extern "C" int64_t rtp_gen_box_grid(float* out, const int res[3], const float start[3], const float end[3])
{
  if (!out || !res || !start || !end || res[0] <= 0 || res[1] <= 0 || res[2] <= 0)
    return RTP_ERR_INVALID;
  float sp[3];
  for (int k = 0; k < 3; ++k)
    sp[k] = (end[k] - start[k]) / res[k];
  int64_t n = 0;
  for (int ix = 0; ix < res[0]; ++ix)
    for (int iy = 0; iy < res[1]; ++iy)
      for (int iz = 0; iz < res[2]; ++iz, ++n)
      {
        float* o = out + 4 * n;
        o[0] = start[0] + ix * sp[0];
        o[1] = start[1] + iy * sp[1];
        o[2] = start[2] + iz * sp[2];
        o[3] = 0.0f;
      }
  return n;
}

// spherical lattice (phi, theta, r) around the box centre, radius = half diagonal  (utils/Geometry.cpp:243-272)
extern "C" int64_t rtp_gen_sphere_grid(float* out, const int res[3], const float start[3], const float end[3])
{
  if (!out || !res || !start || !end || res[0] <= 0 || res[1] <= 0 || res[2] <= 0)
    return RTP_ERR_INVALID;
  const float PI_F = 3.1415927f;
  float v[3], ctr[3];
  for (int k = 0; k < 3; ++k)
  {
    v[k] = end[k] - start[k];
    ctr[k] = start[k] + v[k] / 2.0f;
  }
  const float radius = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) / 2.0f;
  const float dphi = PI_F / res[0], dtheta = 2.0f * PI_F / res[1], dr = radius / res[2];
  int64_t n = 0;
  for (int ip = 0; ip < res[0]; ++ip)
    for (int it = 0; it < res[1]; ++it)
      for (int ir = 0; ir < res[2]; ++ir, ++n)
      {
        float* o = out + 4 * n;
        const float r = (ir + 1) * dr;
        o[0] = ctr[0] + r * cosf(it * dtheta) * sinf(ip * dphi);
        o[1] = ctr[1] + r * sinf(it * dtheta) * sinf(ip * dphi);
        o[2] = ctr[2] + r * cosf(ip * dphi);
        o[3] = 0.0f;
      }
  return n;
}

// planar lattices of the 2D presets (utils/Geometry.cpp:8-196). plane: 0 = XY, 1 = XZ, 2 = YZ; res = points along the
// plane's first / second axis (rectangle) or angular / radial subdivisions (circle). The off-plane coordinate is
// start's (rectangle) or the centre's (circle).
static bool planeAxes(int plane, int& a, int& b)
{
  if (plane < 0 || plane > 2)
    return false;
  a = plane == 2 ? 1 : 0;
  b = plane == 0 ? 1 : 2;
  return true;
}
extern "C" int64_t rtp_gen_rectangle_grid(float* out, int plane, const int res[2], const float start[3], const float end[3])
{
  int a, b;
  if (!out || !res || !start || !end || res[0] <= 0 || res[1] <= 0 || !planeAxes(plane, a, b))
    return RTP_ERR_INVALID;
  const float da = (end[a] - start[a]) / res[0], db = (end[b] - start[b]) / res[1];
  int64_t n = 0;
  for (int ia = 0; ia < res[0]; ++ia)
    for (int ib = 0; ib < res[1]; ++ib, ++n)
    {
      float* o = out + 4 * n;
      // the reference adds index * spacing on every axis; off the plane that is start + 0 * 0
      o[0] = start[0] + 0 * 0.0f;
      o[1] = start[1] + 0 * 0.0f;
      o[2] = start[2] + 0 * 0.0f;
      o[3] = 0.0f;
      o[a] = start[a] + ia * da;
      o[b] = start[b] + ib * db;
    }
  return n;
}
extern "C" int64_t rtp_gen_circle_grid(float* out, int plane, const int res[2], const float start[3], const float end[3])
{
  int a, b;
  if (!out || !res || !start || !end || res[0] <= 0 || res[1] <= 0 || !planeAxes(plane, a, b))
    return RTP_ERR_INVALID;
  const float PI_F = 3.1415927f;
  float v[3], ctr[3];
  for (int k = 0; k < 3; ++k)
  {
    v[k] = end[k] - start[k];
    ctr[k] = start[k] + v[k] / 2.0f;
  }
  const float radius = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) / 2.0f;
  const float dang = 2.0f * PI_F / res[0], dr = radius / res[1];
  int64_t n = 0;
  for (int io = 0; io < res[0]; ++io)
    for (int ir = 0; ir < res[1]; ++ir, ++n)
    {
      float* o = out + 4 * n;
      o[0] = ctr[0];
      o[1] = ctr[1];
      o[2] = ctr[2];
      o[3] = 0.0f;
      o[a] = ctr[a] + ((ir + 1) * dr) * cosf(io * dang);
      o[b] = ctr[b] + ((ir + 1) * dr) * sinf(io * dang);
    }
  return n;
}

// uniform random fill driven by glibc rand(), x, y, z call order  (utils/Geometry.cpp:229-239)
extern "C" int64_t rtp_gen_random_box(float* out, int64_t n, const float start[3], const float end[3], int seed)
{
  if (!out || !start || !end || n < 0)
    return RTP_ERR_INVALID;
  if (seed >= 0)
    srand((unsigned)seed);
  for (int64_t i = 0; i < n; ++i)
    for (int k = 0; k < 4; ++k)
      out[4 * i + k] = (k < 3) ? (float)rand() / (float)RAND_MAX * (end[k] - start[k]) + start[k] : 0.0f;
  return n;
}

// ------------------------------------------------------------------ boids target trajectory (host)

// The boids target wanders on a pseudo-random path driven by three Perlin-noise channels (physics/utils/Target.cpp:11-50,
// physics/utils/PerlinNoise.cpp:10-85); the reference evaluates it on the CPU once per frame (Boids.cpp:351-358) and so
// does this backend. The permutation tables come from std::shuffle over std::default_random_engine(seed), like the
// reference's: the same C++ library gives the same tables.
namespace
{
struct NoiseChannel
{
  std::vector<int> perm;
  explicit NoiseChannel(int seed) : perm(256)
  {
    std::iota(perm.begin(), perm.end(), 0);
    std::default_random_engine engine(seed);
    std::shuffle(perm.begin(), perm.end(), engine);
    perm.insert(perm.end(), perm.begin(), perm.end());
  }
  static float fade(float t) { return t * t * t * (t * (t * 6 - 15) + 10); }
  static float mix(float t, float a, float b) { return a + t * (b - a); }
  static float gradient(int hash, float x, float y, float z)
  {
    const int h = hash & 15;
    const float u = h < 8 ? x : y;
    const float v = h < 4 ? y : ((h == 12 || h == 14) ? x : z);
    return ((h & 1) == 0 ? u : -u) + ((h & 2) == 0 ? v : -v);
  }
  float value(float x, float y, float z) const
  {
    const int X = (int)floor(x) & 255, Y = (int)floor(y) & 255, Z = (int)floor(z) & 255;
    x -= floor(x);
    y -= floor(y);
    z -= floor(z);
    const float u = fade(x), v = fade(y), w = fade(z);
    const int A = perm[X] + Y, AA = perm[A] + Z, AB = perm[A + 1] + Z;
    const int B = perm[X + 1] + Y, BA = perm[B] + Z, BB = perm[B + 1] + Z;
    const float lo = mix(v, mix(u, gradient(perm[AA], x, y, z), gradient(perm[BA], x - 1, y, z)),
        mix(u, gradient(perm[AB], x, y - 1, z), gradient(perm[BB], x - 1, y - 1, z)));
    const float hi = mix(v, mix(u, gradient(perm[AA + 1], x, y, z - 1), gradient(perm[BA + 1], x - 1, y, z - 1)),
        mix(u, gradient(perm[AB + 1], x, y - 1, z - 1), gradient(perm[BB + 1], x - 1, y - 1, z - 1)));
    return (mix(w, lo, hi) + 1.0f) / 2.0f;
  }
};
} // namespace

struct rtp_target
{
  NoiseChannel theta { 1 }, beta { 29 }, radial { 246 };
  float walker[3] = { 0.0f, 0.0f, 0.0f };
  float pos[3] = { 0.0f, 0.0f, 0.0f };
  float maxRadius = 0.0f;
};

extern "C" rtp_target* rtp_target_create(uint32_t box_size)
{
  rtp_target* t = new rtp_target();
  t->maxRadius = 0.48f * box_size;
  return t;
}
extern "C" void rtp_target_destroy(rtp_target* t) { delete t; }
extern "C" int rtp_target_update(rtp_target* t, int dim, float particles_velocity, float out_pos[3])
{
  if (!t || !out_pos)
    return RTP_ERR_INVALID;
  const float PI_T = 3.14f; // Target.cpp:25
  t->walker[0] += 0.001f;
  t->walker[1] += 0.002f;
  t->walker[2] += 0.01f;
  const float nTheta = t->theta.value(t->walker[0], t->walker[1], t->walker[2]);
  const float nBeta = t->beta.value(t->walker[0], t->walker[1], t->walker[2]);
  const float nR = t->radial.value(t->walker[0], t->walker[1], t->walker[2]);
  const float velRatio = particles_velocity / 9.5f;
  const float radius = t->maxRadius * cos(12 * velRatio * PI_T * nR);
  t->pos[0] = dim == 3 ? (radius * cos(5 * velRatio * PI_T * nBeta)) : 0.0f;
  t->pos[1] = radius * sin(5 * velRatio * PI_T * nBeta) * cos(4 * velRatio * PI_T * nTheta);
  t->pos[2] = radius * sin(5 * velRatio * PI_T * nBeta) * sin(4 * velRatio * PI_T * nTheta);
  memcpy(out_pos, t->pos, sizeof t->pos);
  return RTP_OK;
}
