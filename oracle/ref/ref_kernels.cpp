// ref_kernels.cpp -- the reference's own OpenCL C kernels, compiled for the CPU through ocl_shim.hpp.
// TEST INFRASTRUCTURE ONLY. Built by oracle/ref/Makefile into oracle/_ref/libref_kernels.so when /root/reference is
// present; the kernel sources are read from where they lie (REF_GEN_DIR holds a temporary, syntactically rewritten
// copy that is deleted after the build). What is restated here is only the HOST side: which kernel runs when, with
// which buffers (Boids::update physics/ocl/Boids.cpp:323-384, Fluids::update Fluids.cpp:400-471, Clouds::update
// Clouds.cpp:503-627, RadixSort::sort physics/utils/RadixSort.cpp:122-190 as "stable sort + gather").
#include "ocl_shim.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../include/rtp_cuda.h"

namespace ocl
{
thread_local size_t g_globalId = 0;
size_t g_globalSize = 0;

// the -D constants of createProgram() (Fluids.cpp:104-119 etc.) as run-time values
struct BuildOptions
{
  float v_EFFECT_RADIUS, v_EFFECT_RADIUS_SQUARED, v_ABS_WALL_X, v_ABS_WALL_Y, v_ABS_WALL_Z, v_GRID_CELL_SIZE_XYZ, v_POLY6_COEFF, v_SPIKY_COEFF, v_MAX_VEL;
  int v_GRID_RES_X, v_GRID_RES_Y, v_GRID_RES_Z;
  uint v_GRID_NUM_CELLS, v_NUM_MAX_PARTS_IN_CELL;
};
static BuildOptions g_opt;
}

#define EFFECT_RADIUS (ocl::g_opt.v_EFFECT_RADIUS)
#define EFFECT_RADIUS_SQUARED (ocl::g_opt.v_EFFECT_RADIUS_SQUARED)
#define ABS_WALL_X (ocl::g_opt.v_ABS_WALL_X)
#define ABS_WALL_Y (ocl::g_opt.v_ABS_WALL_Y)
#define ABS_WALL_Z (ocl::g_opt.v_ABS_WALL_Z)
#define GRID_CELL_SIZE_XYZ (ocl::g_opt.v_GRID_CELL_SIZE_XYZ)
#define POLY6_COEFF (ocl::g_opt.v_POLY6_COEFF)
#define SPIKY_COEFF (ocl::g_opt.v_SPIKY_COEFF)
#define MAX_VEL (ocl::g_opt.v_MAX_VEL)
#define GRID_RES_X (ocl::g_opt.v_GRID_RES_X)
#define GRID_RES_Y (ocl::g_opt.v_GRID_RES_Y)
#define GRID_RES_Z (ocl::g_opt.v_GRID_RES_Z)
#define GRID_NUM_CELLS (ocl::g_opt.v_GRID_NUM_CELLS)
#define NUM_MAX_PARTS_IN_CELL (ocl::g_opt.v_NUM_MAX_PARTS_IN_CELL)

// One namespace per OpenCL program, files in the order of createProgram() (Boids.cpp:116, Fluids.cpp:123, Clouds.cpp:156)
namespace ocl
{
namespace boids
{
#include "define.cl"
#include "boids.cl"
#include "utils.cl"
#include "grid.cl"
}
#undef ID
#undef FLOAT_EPS
#undef ABS_GRAVITY_ACC_Y
#undef GRAVITY_ACC
#undef FAR_DIST
#undef MAX_STEERING
namespace fluids
{
#include "define.cl"
#include "sph.cl"
#include "fluids.cl"
#include "utils.cl"
#include "grid.cl"
}
#undef ID
#undef FLOAT_EPS
#undef ABS_GRAVITY_ACC_Y
#undef GRAVITY_ACC
#undef FAR_DIST
#undef WALL_COEFF
namespace clouds
{
#include "define.cl"
#include "sph.cl"
#include "clouds.cl"
#include "grid.cl"
#include "utils.cl"
}
} // namespace ocl

using ocl::float4;
using ocl::float8;
using ocl::uint2;

// Context::runKernel(name, global size): a 1-D NDRange executed work-item by work-item (every kernel on this path
// writes only its own element or a cell no other work-item writes)
template <typename K, typename... A>
static void run(size_t n, K kernel, A... args)
{
  ocl::g_globalSize = n;
#pragma omp parallel for schedule(dynamic, 256)
  for (size_t id = 0; id < n; ++id)
  {
    ocl::g_globalId = id;
    kernel(args...);
  }
}

struct ref_world
{
  int model;
  size_t M, N, C;
  int dim, boundary, jacobi;
  ocl::BuildOptions opt;
  rtp_boids_params boids;
  rtp_target_params target;
  float targetPos[4];
  int targetActive;
  rtp_fluid_params fluid;
  rtp_cloud_params cloud;
  float cam[4];
  int dispField;
  float dispMin, dispMax;
  std::vector<float4> pos, col, vel, acc, predPos, corrPos, velInVisc, vort, totCorrPos, tmp4;
  std::vector<float> density, constFactor, temp, tempIn, lapTemp, corrTemp, constFactorTemp, vaporDens, vaporDensIn, cloudDens,
      cloudDensIn, buoyancy, cloudGen, partID, tmp1;
  std::vector<float8> partDetector;
  std::vector<uint> cellID, cameraDist, perm, cameraPerm;
  std::vector<uint2> startEnd;
};

static float baked(float v)
{
  char buf[128];
  snprintf(buf, sizeof buf, "%.10f", (double)v);
  return strtof(buf, nullptr);
}

template <typename P>
static P asParams(const void* p)
{
  P out;
  static_assert(sizeof(P) == sizeof(rtp_fluid_params) || sizeof(P) == sizeof(rtp_cloud_params) || sizeof(P) == sizeof(rtp_boids_params)
          || sizeof(P) == sizeof(rtp_target_params),
      "parameter block layout");
  memcpy(&out, p, sizeof(P));
  return out;
}

extern "C" {

int ref_create(const rtp_config* cfg, ref_world** out)
{
  ref_world* w = new ref_world();
  w->model = cfg->model;
  w->M = cfg->max_particles;
  w->N = cfg->nb_particles;
  w->C = (size_t)cfg->grid[0] * cfg->grid[1] * cfg->grid[2];
  w->dim = cfg->dim == 2 ? 2 : 3;
  w->boundary = 0;
  w->jacobi = 2;
  const float effectRadius = (float)cfg->box[0] / (float)cfg->grid[0];
  const float PI_F = 3.1415927f;
  ocl::BuildOptions& o = w->opt;
  o.v_EFFECT_RADIUS = baked(effectRadius);
  o.v_EFFECT_RADIUS_SQUARED = baked(1.0f * (float)cfg->box[0] * (float)cfg->box[0] / (float)((size_t)cfg->grid[0] * cfg->grid[0]));
  o.v_ABS_WALL_X = baked((float)cfg->box[0] / 2.0f);
  o.v_ABS_WALL_Y = baked((float)cfg->box[1] / 2.0f);
  o.v_ABS_WALL_Z = baked((float)cfg->box[2] / 2.0f);
  o.v_GRID_CELL_SIZE_XYZ = baked((float)cfg->box[0] / (float)cfg->grid[0]);
  o.v_POLY6_COEFF = baked(315.0f / (64.0f * PI_F * std::pow(effectRadius, 9.f)));
  o.v_SPIKY_COEFF = baked(15.0f / (PI_F * std::pow(effectRadius, 6.f)));
  o.v_MAX_VEL = baked(30.0f);
  o.v_GRID_RES_X = (int)cfg->grid[0];
  o.v_GRID_RES_Y = (int)cfg->grid[1];
  o.v_GRID_RES_Z = (int)cfg->grid[2];
  o.v_GRID_NUM_CELLS = (uint)w->C;
  o.v_NUM_MAX_PARTS_IN_CELL = cfg->max_parts_in_cell ? cfg->max_parts_in_cell : (cfg->model == RTP_MODEL_BOIDS ? 3000u : 100u);
  w->boids = rtp_boids_params { 0.5f, 1.6f, 1.6f, 1.45f };
  w->target = rtp_target_params { 2.0f, 1 };
  w->targetActive = 0;
  w->fluid = rtp_fluid_params { 450.0f, 600.0f, 0.010f, (uint32_t)w->dim, 1, 0.006f, 0.001f, 4, 1, 0.0004f, 0.0001f };
  w->cloud = rtp_cloud_params { (uint32_t)w->dim, 0.01f, 450.0f, 10.0f, 0.10f, 0.0005f, 5.0f, 0.3485f, 0.07f, 1, 600.0f, 0.75f, 1.0f };
  w->cam[0] = 32.0f, w->cam[1] = -1.2f, w->cam[2] = 0.0f, w->cam[3] = 0.0f;
  w->dispField = RTP_F_CLOUD_DENS, w->dispMin = 1.0f, w->dispMax = 15.0f;
  const size_t M = w->M;
  for (auto* v : { &w->pos, &w->col, &w->vel, &w->acc, &w->predPos, &w->corrPos, &w->velInVisc, &w->vort, &w->totCorrPos, &w->tmp4 })
    v->assign(M, float4(0.0f));
  for (auto* v : { &w->density, &w->constFactor, &w->temp, &w->tempIn, &w->lapTemp, &w->corrTemp, &w->constFactorTemp, &w->vaporDens,
           &w->vaporDensIn, &w->cloudDens, &w->cloudDensIn, &w->buoyancy, &w->cloudGen, &w->partID, &w->tmp1 })
    v->assign(M, 0.0f);
  w->partDetector.assign(w->C, float8(0.0f));
  w->cellID.assign(M + 1, 0xFFFFFFFFu); // +1: fillEndCell reads pCellID[N] also when N == M (grid.cl:131)
  w->cameraDist.assign(M, 0u);
  w->perm.assign(M, 0u);
  w->cameraPerm.assign(M, 0u);
  w->startEnd.assign(w->C, uint2(0u, 0u));
  *out = w;
  return 0;
}

void ref_destroy(ref_world* w) { delete w; }

void* ref_field_ptr(ref_world* w, int field, size_t* bytes)
{
  void* p = nullptr;
  size_t b = 0;
  const size_t M = w->M;
#define F4(v) p = v.data(), b = 16 * M
#define F1(v) p = v.data(), b = 4 * M
  switch (field)
  {
  case RTP_F_POS: F4(w->pos); break;
  case RTP_F_COL: F4(w->col); break;
  case RTP_F_VEL: F4(w->vel); break;
  case RTP_F_ACC: F4(w->acc); break;
  case RTP_F_PRED_POS: F4(w->predPos); break;
  case RTP_F_CORR_POS: F4(w->corrPos); break;
  case RTP_F_VORT: F4(w->vort); break;
  case RTP_F_TOT_CORR_POS: F4(w->totCorrPos); break;
  case RTP_F_DENSITY: F1(w->density); break;
  case RTP_F_CONST_FACTOR: F1(w->constFactor); break;
  case RTP_F_TEMP: F1(w->temp); break;
  case RTP_F_VAPOR_DENS: F1(w->vaporDens); break;
  case RTP_F_CLOUD_DENS: F1(w->cloudDens); break;
  case RTP_F_BUOYANCY: F1(w->buoyancy); break;
  case RTP_F_CLOUD_GEN: F1(w->cloudGen); break;
  case RTP_F_PART_ID: F1(w->partID); break;
  case RTP_F_LAPLACIAN_TEMP: F1(w->lapTemp); break;
  case RTP_F_CONST_FACTOR_TEMP: F1(w->constFactorTemp); break;
  case RTP_F_CORR_TEMP: F1(w->corrTemp); break;
  case RTP_F_CELL_ID: p = w->cellID.data(), b = 4 * M; break;
  case RTP_F_CAMERA_DIST: p = w->cameraDist.data(), b = 4 * M; break;
  case RTP_F_START_END_CELL: p = w->startEnd.data(), b = 8 * w->C; break;
  case RTP_F_PERM: p = w->perm.data(), b = 4 * M; break;
  case RTP_F_CAMERA_PERM: p = w->cameraPerm.data(), b = 4 * M; break;
  case RTP_F_PART_DETECTOR: p = w->partDetector.data(), b = 32 * w->C; break;
  default: break;
  }
#undef F4
#undef F1
  if (bytes)
    *bytes = b;
  return p;
}

int ref_set_boids_params(ref_world* w, const rtp_boids_params* r, const rtp_target_params* t, const float tp[4], int active)
{
  if (r) w->boids = *r;
  if (t) w->target = *t;
  if (tp) memcpy(w->targetPos, tp, 16);
  w->targetActive = active;
  return 0;
}
int ref_set_fluid_params(ref_world* w, const rtp_fluid_params* f, int jacobi)
{
  if (f) w->fluid = *f;
  if (jacobi > 0) w->jacobi = jacobi;
  return 0;
}
int ref_set_cloud_params(ref_world* w, const rtp_cloud_params* c)
{
  if (c) w->cloud = *c;
  return 0;
}
int ref_set_boundary(ref_world* w, int b) { w->boundary = b; return 0; }
int ref_set_nb_particles(ref_world* w, uint64_t n) { w->N = n; return 0; }
int ref_set_dimension(ref_world* w, int d)
{
  w->dim = d == 2 ? 2 : 3;
  w->fluid.dim = w->cloud.dim = (uint32_t)w->dim;
  return 0;
}
int ref_set_displayed_quantity(ref_world* w, int field, float lo, float hi)
{
  w->dispField = field, w->dispMin = lo, w->dispMax = hi;
  return 0;
}
void ref_set_camera(ref_world* w, const float cam[3])
{
  if (cam) memcpy(w->cam, cam, 12);
}
float ref_constant(const ref_world* w, const char* name)
{
  const ocl::BuildOptions& o = w->opt;
  if (!strcmp(name, "EFFECT_RADIUS")) return o.v_EFFECT_RADIUS;
  if (!strcmp(name, "EFFECT_RADIUS_SQUARED")) return o.v_EFFECT_RADIUS_SQUARED;
  if (!strcmp(name, "GRID_CELL_SIZE_XYZ")) return o.v_GRID_CELL_SIZE_XYZ;
  if (!strcmp(name, "ABS_WALL_X")) return o.v_ABS_WALL_X;
  if (!strcmp(name, "ABS_WALL_Y")) return o.v_ABS_WALL_Y;
  if (!strcmp(name, "ABS_WALL_Z")) return o.v_ABS_WALL_Z;
  if (!strcmp(name, "POLY6_COEFF")) return o.v_POLY6_COEFF;
  if (!strcmp(name, "SPIKY_COEFF")) return o.v_SPIKY_COEFF;
  if (!strcmp(name, "MAX_VEL")) return o.v_MAX_VEL;
  return NAN;
}
} // extern "C"

// ---------------------------------------------------------------- RadixSort::sort as "stable sort + permutate"
static void sortAndPermute(ref_world* w, std::vector<uint>& keys, std::vector<uint>& perm, bool cameraSort)
{
  const size_t M = w->M;
  std::vector<uint> idx(M);
  std::iota(idx.begin(), idx.end(), 0u); // resetIndex radixSort.cl:171-174
  std::stable_sort(idx.begin(), idx.end(), [&](uint a, uint b) { return keys[a] < keys[b]; });
  std::vector<uint> sorted(M);
  for (size_t i = 0; i < M; ++i)
    sorted[i] = keys[idx[i]];
  std::copy(sorted.begin(), sorted.end(), keys.begin());
  perm = idx;
  // copyBuffer + permutateFloat4 / permutateFloat (RadixSort.cpp:164-189; radixSort.cl:179-202: out[i] = in[perm[i]]).
  // radixSort.cl itself is not compiled: its kernels need work-group barriers; only its contract is used.
  auto g4 = [&](std::vector<float4>& buf)
  {
    w->tmp4 = buf;
    for (size_t i = 0; i < M; ++i)
      buf[i] = w->tmp4[perm[i]];
  };
  auto g1 = [&](std::vector<float>& buf)
  {
    w->tmp1 = buf;
    for (size_t i = 0; i < M; ++i)
      buf[i] = w->tmp1[perm[i]];
  };
  switch (w->model)
  {
  case RTP_MODEL_BOIDS: g4(w->pos); g4(w->col); g4(w->vel); g4(w->acc); break; // Boids.cpp:337, :381
  case RTP_MODEL_FLUIDS: g4(w->pos); g4(w->col); g4(w->vel); g4(w->predPos); break; // Fluids.cpp:417, :468
  case RTP_MODEL_CLOUDS: // Clouds.cpp:543, :624
    g4(w->pos); g4(w->col); g4(w->vel); g4(w->predPos);
    if (!cameraSort) g4(w->totCorrPos);
    g1(w->temp); g1(w->buoyancy); g1(w->vaporDens); g1(w->cloudDens); g1(w->partID);
    break;
  }
}

// ---------------------------------------------------------------- stages == reference kernel launches
// stage ids are those of oracle/rtp_oracle.h (orc_stage)
enum
{
  ST_FILL_CELL_IDS = 0, ST_SORT_BY_CELL, ST_BUILD_CELL_TABLE, ST_BD_RULES, ST_BD_TARGET, ST_BD_UPDATE_VEL, ST_BD_UPDATE_POS,
  ST_PREDICT_POS, ST_APPLY_BOUNDARY, ST_DENSITY, ST_CONSTRAINT_FACTOR, ST_CONSTRAINT_CORRECTION, ST_CORRECT_POS, ST_UPDATE_VEL,
  ST_VORTICITY, ST_VORTICITY_CONFINEMENT, ST_XSPH, ST_UPDATE_POS, ST_CLD_THERMO, ST_CLD_LAPLACIAN_TEMP,
  ST_CLD_CONSTRAINT_FACTOR_TEMP, ST_CLD_CONSTRAINT_CORRECTION_TEMP, ST_CLD_CORRECT_TEMP, ST_RENDER_AUX, ST_CAMERA_SORT
};

extern "C" int ref_reset_ids(ref_world* w)
{
  ocl::g_opt = w->opt;
  run(w->M, ocl::fluids::resetCellIDs, w->cellID.data());
  run(w->M, ocl::fluids::resetCameraDist, w->cameraDist.data());
  w->cellID[w->M] = 0xFFFFFFFFu;
  return 0;
}

extern "C" int ref_init_clouds_fields(ref_world* w)
{
  ocl::g_opt = w->opt;
  namespace K = ocl::clouds;
  const auto cp = asParams<K::CloudParams>(&w->cloud);
  run(w->M, K::cld_initTemperature, (const float4*)w->pos.data(), w->temp.data()); // Clouds.cpp:495
  run(w->M, K::cld_initVaporDensity, cp, (const float*)w->temp.data(), w->vaporDens.data()); // Clouds.cpp:497
  return 0;
}

static int stageBoids(ref_world* w, int stage)
{
  namespace K = ocl::boids;
  const size_t N = w->N;
  const float dt = 0.1f; // Boids.cpp:334
  switch (stage)
  {
  case ST_FILL_CELL_IDS: run(N, K::fillCellIDs, (const float4*)w->pos.data(), w->cellID.data()); break;
  case ST_SORT_BY_CELL: sortAndPermute(w, w->cellID, w->perm, false); break;
  case ST_BUILD_CELL_TABLE:
    run(w->C, K::resetStartEndCell, w->startEnd.data());
    run(N, K::fillStartCell, (const uint*)w->cellID.data(), w->startEnd.data());
    run(N, K::fillEndCell, (const uint*)w->cellID.data(), w->startEnd.data());
    run(w->C, K::adjustEndCell, w->startEnd.data());
    break;
  case ST_BD_RULES:
    if (w->dim == 2)
      run(N, K::bd_applyBoidsRulesWithGrid2D, (const float4*)w->pos.data(), (const float4*)w->vel.data(), (const uint2*)w->startEnd.data(),
          asParams<K::BoidsRuleParams>(&w->boids), w->acc.data());
    else
      run(N, K::bd_applyBoidsRulesWithGrid3D, (const float4*)w->pos.data(), (const float4*)w->vel.data(), (const uint2*)w->startEnd.data(),
          asParams<K::BoidsRuleParams>(&w->boids), w->acc.data());
    break;
  case ST_BD_TARGET:
    if (w->targetActive)
      run(N, K::bd_addTargetRule, (const float4*)w->pos.data(), float4(w->targetPos[0], w->targetPos[1], w->targetPos[2], w->targetPos[3]),
          asParams<K::TargetParams>(&w->target), w->acc.data());
    break;
  case ST_BD_UPDATE_VEL: run(N, K::bd_updateVel, (const float4*)w->acc.data(), dt, w->boids.velocityScale, w->vel.data()); break;
  case ST_BD_UPDATE_POS:
    if (w->boundary == RTP_BOUNDARY_CYCLIC_WALL)
      run(N, K::bd_updatePosAndApplyPeriodicBC, (const float4*)w->vel.data(), dt, w->pos.data());
    else
      run(N, K::bd_updatePosAndApplyWallBC, w->vel.data(), dt, w->pos.data());
    break;
  case ST_RENDER_AUX:
    run(w->C, K::resetGridDetector, w->partDetector.data());
    run(N, K::fillGridDetector, w->pos.data(), w->partDetector.data());
    break;
  case ST_CAMERA_SORT:
    run(N, K::fillCameraDist, (const float4*)w->pos.data(), (const ocl::float3*)w->cam, w->cameraDist.data());
    sortAndPermute(w, w->cameraDist, w->cameraPerm, true);
    break;
  default: return -1;
  }
  return 0;
}

static int stageFluids(ref_world* w, int stage)
{
  namespace K = ocl::fluids;
  const size_t N = w->N;
  const auto fp = asParams<K::FluidParams>(&w->fluid);
  const float4* pred = w->predPos.data();
  const uint2* table = w->startEnd.data();
  switch (stage)
  {
  case ST_PREDICT_POS: run(N, K::fld_predictPosition, (const float4*)w->pos.data(), (const float4*)w->vel.data(), fp, w->predPos.data()); break;
  case ST_FILL_CELL_IDS: run(N, K::fillCellIDs, pred, w->cellID.data()); break;
  case ST_SORT_BY_CELL: sortAndPermute(w, w->cellID, w->perm, false); break;
  case ST_BUILD_CELL_TABLE:
    run(w->C, K::resetStartEndCell, w->startEnd.data());
    run(N, K::fillStartCell, (const uint*)w->cellID.data(), w->startEnd.data());
    run(N, K::fillEndCell, (const uint*)w->cellID.data(), w->startEnd.data());
    run(w->C, K::adjustEndCell, w->startEnd.data());
    break;
  case ST_APPLY_BOUNDARY: run(N, K::fld_applyBoundaryCondition, w->predPos.data()); break;
  case ST_DENSITY: run(N, K::fld_computeDensity, pred, table, fp, w->density.data()); break;
  case ST_CONSTRAINT_FACTOR: run(N, K::fld_computeConstraintFactor, pred, (const float*)w->density.data(), table, fp, w->constFactor.data()); break;
  case ST_CONSTRAINT_CORRECTION: run(N, K::fld_computeConstraintCorrection, (const float*)w->constFactor.data(), table, pred, fp, w->corrPos.data()); break;
  case ST_CORRECT_POS: run(N, K::fld_correctPosition, (const float4*)w->corrPos.data(), w->predPos.data()); break;
  case ST_UPDATE_VEL: run(N, K::fld_updateVel, pred, (const float4*)w->pos.data(), fp, w->vel.data()); break;
  case ST_VORTICITY: run(N, K::fld_computeVorticity, pred, table, (const float4*)w->vel.data(), fp, w->vort.data()); break;
  case ST_VORTICITY_CONFINEMENT: run(N, K::fld_applyVorticityConfinement, pred, table, (const float4*)w->vort.data(), fp, w->vel.data()); break;
  case ST_XSPH:
    w->velInVisc = w->vel; // copyBuffer, Fluids.cpp:451
    run(N, K::fld_applyXsphViscosityCorrection, pred, table, (const float4*)w->velInVisc.data(), fp, w->vel.data());
    break;
  case ST_UPDATE_POS: run(N, K::fld_updatePosition, pred, w->pos.data()); break;
  case ST_RENDER_AUX:
    run(w->C, K::resetGridDetector, w->partDetector.data());
    run(N, K::fillGridDetector, w->pos.data(), w->partDetector.data());
    run(N, K::fld_fillFluidColor, (const float*)w->density.data(), fp, w->col.data());
    break;
  case ST_CAMERA_SORT:
    run(N, K::fillCameraDist, (const float4*)w->pos.data(), (const ocl::float3*)w->cam, w->cameraDist.data());
    sortAndPermute(w, w->cameraDist, w->cameraPerm, true);
    break;
  default: return -1;
  }
  return 0;
}

static int stageClouds(ref_world* w, int stage)
{
  namespace K = ocl::clouds;
  const size_t N = w->N;
  const auto fp = asParams<K::FluidParams>(&w->fluid);
  const auto cp = asParams<K::CloudParams>(&w->cloud);
  const float4* pred = w->predPos.data();
  const float4* pos = w->pos.data();
  const uint2* table = w->startEnd.data();
  switch (stage)
  {
  case ST_CLD_THERMO: // Clouds.cpp:514-529, kernel <-> buffer bindings Clouds.cpp:247-252
    w->tempIn = w->temp;
    run(N, K::cld_heatFromGround, (const float*)w->tempIn.data(), pos, cp, w->temp.data());
    run(N, K::cld_computeBuoyancy, (const float*)w->temp.data(), pos, (const float*)w->cloudDens.data(), cp, w->buoyancy.data());
    run(N, K::cld_applyAdiabaticCooling, (const float*)w->temp.data(), (const float4*)w->vel.data(), cp, w->tempIn.data());
    run(N, K::cld_generateCloud, (const float*)w->tempIn.data(), (const float*)w->vaporDens.data(), (const float*)w->cloudDens.data(), cp, w->cloudGen.data());
    w->vaporDensIn = w->vaporDens;
    w->cloudDensIn = w->cloudDens;
    run(N, K::cld_applyPhaseTransition, (const float*)w->vaporDensIn.data(), (const float*)w->cloudDensIn.data(), (const float*)w->cloudGen.data(), cp,
        w->vaporDens.data(), w->cloudDens.data());
    run(N, K::cld_applyLatentHeat, (const float*)w->tempIn.data(), (const float*)w->cloudGen.data(), cp, w->temp.data());
    break;
  case ST_PREDICT_POS:
    run(N, K::cld_predictPosition, pos, (const float4*)w->vel.data(), (const float*)w->buoyancy.data(), cp, w->predPos.data(), w->totCorrPos.data());
    break;
  case ST_APPLY_BOUNDARY: run(N, K::cld_applyMixedBoundaryConditions, w->predPos.data()); break;
  case ST_FILL_CELL_IDS: run(N, K::fillCellIDs, pred, w->cellID.data()); break;
  case ST_SORT_BY_CELL: sortAndPermute(w, w->cellID, w->perm, false); break;
  case ST_BUILD_CELL_TABLE:
    run(w->C, K::resetStartEndCell, w->startEnd.data());
    run(N, K::fillStartCell, (const uint*)w->cellID.data(), w->startEnd.data());
    run(N, K::fillEndCell, (const uint*)w->cellID.data(), w->startEnd.data());
    run(w->C, K::adjustEndCell, w->startEnd.data());
    break;
  case ST_CLD_LAPLACIAN_TEMP: run(N, K::cld_computeLaplacianTemp, pos, (const float*)w->temp.data(), table, cp, w->lapTemp.data()); break;
  case ST_CLD_CONSTRAINT_FACTOR_TEMP: run(N, K::cld_computeConstraintFactorTemp, pos, (const float*)w->lapTemp.data(), table, cp, w->constFactorTemp.data()); break;
  case ST_CLD_CONSTRAINT_CORRECTION_TEMP: run(N, K::cld_computeConstraintCorrectionTemp, (const float*)w->constFactorTemp.data(), table, pos, cp, w->corrTemp.data()); break;
  case ST_CLD_CORRECT_TEMP: run(N, K::cld_correctTemperature, (const float*)w->corrTemp.data(), w->temp.data()); break;
  case ST_DENSITY: run(N, K::cld_computeDensity, pred, table, fp, w->density.data()); break;
  case ST_CONSTRAINT_FACTOR: run(N, K::cld_computeConstraintFactor, pred, (const float*)w->density.data(), table, fp, w->constFactor.data()); break;
  case ST_CONSTRAINT_CORRECTION: run(N, K::cld_computeConstraintCorrection, (const float*)w->constFactor.data(), table, pred, fp, w->corrPos.data()); break;
  case ST_CORRECT_POS: // Clouds.cpp:579-583
    run(N, K::cld_correctPosition, (const float4*)w->corrPos.data(), w->predPos.data());
    run(N, K::cld_correctPosition, (const float4*)w->corrPos.data(), w->totCorrPos.data());
    break;
  case ST_UPDATE_VEL: run(N, K::cld_updateVel, (const float4*)w->totCorrPos.data(), fp, w->vel.data()); break;
  case ST_VORTICITY: run(N, K::cld_computeVorticity, pred, table, (const float4*)w->vel.data(), fp, w->vort.data()); break;
  case ST_VORTICITY_CONFINEMENT: run(N, K::cld_applyVorticityConfinement, pred, table, (const float4*)w->vort.data(), fp, w->vel.data()); break;
  case ST_XSPH:
    w->velInVisc = w->vel;
    run(N, K::cld_applyXsphViscosityCorrection, pred, table, (const float4*)w->velInVisc.data(), fp, w->vel.data());
    break;
  case ST_UPDATE_POS: run(N, K::cld_updatePosition, pred, cp, w->pos.data()); break;
  case ST_RENDER_AUX:
  {
    run(w->C, K::resetGridDetector, w->partDetector.data());
    run(N, K::fillGridDetector, w->pos.data(), w->partDetector.data());
    size_t b;
    const float* q = (const float*)ref_field_ptr(w, w->dispField, &b);
    if (q && b == 4 * w->M)
      run(N, K::fillColorFloat, q, w->dispMin, w->dispMax, w->col.data());
    break;
  }
  case ST_CAMERA_SORT:
    run(N, K::fillCameraDist, pos, (const ocl::float3*)w->cam, w->cameraDist.data());
    sortAndPermute(w, w->cameraDist, w->cameraPerm, true);
    break;
  default: return -1;
  }
  return 0;
}

extern "C" int ref_run_stage(ref_world* w, int stage)
{
  ocl::g_opt = w->opt;
  switch (w->model)
  {
  case RTP_MODEL_BOIDS: return stageBoids(w, stage);
  case RTP_MODEL_FLUIDS: return stageFluids(w, stage);
  case RTP_MODEL_CLOUDS: return stageClouds(w, stage);
  }
  return -1;
}

// {Boids,Fluids,Clouds}::update() (Boids.cpp:323-384, Fluids.cpp:400-471, Clouds.cpp:503-627)
extern "C" int ref_step(ref_world* w, unsigned flags, const float cam[3])
{
  ref_set_camera(w, cam);
  auto st = [&](int s) { ref_run_stage(w, s); };
  if (flags & RTP_STEP_PHYSICS)
  {
    if (w->model == RTP_MODEL_BOIDS)
    {
      st(ST_FILL_CELL_IDS); st(ST_SORT_BY_CELL); st(ST_BUILD_CELL_TABLE); st(ST_BD_RULES); st(ST_BD_TARGET); st(ST_BD_UPDATE_VEL); st(ST_BD_UPDATE_POS);
    }
    else if (w->model == RTP_MODEL_FLUIDS)
    {
      st(ST_PREDICT_POS); st(ST_FILL_CELL_IDS); st(ST_SORT_BY_CELL); st(ST_BUILD_CELL_TABLE);
      for (int it = 0; it < w->jacobi; ++it)
      {
        st(ST_APPLY_BOUNDARY); st(ST_DENSITY); st(ST_CONSTRAINT_FACTOR); st(ST_CONSTRAINT_CORRECTION); st(ST_CORRECT_POS);
      }
      st(ST_UPDATE_VEL);
      if (w->fluid.isVorticityConfEnabled)
      {
        st(ST_VORTICITY); st(ST_VORTICITY_CONFINEMENT); st(ST_XSPH);
      }
      st(ST_UPDATE_POS);
    }
    else
    {
      st(ST_CLD_THERMO); st(ST_PREDICT_POS); st(ST_APPLY_BOUNDARY); st(ST_FILL_CELL_IDS); st(ST_SORT_BY_CELL); st(ST_BUILD_CELL_TABLE);
      if (w->cloud.isTempSmoothingEnabled)
      {
        st(ST_CLD_LAPLACIAN_TEMP); st(ST_CLD_CONSTRAINT_FACTOR_TEMP); st(ST_CLD_CONSTRAINT_CORRECTION_TEMP); st(ST_CLD_CORRECT_TEMP);
      }
      for (int it = 0; it < w->jacobi; ++it)
      {
        st(ST_DENSITY); st(ST_CONSTRAINT_FACTOR); st(ST_CONSTRAINT_CORRECTION); st(ST_CORRECT_POS); st(ST_APPLY_BOUNDARY);
      }
      st(ST_UPDATE_VEL);
      if (w->fluid.isVorticityConfEnabled)
      {
        st(ST_VORTICITY); st(ST_VORTICITY_CONFINEMENT); st(ST_XSPH);
      }
      st(ST_UPDATE_POS);
    }
  }
  if ((flags & RTP_STEP_RENDER_AUX) && ((flags & RTP_STEP_PHYSICS) || w->model == RTP_MODEL_CLOUDS))
    st(ST_RENDER_AUX);
  if (flags & RTP_STEP_CAMERA_SORT)
    st(ST_CAMERA_SORT);
  return 0;
}
