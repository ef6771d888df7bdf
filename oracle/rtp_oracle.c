/*
 * rtp_oracle.c -- CPU restatement of the RealTimeParticles hot path (parity oracle + CPU baseline).
 * TEST INFRASTRUCTURE ONLY -- see rtp_oracle.h for who may use it and for the parity-pin statement.
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference). One function
 * == one reference kernel; expression order follows the OpenCL C source left to right so that the result is
 * comparable bit for bit with oracle/_ref (the reference's .cl sources compiled through the shim).
 *
 * Canonical arithmetic (OpenCL leaves built-in precision to the driver and the reference builds with
 * -cl-fast-relaxed-math, i.e. mad contraction allowed; DESIGN.md "Canonical arithmetic" lists every choice):
 *   all arithmetic IEEE-754 binary32, round-to-nearest; the ONLY fused operations are the explicit fmaf() below
 *   (build with -ffp-contract=off). Element-wise stages follow the .cl expression order literally.
 *   dot(a,b)      = fma(a.z,b.z, fma(a.y,b.y, a.x*b.x))   (w is always 0 on this path)
 *   fast_length   = length = sqrtf(dot(v,v))
 *   fast_normalize(v) = v * (1.0f / sqrtf(dot(v,v))), fast_normalize(0) = 0   (OpenCL 1.2 s6.12.5: "if all elements
 *                   of x are zero, returns x"; tested as dot(v,v) == 0)
 *   normalize(v)  = v / sqrtf(dot(v,v)), normalize(0) = 0
 *   step(e,x)     = x < e ? 0 : 1 ; clamp(x,lo,hi) = fmin(fmax(x,lo),hi) ; convert_uint = truncation
 *   exp           = canon_expf / canon_exp below: Cody-Waite reduction by ln2 (hi+lo) and a Taylor-Horner polynomial
 *                   (degree 7 in float, 13 in double: < 1.5 ulp), every step an explicit fma; double is used in
 *                   saturationVaporDensity whose literals are double (clouds.cl:98)
 *  Pair terms inside the 27-cell sums (sph.cl kernels), with sq = dot(vec,vec), len = sqrtf(sq):
 *   poly6(vec)    = (len < h) ? POLY6_COEFF * ((t*t)*t) : 0,  t = h*h - sq          (len^2 taken as sq)
 *   gradSpiky(vec)= vec * c,  c = (len <= FLOAT_EPS || !(len < h)) ? 0 : ((K * (hl*hl)) * (1.0f/len)),
 *                   hl = h - len, K = SPIKY_COEFF * -3.0f
 *   x / d inside a pair term = x * (1.0f/d) with one IEEE reciprocal
 *   sum += a*b    = fmaf(a, b, sum) where noted at each kernel; pow(r, n) = r*r*...*r (n-1 sequential products)
 *  Sums run in the reference's order: 27 cells (iX, iY, iZ ascending), e ascending, one fp32 accumulator.
 */
#include "rtp_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct
{
  float x, y, z, w;
} f4;
typedef struct
{
  uint32_t x, y;
} u2;
typedef struct
{
  int x, y, z;
} i3;

#define FLOAT_EPS 0.00000001f /* define.cl:6 */
#define ABS_GRAVITY_ACC_Y 9.81f /* define.cl:8 */
#define FAR_DIST 1000000.0f /* define.cl:10 */
#define MAX_STEERING 0.5f /* boids.cl:10 */

struct orc_world
{
  int model;
  size_t M, N, C;
  uint32_t box[3];
  int res[3];
  int dim, boundary, jacobi;
  uint32_t maxPartsInCell;
  /* baked -D constants (Boids.cpp:103-113, Fluids.cpp:104-119, Clouds.cpp:132-147) */
  float absW[3], cellSize, effectRadius, effectRadiusSq, poly6Coeff, spikyCoeff, maxVel;
  rtp_boids_params boids;
  rtp_target_params target;
  float targetPos[4];
  int targetActive;
  rtp_fluid_params fluid;
  rtp_cloud_params cloud;
  float cam[3];
  int dispField;
  float dispMin, dispMax;
  /* buffers */
  f4 *pos, *col, *vel, *acc, *predPos, *corrPos, *velInVisc, *vort, *totCorrPos, *tmp4;
  float *density, *constFactor, *temp, *tempIn, *lapTemp, *corrTemp, *constFactorTemp;
  float *vaporDens, *vaporDensIn, *cloudDens, *cloudDensIn, *buoyancy, *cloudGen, *partID, *tmp1;
  float* partDetector; /* float8[C] */
  uint32_t *cellID, *cameraDist, *perm, *cameraPerm, *keysTmp, *idxTmp;
  u2* startEnd;
};

/* ------------------------------------------------------------------ helpers */

/* utils/Utils.cpp:24-29 FloatToStr (fixed, 10 decimals, 'f' suffix) then the OpenCL compiler parses the literal */
float orc_baked_constant(float v)
{
  char buf[128];
  snprintf(buf, sizeof buf, "%.10f", (double)v);
  return strtof(buf, NULL);
}

static inline f4 mk4(float x, float y, float z, float w)
{
  f4 r = { x, y, z, w };
  return r;
}
static inline f4 add4(f4 a, f4 b) { return mk4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
static inline f4 sub4(f4 a, f4 b) { return mk4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
static inline f4 mul4s(f4 a, float s) { return mk4(a.x * s, a.y * s, a.z * s, a.w * s); }
static inline f4 div4s(f4 a, float s) { return mk4(a.x / s, a.y / s, a.z / s, a.w / s); }
/* a * s + c with one rounding per component */
static inline f4 fma4s(f4 a, float s, f4 c) { return mk4(fmaf(a.x, s, c.x), fmaf(a.y, s, c.y), fmaf(a.z, s, c.z), fmaf(a.w, s, c.w)); }
static inline float dotc(f4 a, f4 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
static inline float lengthc(f4 v) { return sqrtf(dotc(v, v)); }
static inline f4 fast_normalizec(f4 v)
{
  const float d = dotc(v, v);
  if (d == 0.0f)
    return v;
  const float r = 1.0f / sqrtf(d);
  return mul4s(v, r);
}
static inline f4 normalizec(f4 v)
{
  const float l = sqrtf(dotc(v, v));
  if (l == 0.0f)
    return mk4(0.0f, 0.0f, 0.0f, 0.0f);
  return div4s(v, l);
}
static inline f4 crossc(f4 a, f4 b)
{
  return mk4(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x, 0.0f);
}
/* cross product with one fused multiply-add per component (pair terms) */
static inline f4 crossf(f4 a, f4 b)
{
  return mk4(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)), 0.0f);
}
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

/* canonical exp (see header): exp(x) = 2^n * P(r), n = rint(x log2 e), r = x - n ln2 (two-step), P = Taylor-Horner */
static inline float canon_expf(float x)
{
  x = fminf(fmaxf(x, -87.0f), 88.0f);
  const float n = rintf(x * 1.44269504f);
  float r = fmaf(n, -0.693145751953125f, x);
  r = fmaf(n, -1.428606765330187e-06f, r);
  float p = 1.98412698e-4f;
  p = fmaf(p, r, 1.38888889e-3f);
  p = fmaf(p, r, 8.33333333e-3f);
  p = fmaf(p, r, 4.16666667e-2f);
  p = fmaf(p, r, 1.66666667e-1f);
  p = fmaf(p, r, 0.5f);
  p = fmaf(p, r, 1.0f);
  p = fmaf(p, r, 1.0f);
  union { uint32_t u; float f; } s;
  s.u = (uint32_t)((int)n + 127) << 23;
  return p * s.f;
}
static inline double canon_exp(double x)
{
  x = fmin(fmax(x, -700.0), 700.0);
  const double n = rint(x * 1.4426950408889634);
  double r = fma(n, -6.93147180369123816490e-01, x);
  r = fma(n, -1.90821492927058770002e-10, r);
  double p = 1.0 / 6227020800.0;
  p = fma(p, r, 1.0 / 479001600.0);
  p = fma(p, r, 1.0 / 39916800.0);
  p = fma(p, r, 1.0 / 3628800.0);
  p = fma(p, r, 1.0 / 362880.0);
  p = fma(p, r, 1.0 / 40320.0);
  p = fma(p, r, 1.0 / 5040.0);
  p = fma(p, r, 1.0 / 720.0);
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  union { uint64_t u; double f; } s;
  s.u = (uint64_t)((int64_t)n + 1023) << 52;
  return p * s.f;
}

/* grid.cl:14-24 getCell3DIndexFromPos */
static inline i3 cell3D(const orc_world* w, f4 p)
{
  const float px = clampf(p.x, -w->absW[0], w->absW[0]) + w->absW[0];
  const float py = clampf(p.y, -w->absW[1], w->absW[1]) + w->absW[1];
  const float pz = clampf(p.z, -w->absW[2], w->absW[2]) + w->absW[2];
  i3 c;
  c.x = (int)(uint32_t)floorf(px / w->cellSize);
  c.y = (int)(uint32_t)floorf(py / w->cellSize);
  c.z = (int)(uint32_t)floorf(pz / w->cellSize);
  return c;
}
/* grid.cl:29-38 getCell1DIndexFromPos */
static inline uint32_t cell1D(const orc_world* w, f4 p)
{
  const i3 c = cell3D(w, p);
  return (uint32_t)c.x * (uint32_t)w->res[2] * (uint32_t)w->res[1] + (uint32_t)c.y * (uint32_t)w->res[2] + (uint32_t)c.z;
}

/* sph.cl:10-14 poly6 on the squared distance (canonical: len^2 == sq); coefficient applied by the caller */
static inline float poly6nc(const orc_world* w, float sq)
{
  const float h = w->effectRadius;
  const float len = sqrtf(sq);
  if (!(len < h))
    return 0.0f;
  const float t = h * h - sq;
  return (t * t) * t;
}
static inline float poly6(const orc_world* w, float sq) { return w->poly6Coeff * poly6nc(w, sq); }
/* sph.cl:26-34 gradSpiky = vec * spikyCoef(sq) */
static inline float spikyCoef(const orc_world* w, float sq)
{
  const float h = w->effectRadius;
  const float len = sqrtf(sq);
  if (len <= FLOAT_EPS || !(len < h))
    return 0.0f;
  const float hl = h - len;
  const float K = w->spikyCoeff * -3.0f;
  return (K * (hl * hl)) * (1.0f / len);
}
/* 1 / (poly6L(artPressureRadius * EFFECT_RADIUS) / POLY6_COEFF), fluids.cl:56 (the coefficient cancels in the ratio) */
static inline float artInvDenominator(const orc_world* w)
{
  const float h = w->effectRadius;
  const float len = w->fluid.artPressureRadius * h;
  const float t = h * h - len * len;
  const float den = (len < h) ? (t * t) * t : 0.0f;
  return 1.0f / den;
}
/* fluids.cl:51-57 / clouds.cl:105-111 artPressure = -k * (W(vec)/W(dq h))^n */
static inline float artPressure(const orc_world* w, float sq, float invDen)
{
  const rtp_fluid_params* f = &w->fluid;
  if (f->isArtPressureEnabled == 0)
    return 0.0f;
  const float ratio = poly6nc(w, sq) * invDen;
  float pw = ratio;
  for (uint32_t q = 1; q < f->artPressureExp; ++q)
    pw = pw * ratio;
  return -(f->artPressureCoeff * pw);
}

/* Neighbour-cell resolution for the three traversal flavours.
 *  fluids: modulo wrap, fluids.cl:107 (range test :110 is dead code after the modulo)
 *  boids : out-of-range cells skipped, boids.cl:84-88
 *  clouds: modulo wrap in x/z with a +-2W image shift, y out of range skipped, clouds.cl:334-347
 * Returns 0 when the cell is skipped. shift = absWall * signAbsWall (exact: 0 or +-2W). */
static inline int neighbour_cell(const orc_world* w, i3 ci, int iX, int iY, int iZ, uint32_t* c1, f4* shift)
{
  const int RX = w->res[0], RY = w->res[1], RZ = w->res[2];
  int cx, cy, cz;
  *shift = mk4(0.0f, 0.0f, 0.0f, 0.0f);
  if (w->model == RTP_MODEL_BOIDS)
  {
    cx = ci.x + iX;
    cy = ci.y + iY;
    cz = ci.z + iZ;
    if (cx < 0 || cy < 0 || cz < 0 || cx >= RX || cy >= RY || cz >= RZ)
      return 0;
  }
  else
  {
    cx = (ci.x + iX + RX) % RX;
    cy = (ci.y + iY + RY) % RY;
    cz = (ci.z + iZ + RZ) % RZ;
    if (w->model == RTP_MODEL_CLOUDS)
    {
      float sx = 0.0f, sz = 0.0f;
      if ((ci.x + iX) >= RX)
        sx = 2.0f;
      else if ((ci.x + iX) < 0)
        sx = -2.0f;
      if (((ci.y + iY) >= RY) || ((ci.y + iY) < 0))
        return 0;
      if ((ci.z + iZ) >= RZ)
        sz = 2.0f;
      else if ((ci.z + iZ) < 0)
        sz = -2.0f;
      *shift = mk4(w->absW[0] * sx, w->absW[1] * 0.0f, w->absW[2] * sz, 0.0f);
    }
  }
  *c1 = (uint32_t)((cx * RY + cy) * RZ + cz);
  return 1;
}

/* 27-cell traversal in the reference's order (iX, iY, iZ ascending; e ascending, inclusive range) */
#define FOR_EACH_NEIGHBOUR(W, CI, E, SHIFT, ...)                          \
  for (int iX_ = -1; iX_ <= 1; ++iX_)                                     \
    for (int iY_ = -1; iY_ <= 1; ++iY_)                                   \
      for (int iZ_ = -1; iZ_ <= 1; ++iZ_)                                 \
      {                                                                   \
        uint32_t c1_;                                                     \
        f4 SHIFT;                                                         \
        if (!neighbour_cell((W), (CI), iX_, iY_, iZ_, &c1_, &SHIFT))      \
          continue;                                                       \
        const u2 se_ = (W)->startEnd[c1_];                                \
        for (uint32_t E = se_.x; E <= se_.y; ++E)                         \
        {                                                                 \
          __VA_ARGS__                                                     \
        }                                                                 \
      }

/* pos - posN - absWall*sign  (clouds.cl:356); for boids/fluids the shift is +0 and is not applied */
static inline f4 pair_vec(const orc_world* w, f4 pos, f4 posN, f4 shift)
{
  f4 v = sub4(pos, posN);
  if (w->model == RTP_MODEL_CLOUDS)
    v = sub4(v, shift);
  return v;
}

/* ------------------------------------------------------------------ life cycle */

static void* zalloc(size_t n, size_t sz)
{
  void* p = calloc(n ? n : 1, sz);
  return p;
}

int orc_create(const rtp_config* cfg, orc_world** out)
{
  if (!cfg || !out || cfg->max_particles == 0 || cfg->nb_particles > cfg->max_particles)
    return RTP_ERR_INVALID;
  for (int k = 0; k < 3; ++k)
    if (cfg->box[k] == 0 || cfg->grid[k] == 0)
      return RTP_ERR_INVALID;
  orc_world* w = (orc_world*)calloc(1, sizeof *w);
  w->model = cfg->model;
  w->M = cfg->max_particles;
  w->N = cfg->nb_particles;
  for (int k = 0; k < 3; ++k)
  {
    w->box[k] = cfg->box[k];
    w->res[k] = (int)cfg->grid[k];
  }
  w->C = (size_t)cfg->grid[0] * cfg->grid[1] * cfg->grid[2];
  w->dim = cfg->dim == 2 ? 2 : 3;
  w->boundary = RTP_BOUNDARY_BOUNCING_WALL;
  w->jacobi = 2; /* Fluids.cpp:58, Clouds.cpp:72 */
  w->maxPartsInCell = cfg->max_parts_in_cell ? cfg->max_parts_in_cell : (cfg->model == RTP_MODEL_BOIDS ? 3000u : 100u);

  /* Fluids.cpp:104-119 / Boids.cpp:103-113 / Clouds.cpp:132-147 */
  const float effectRadius = ((float)w->box[0]) / (float)w->res[0];
  w->effectRadius = orc_baked_constant(effectRadius);
  w->cellSize = orc_baked_constant((float)w->box[0] / (float)w->res[0]);
  for (int k = 0; k < 3; ++k)
    w->absW[k] = orc_baked_constant((float)w->box[k] / 2.0f);
  w->effectRadiusSq = orc_baked_constant(1.0f * (float)w->box[0] * (float)w->box[0] / (float)((size_t)w->res[0] * (size_t)w->res[0]));
  const float PI_F = 3.1415927f; /* utils/Math.hpp:69 */
  w->poly6Coeff = orc_baked_constant(315.0f / (64.0f * PI_F * powf(effectRadius, 9.f)));
  w->spikyCoeff = orc_baked_constant(15.0f / (PI_F * powf(effectRadius, 6.f)));
  w->maxVel = orc_baked_constant(30.0f);

  /* defaults: Boids.hpp:14-26, Fluids.hpp:17-32, Clouds.hpp:15-45 (+ Clouds.cpp:308-311 copies) */
  w->boids = (rtp_boids_params) { 0.5f, 1.6f, 1.6f, 1.45f };
  w->target = (rtp_target_params) { 2.0f, 1 };
  w->targetActive = 0;
  w->fluid = (rtp_fluid_params) { 450.0f, 600.0f, 0.010f, (uint32_t)w->dim, 1, 0.006f, 0.001f, 4, 1, 0.0004f, 0.0001f };
  w->cloud = (rtp_cloud_params) { (uint32_t)w->dim, 0.01f, 450.0f, 10.0f, 0.10f, 0.0005f, 5.0f, 0.3485f, 0.07f, 1, 600.0f, 0.75f, 1.0f };
  w->cam[0] = 32.0f; /* render/Camera.cpp:11 */
  w->cam[1] = -1.2f;
  w->cam[2] = 0.0f;
  w->dispField = RTP_F_CLOUD_DENS; /* Clouds.cpp:213 */
  w->dispMin = 1.0f;
  w->dispMax = 15.0f;

  const size_t M = w->M;
  w->pos = zalloc(M, sizeof(f4));
  w->col = zalloc(M, sizeof(f4));
  w->vel = zalloc(M, sizeof(f4));
  w->acc = zalloc(M, sizeof(f4));
  w->predPos = zalloc(M, sizeof(f4));
  w->corrPos = zalloc(M, sizeof(f4));
  w->velInVisc = zalloc(M, sizeof(f4));
  w->vort = zalloc(M, sizeof(f4));
  w->totCorrPos = zalloc(M, sizeof(f4));
  w->tmp4 = zalloc(M, sizeof(f4));
  float** f1[] = { &w->density, &w->constFactor, &w->temp, &w->tempIn, &w->lapTemp, &w->corrTemp, &w->constFactorTemp,
    &w->vaporDens, &w->vaporDensIn, &w->cloudDens, &w->cloudDensIn, &w->buoyancy, &w->cloudGen, &w->partID, &w->tmp1 };
  for (size_t i = 0; i < sizeof f1 / sizeof f1[0]; ++i)
    *f1[i] = zalloc(M, sizeof(float));
  w->partDetector = zalloc(w->C * 8, sizeof(float));
  w->cellID = zalloc(M + 1, sizeof(uint32_t)); /* +1: fillEndCell reads cellID[N] even when N == M (grid.cl:131) */
  w->cameraDist = zalloc(M, sizeof(uint32_t));
  w->perm = zalloc(M, sizeof(uint32_t));
  w->cameraPerm = zalloc(M, sizeof(uint32_t));
  w->keysTmp = zalloc(M, sizeof(uint32_t));
  w->idxTmp = zalloc(M, sizeof(uint32_t));
  w->startEnd = zalloc(w->C, sizeof(u2));
  w->cellID[w->M] = 0xFFFFFFFFu;
  *out = w;
  return RTP_OK;
}

void orc_destroy(orc_world* w)
{
  if (!w)
    return;
  void* ptrs[] = { w->pos, w->col, w->vel, w->acc, w->predPos, w->corrPos, w->velInVisc, w->vort, w->totCorrPos, w->tmp4,
    w->density, w->constFactor, w->temp, w->tempIn, w->lapTemp, w->corrTemp, w->constFactorTemp, w->vaporDens, w->vaporDensIn,
    w->cloudDens, w->cloudDensIn, w->buoyancy, w->cloudGen, w->partID, w->tmp1, w->partDetector, w->cellID, w->cameraDist,
    w->perm, w->cameraPerm, w->keysTmp, w->idxTmp, w->startEnd };
  for (size_t i = 0; i < sizeof ptrs / sizeof ptrs[0]; ++i)
    free(ptrs[i]);
  free(w);
}

void* orc_field_ptr(orc_world* w, int field, size_t* bytes)
{
  const size_t M = w->M;
  void* p = NULL;
  size_t b = 0;
  switch (field)
  {
  case RTP_F_POS: p = w->pos; b = 16 * M; break;
  case RTP_F_COL: p = w->col; b = 16 * M; break;
  case RTP_F_VEL: p = w->vel; b = 16 * M; break;
  case RTP_F_ACC: p = w->acc; b = 16 * M; break;
  case RTP_F_PRED_POS: p = w->predPos; b = 16 * M; break;
  case RTP_F_CORR_POS: p = w->corrPos; b = 16 * M; break;
  case RTP_F_VORT: p = w->vort; b = 16 * M; break;
  case RTP_F_TOT_CORR_POS: p = w->totCorrPos; b = 16 * M; break;
  case RTP_F_DENSITY: p = w->density; b = 4 * M; break;
  case RTP_F_CONST_FACTOR: p = w->constFactor; b = 4 * M; break;
  case RTP_F_TEMP: p = w->temp; b = 4 * M; break;
  case RTP_F_VAPOR_DENS: p = w->vaporDens; b = 4 * M; break;
  case RTP_F_CLOUD_DENS: p = w->cloudDens; b = 4 * M; break;
  case RTP_F_BUOYANCY: p = w->buoyancy; b = 4 * M; break;
  case RTP_F_CLOUD_GEN: p = w->cloudGen; b = 4 * M; break;
  case RTP_F_PART_ID: p = w->partID; b = 4 * M; break;
  case RTP_F_LAPLACIAN_TEMP: p = w->lapTemp; b = 4 * M; break;
  case RTP_F_CONST_FACTOR_TEMP: p = w->constFactorTemp; b = 4 * M; break;
  case RTP_F_CORR_TEMP: p = w->corrTemp; b = 4 * M; break;
  case RTP_F_CELL_ID: p = w->cellID; b = 4 * M; break;
  case RTP_F_CAMERA_DIST: p = w->cameraDist; b = 4 * M; break;
  case RTP_F_START_END_CELL: p = w->startEnd; b = 8 * w->C; break;
  case RTP_F_PERM: p = w->perm; b = 4 * M; break;
  case RTP_F_CAMERA_PERM: p = w->cameraPerm; b = 4 * M; break;
  case RTP_F_PART_DETECTOR: p = w->partDetector; b = 32 * w->C; break;
  default: break;
  }
  if (bytes)
    *bytes = b;
  return p;
}

int orc_set_boids_params(orc_world* w, const rtp_boids_params* rules, const rtp_target_params* target,
    const float target_pos[4], int target_active)
{
  if (rules)
    w->boids = *rules;
  if (target)
    w->target = *target;
  if (target_pos)
    memcpy(w->targetPos, target_pos, sizeof w->targetPos);
  w->targetActive = target_active;
  return RTP_OK;
}
int orc_set_fluid_params(orc_world* w, const rtp_fluid_params* fluid, int nb_jacobi_iters)
{
  if (fluid)
    w->fluid = *fluid;
  if (nb_jacobi_iters > 0)
    w->jacobi = nb_jacobi_iters;
  return RTP_OK;
}
int orc_set_cloud_params(orc_world* w, const rtp_cloud_params* cloud)
{
  if (cloud)
    w->cloud = *cloud;
  return RTP_OK;
}
int orc_set_boundary(orc_world* w, int boundary)
{
  w->boundary = boundary;
  return RTP_OK;
}
int orc_set_nb_particles(orc_world* w, uint64_t n)
{
  if (n > w->M)
    return RTP_ERR_INVALID;
  w->N = n;
  return RTP_OK;
}
int orc_set_dimension(orc_world* w, int dim)
{
  w->dim = dim == 2 ? 2 : 3;
  w->fluid.dim = (uint32_t)w->dim;
  w->cloud.dim = (uint32_t)w->dim;
  return RTP_OK;
}
int orc_set_displayed_quantity(orc_world* w, int field, float min_val, float max_val)
{
  w->dispField = field;
  w->dispMin = min_val;
  w->dispMax = max_val;
  return RTP_OK;
}
void orc_set_camera(orc_world* w, const float cam[3])
{
  if (cam)
    memcpy(w->cam, cam, sizeof w->cam);
}

float orc_constant(const orc_world* w, const char* name)
{
  if (!strcmp(name, "EFFECT_RADIUS")) return w->effectRadius;
  if (!strcmp(name, "EFFECT_RADIUS_SQUARED")) return w->effectRadiusSq;
  if (!strcmp(name, "GRID_CELL_SIZE_XYZ")) return w->cellSize;
  if (!strcmp(name, "ABS_WALL_X")) return w->absW[0];
  if (!strcmp(name, "ABS_WALL_Y")) return w->absW[1];
  if (!strcmp(name, "ABS_WALL_Z")) return w->absW[2];
  if (!strcmp(name, "POLY6_COEFF")) return w->poly6Coeff;
  if (!strcmp(name, "SPIKY_COEFF")) return w->spikyCoeff;
  if (!strcmp(name, "MAX_VEL")) return w->maxVel;
  return NAN;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_threads(int n)
{
#ifdef _OPENMP
  if (n > 0)
    omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------ sort (RadixSort.cpp:122-190) */

/* Semantics of RadixSort::sort's key phase: stable ascending LSD sort of n 32-bit keys in 4 passes of 8 bits
 * (RadixSort.cpp:25-32, :136-161), carrying the permutation initialised by resetIndex (radixSort.cl:171-174). */
void orc_sort_keys(const uint32_t* keys_in, uint32_t* keys_out, uint32_t* perm_out, uint64_t n)
{
  uint32_t* ka = (uint32_t*)malloc((n ? n : 1) * 4);
  uint32_t* kb = (uint32_t*)malloc((n ? n : 1) * 4);
  uint32_t* ia = (uint32_t*)malloc((n ? n : 1) * 4);
  uint32_t* ib = (uint32_t*)malloc((n ? n : 1) * 4);
  memcpy(ka, keys_in, n * 4);
  for (uint64_t i = 0; i < n; ++i)
    ia[i] = (uint32_t)i;
  for (int pass = 0; pass < 4; ++pass)
  {
    uint64_t hist[257];
    memset(hist, 0, sizeof hist);
    for (uint64_t i = 0; i < n; ++i)
      ++hist[((ka[i] >> (pass * 8)) & 255u) + 1];
    for (int d = 0; d < 256; ++d)
      hist[d + 1] += hist[d];
    for (uint64_t i = 0; i < n; ++i)
    {
      const uint64_t o = hist[(ka[i] >> (pass * 8)) & 255u]++;
      kb[o] = ka[i];
      ib[o] = ia[i];
    }
    uint32_t* t = ka;
    ka = kb;
    kb = t;
    t = ia;
    ia = ib;
    ib = t;
  }
  if (keys_out)
    memcpy(keys_out, ka, n * 4);
  if (perm_out)
    memcpy(perm_out, ia, n * 4);
  free(ka);
  free(kb);
  free(ia);
  free(ib);
}

/* permutateFloat4 / permutateFloat after a full copy into a temp (RadixSort.cpp:164-189, radixSort.cl:179-202) */
static void gather4(orc_world* w, f4* buf, const uint32_t* perm)
{
  const size_t M = w->M;
  memcpy(w->tmp4, buf, M * sizeof(f4));
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < M; ++i)
    buf[i] = w->tmp4[perm[i]];
}
static void gather1(orc_world* w, float* buf, const uint32_t* perm)
{
  const size_t M = w->M;
  memcpy(w->tmp1, buf, M * sizeof(float));
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < M; ++i)
    buf[i] = w->tmp1[perm[i]];
}

/* RadixSort::sort(key, float4 list, float list) over all M entries (RadixSort.cpp:122-190) */
static void sort_and_permute(orc_world* w, uint32_t* keys, uint32_t* perm, int cameraSort)
{
  orc_sort_keys(keys, w->keysTmp, perm, w->M);
  memcpy(keys, w->keysTmp, w->M * 4);
  switch (w->model)
  {
  case RTP_MODEL_BOIDS: /* Boids.cpp:337, :381 */
    gather4(w, w->pos, perm);
    gather4(w, w->col, perm);
    gather4(w, w->vel, perm);
    gather4(w, w->acc, perm);
    break;
  case RTP_MODEL_FLUIDS: /* Fluids.cpp:417, :468 */
    gather4(w, w->pos, perm);
    gather4(w, w->col, perm);
    gather4(w, w->vel, perm);
    gather4(w, w->predPos, perm);
    break;
  case RTP_MODEL_CLOUDS: /* Clouds.cpp:543, :624 */
    gather4(w, w->pos, perm);
    gather4(w, w->col, perm);
    gather4(w, w->vel, perm);
    gather4(w, w->predPos, perm);
    if (!cameraSort)
      gather4(w, w->totCorrPos, perm);
    gather1(w, w->temp, perm);
    gather1(w, w->buoyancy, perm);
    gather1(w, w->vaporDens, perm);
    gather1(w, w->cloudDens, perm);
    gather1(w, w->partID, perm);
    break;
  }
}

/* ------------------------------------------------------------------ grid.cl */

/* resetCellIDs grid.cl:65-71 + resetCameraDist utils.cl:35-38 (Fluids.cpp:214-215) */
int orc_reset_ids(orc_world* w)
{
  for (size_t i = 0; i < w->M; ++i)
  {
    w->cellID[i] = (uint32_t)(w->C * 2 + i);
    w->cameraDist[i] = (uint32_t)(FAR_DIST);
  }
  w->cellID[w->M] = 0xFFFFFFFFu; /* out-of-bounds sentinel, never a valid id */
  return RTP_OK;
}

/* fillCellIDs grid.cl:76-86 */
static void k_fillCellIDs(orc_world* w, const f4* p)
{
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < w->N; ++i)
    w->cellID[i] = cell1D(w, p[i]);
}

/* resetStartEndCell :91-96, fillStartCell :101-117, fillEndCell :122-138, adjustEndCell :143-152 */
static void k_buildCellTable(orc_world* w)
{
  const uint32_t C = (uint32_t)w->C;
  for (size_t c = 0; c < w->C; ++c)
  {
    w->startEnd[c].x = 1;
    w->startEnd[c].y = 0;
  }
  for (size_t i = 0; i < w->N; ++i)
  {
    const uint32_t id = w->cellID[i];
    if (i > 0 && id < C)
    {
      if (id != w->cellID[i - 1])
        w->startEnd[id].x = (uint32_t)i;
    }
  }
  for (size_t i = 0; i < w->N; ++i)
  {
    const uint32_t id = w->cellID[i];
    if (id < C) /* "ID != get_global_size(0)" is always true (grid.cl:129) */
    {
      if (id != w->cellID[i + 1])
        w->startEnd[id].y = (uint32_t)i;
    }
  }
  for (size_t c = 0; c < w->C; ++c)
  {
    const u2 se = w->startEnd[c];
    if (se.y > se.x)
    {
      const uint32_t d = se.y - se.x;
      w->startEnd[c].y = se.x + (d < w->maxPartsInCell ? d : w->maxPartsInCell);
    }
  }
}

/* resetGridDetector / fillGridDetector grid.cl:43-60 */
static void k_gridDetector(orc_world* w)
{
  memset(w->partDetector, 0, w->C * 8 * sizeof(float));
  for (size_t i = 0; i < w->N; ++i)
  {
    const uint32_t c = cell1D(w, w->pos[i]);
    if (c < w->C)
      for (int k = 0; k < 8; ++k)
        w->partDetector[(size_t)c * 8 + k] = 1.0f;
  }
}

/* fillCameraDist utils.cl:43-52 */
static void k_fillCameraDist(orc_world* w)
{
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < w->N; ++i)
  {
    const f4 d = mk4(w->pos[i].x - w->cam[0], w->pos[i].y - w->cam[1], w->pos[i].z - w->cam[2], 0.0f);
    const float len = sqrtf(dotc(d, d));
    w->cameraDist[i] = (uint32_t)(fmaxf(FAR_DIST - len * 100.0f, 0.0f));
  }
}

/* ------------------------------------------------------------------ boids.cl */

/* bd_applyBoidsRulesWithGrid3D boids.cl:46-132 */
static void k_bd_rules3D(orc_world* w)
{
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t i = 0; i < w->N; ++i)
  {
    const f4 pos = w->pos[i];
    const i3 ci = cell3D(w, pos);
    int count = 0;
    f4 newAcc = mk4(0, 0, 0, 0), avgPos = mk4(0, 0, 0, 0), avgVel = mk4(0, 0, 0, 0), repulse = mk4(0, 0, 0, 0);
    FOR_EACH_NEIGHBOUR(w, ci, e, shift, {
      (void)shift;
      const f4 posN = w->pos[e];
      const f4 vec = sub4(pos, posN);
      const float sq = dotc(vec, vec);
      if (sq < w->effectRadiusSq && sq > FLOAT_EPS)
      {
        avgPos = add4(avgPos, posN);
        avgVel = add4(avgVel, fast_normalizec(w->vel[e]));
        repulse = fma4s(vec, 1.0f / sq, repulse); /* vec / squaredDist, boids.cl:105 */
        ++count;
      }
    })
    if (count != 0)
    {
      const rtp_boids_params* p = &w->boids;
      avgPos = div4s(avgPos, (float)count);
      avgPos = sub4(avgPos, pos);
      avgPos = mul4s(fast_normalizec(avgPos), p->velocityScale);
      avgVel = mul4s(fast_normalizec(avgVel), p->velocityScale);
      repulse = mul4s(fast_normalizec(repulse), p->velocityScale);
      newAcc = add4(add4(mul4s(avgVel, p->alignmentScale), mul4s(repulse, p->separationScale)), mul4s(avgPos, p->cohesionScale));
    }
    w->acc[i] = newAcc;
  }
}

/* bd_applyBoidsRulesWithGrid2D boids.cl:137-221 (9 YZ cells, index formula :180, range test against GRID_RES_X :177) */
static void k_bd_rules2D(orc_world* w)
{
  const int RX = w->res[0], RY = w->res[1];
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t i = 0; i < w->N; ++i)
  {
    const f4 pos = w->pos[i];
    const i3 ci = cell3D(w, pos);
    int count = 0;
    f4 newAcc = mk4(0, 0, 0, 0), avgPos = mk4(0, 0, 0, 0), avgVel = mk4(0, 0, 0, 0), repulse = mk4(0, 0, 0, 0);
    for (int iY = -1; iY <= 1; ++iY)
      for (int iZ = -1; iZ <= 1; ++iZ)
      {
        const int cx = ci.x, cy = ci.y + iY, cz = ci.z + iZ;
        if (cx < 0 || cy < 0 || cz < 0 || cx >= RX || cy >= RX || cz >= RX)
          continue;
        const uint32_t c1 = (uint32_t)((RX / 2 * RX + cy) * RY + cz);
        const u2 se = w->startEnd[c1];
        for (uint32_t e = se.x; e <= se.y; ++e)
        {
          const f4 posN = w->pos[e];
          const f4 vec = sub4(pos, posN);
          const float sq = dotc(vec, vec);
          if (sq < w->effectRadiusSq && sq > FLOAT_EPS)
          {
            avgPos = add4(avgPos, posN);
            avgVel = add4(avgVel, fast_normalizec(w->vel[e]));
            repulse = fma4s(vec, 1.0f / sq, repulse);
            ++count;
          }
        }
      }
    if (count != 0)
    {
      const rtp_boids_params* p = &w->boids;
      avgPos = div4s(avgPos, (float)count);
      avgPos = sub4(avgPos, pos);
      avgPos = mul4s(fast_normalizec(avgPos), p->velocityScale);
      avgVel = mul4s(fast_normalizec(avgVel), p->velocityScale);
      repulse = mul4s(fast_normalizec(repulse), p->velocityScale);
      newAcc = add4(add4(mul4s(avgVel, p->alignmentScale), mul4s(repulse, p->separationScale)), mul4s(avgPos, p->cohesionScale));
    }
    w->acc[i] = newAcc;
  }
}

/* bd_addTargetRule boids.cl:226-241 */
static void k_bd_target(orc_world* w)
{
  const f4 t = mk4(w->targetPos[0], w->targetPos[1], w->targetPos[2], w->targetPos[3]);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < w->N; ++i)
  {
    const f4 vec = sub4(t, w->pos[i]);
    const float dist = lengthc(vec);
    if (dist < w->target.targetRadiusEffect)
    {
      const float s = clampf(1.3f / dist, 0.0f, 1.4f * MAX_STEERING);
      w->acc[i] = add4(w->acc[i], mul4s(mul4s(vec, (float)w->target.targetSignEffect), s));
    }
  }
}

/* bd_updateVel boids.cl:246-259 (timeStep 0.1 Boids.cpp:334, maxVelocity = velocityScale Boids.cpp:216) */
static void k_bd_updateVel(orc_world* w)
{
  const float dt = 0.1f, maxV = w->boids.velocityScale;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < w->N; ++i)
  {
    const f4 nv = add4(w->vel[i], mul4s(w->acc[i], dt));
    const float norm = clampf(lengthc(nv), 0.2f * maxV, maxV);
    w->vel[i] = mul4s(fast_normalizec(nv), norm);
  }
}

/* bd_updatePosAndApplyWallBC boids.cl:264-284 / bd_updatePosAndApplyPeriodicBC :289-315 */
static void k_bd_updatePos(orc_world* w)
{
  const float dt = 0.1f;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < w->N; ++i)
  {
    const f4 np = add4(w->pos[i], mul4s(w->vel[i], dt));
    f4 cp = mk4(clampf(np.x, -w->absW[0], w->absW[0]), clampf(np.y, -w->absW[1], w->absW[1]),
        clampf(np.z, -w->absW[2], w->absW[2]), clampf(np.w, 0.0f, 0.0f));
    if (w->boundary == RTP_BOUNDARY_CYCLIC_WALL)
    {
      if (!(cp.x == np.x)) cp.x *= -1.0f;
      if (!(cp.y == np.y)) cp.y *= -1.0f;
      if (!(cp.z == np.z)) cp.z *= -1.0f;
      w->pos[i] = cp;
    }
    else
    {
      w->pos[i] = cp;
      if (!(cp.x == np.x && cp.y == np.y && cp.z == np.z))
        w->vel[i] = mul4s(w->vel[i], -0.5f);
    }
  }
}

/* ------------------------------------------------------------------ fluids.cl / clouds.cl */

/* fld_predictPosition fluids.cl:62-74 ; cld_predictPosition clouds.cl:257-273 */
static void k_predictPosition(orc_world* w)
{
  if (w->model == RTP_MODEL_FLUIDS)
  {
    const float dt = w->fluid.timeStep;
    const f4 g = mul4s(mk4(0.0f, -ABS_GRAVITY_ACC_Y, 0.0f, 0.0f), dt);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < w->N; ++i)
    {
      const f4 nv = add4(w->vel[i], g);
      w->predPos[i] = add4(w->pos[i], mul4s(nv, dt));
    }
  }
  else
  {
    const float dt = w->cloud.timeStep;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < w->N; ++i)
    {
      const f4 pv = add4(w->vel[i], mul4s(mk4(0.0f, w->buoyancy[i], 0.0f, 0.0f), dt));
      w->totCorrPos[i] = mul4s(pv, dt);
      w->predPos[i] = add4(w->pos[i], mul4s(pv, dt));
    }
  }
}

/* fld_applyBoundaryCondition fluids.cl:435-439 ; cld_applyMixedBoundaryConditions clouds.cl:279-299 */
static void k_applyBoundary(orc_world* w)
{
  const float WX = w->absW[0], WY = w->absW[1], WZ = w->absW[2];
  if (w->model == RTP_MODEL_FLUIDS)
  {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < w->N; ++i)
    {
      f4 p = w->predPos[i];
      p.x = clampf(p.x, -WX + 0.01f, WX - 0.1f);
      p.y = clampf(p.y, -WY + 0.01f, WY - 0.1f);
      p.z = clampf(p.z, -WZ + 0.01f, WZ - 0.1f);
      p.w = clampf(p.w, 0.0f, 0.0f);
      w->predPos[i] = p;
    }
  }
  else
  {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < w->N; ++i)
    {
      const f4 np = w->predPos[i];
      const f4 cp = mk4(clampf(np.x, -WX, WX), clampf(np.y, -WY, WY), clampf(np.z, -WZ, WZ), 0.0f);
      f4 p = np;
      if (fabsf(np.x) > WX) p.x = np.x - 2 * cp.x;
      if (fabsf(np.y) > WY) p.y = cp.y;
      if (fabsf(np.z) > WZ) p.z = np.z - 2 * cp.z;
      w->predPos[i] = p;
    }
  }
}

/* fld_computeDensity fluids.cl:80-126 ; cld_computeDensity clouds.cl:305-360 */
static void k_density(orc_world* w)
{
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t i = 0; i < w->N; ++i)
  {
    const f4 pos = w->predPos[i];
    const i3 ci = cell3D(w, pos);
    float d = 0.0f;
    FOR_EACH_NEIGHBOUR(w, ci, e, shift, {
      const f4 vec = pair_vec(w, pos, w->predPos[e], shift);
      d += poly6(w, dotc(vec, vec));
    })
    w->density[i] = d;
  }
}

/* fld_computeConstraintFactor fluids.cl:131-193 ; cld_computeConstraintFactor clouds.cl:365-436 */
static void k_constraintFactor(orc_world* w)
{
  const float rho0 = w->fluid.restDensity;
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t i = 0; i < w->N; ++i)
  {
    const f4 pos = w->predPos[i];
    const i3 ci = cell3D(w, pos);
    const float densityC = w->density[i] / rho0 - 1.0f;
    f4 sumGradCi = mk4(0, 0, 0, 0);
    float sumSqGradC = 0.0f;
    FOR_EACH_NEIGHBOUR(w, ci, e, shift, {
      const f4 vec = pair_vec(w, pos, w->predPos[e], shift);
      const float sq = dotc(vec, vec);
      const float c = spikyCoef(w, sq);
      sumGradCi = fma4s(vec, c, sumGradCi); /* sumGradCi += grad */
      sumSqGradC += (c * c) * sq; /* dot(grad, grad) */
    })
    sumSqGradC += dotc(sumGradCi, sumGradCi);
    sumSqGradC /= rho0 * rho0;
    w->constFactor[i] = -densityC / (sumSqGradC + w->fluid.relaxCFM);
  }
}

/* fld_computeConstraintCorrection fluids.cl:198-247 ; cld_computeConstraintCorrection clouds.cl:441-502 */
static void k_constraintCorrection(orc_world* w)
{
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t i = 0; i < w->N; ++i)
  {
    const f4 pos = w->predPos[i];
    const float lambdaI = w->constFactor[i];
    const i3 ci = cell3D(w, pos);
    f4 corr = mk4(0, 0, 0, 0);
    const float invDen = artInvDenominator(w);
    FOR_EACH_NEIGHBOUR(w, ci, e, shift, {
      const f4 vec = pair_vec(w, pos, w->predPos[e], shift);
      const float sq = dotc(vec, vec);
      const float s = lambdaI + w->constFactor[e] + artPressure(w, sq, invDen);
      corr = fma4s(vec, s * spikyCoef(w, sq), corr);
    })
    w->corrPos[i] = div4s(corr, w->fluid.restDensity);
  }
}

/* fld_correctPosition fluids.cl:252-258 ; clouds: cld_correctPosition twice, Clouds.cpp:579-583 */
static void k_correctPosition(orc_world* w)
{
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < w->N; ++i)
  {
    w->predPos[i] = add4(w->predPos[i], w->corrPos[i]);
    if (w->model == RTP_MODEL_CLOUDS)
      w->totCorrPos[i] = add4(w->totCorrPos[i], w->corrPos[i]);
  }
}

/* fld_updateVel fluids.cl:263-273 ; cld_updateVel clouds.cl:958-967 */
static void k_updateVel(orc_world* w)
{
  const float dt = w->fluid.timeStep, mv = w->maxVel;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < w->N; ++i)
  {
    const f4 d = (w->model == RTP_MODEL_FLUIDS) ? sub4(w->predPos[i], w->pos[i]) : w->totCorrPos[i];
    const f4 v = div4s(d, dt + FLOAT_EPS);
    w->vel[i] = mk4(clampf(v.x, -mv, mv), clampf(v.y, -mv, mv), clampf(v.z, -mv, mv), clampf(v.w, -mv, mv));
  }
}

/* fld_computeVorticity fluids.cl:278-324 ; cld_computeVorticity clouds.cl:727-785 */
static void k_vorticity(orc_world* w)
{
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t i = 0; i < w->N; ++i)
  {
    const f4 pos = w->predPos[i];
    const f4 velocity = w->vel[i];
    const i3 ci = cell3D(w, pos);
    f4 vort = mk4(0, 0, 0, 0);
    FOR_EACH_NEIGHBOUR(w, ci, e, shift, {
      const f4 vec = pair_vec(w, pos, w->predPos[e], shift);
      const float c = spikyCoef(w, dotc(vec, vec));
      /* cross(dv, vec * c) = cross(dv, vec) * c with cross(a,b).x = fma(a.y, b.z, -(a.z * b.y)) */
      vort = fma4s(crossf(sub4(w->vel[e], velocity), vec), c, vort);
    })
    w->vort[i] = vort;
  }
}

/* fld_applyVorticityConfinement fluids.cl:329-377 ; cld_applyVorticityConfinement clouds.cl:790-850 */
static void k_vorticityConfinement(orc_world* w)
{
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t i = 0; i < w->N; ++i)
  {
    const f4 pos = w->predPos[i];
    const f4 vorticity = w->vort[i];
    const i3 ci = cell3D(w, pos);
    f4 n = mk4(0, 0, 0, 0);
    FOR_EACH_NEIGHBOUR(w, ci, e, shift, {
      const f4 vec = pair_vec(w, pos, w->predPos[e], shift);
      n = fma4s(vec, lengthc(w->vort[e]) * spikyCoef(w, dotc(vec, vec)), n);
    })
    const f4 c = crossc(normalizec(n), vorticity);
    w->vel[i] = add4(w->vel[i], mul4s(mul4s(c, w->fluid.vorticityConfCoeff), w->fluid.timeStep));
  }
}

/* copyBuffer(p_vel -> p_velInViscosity) Fluids.cpp:451 + fld_applyXsphViscosityCorrection fluids.cl:383-430 ;
 * cld_applyXsphViscosityCorrection clouds.cl:856-915 */
static void k_xsph(orc_world* w)
{
  memcpy(w->velInVisc, w->vel, w->M * sizeof(f4));
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t i = 0; i < w->N; ++i)
  {
    const f4 pos = w->predPos[i];
    const f4 velocity = w->velInVisc[i];
    const i3 ci = cell3D(w, pos);
    f4 visc = mk4(0, 0, 0, 0);
    FOR_EACH_NEIGHBOUR(w, ci, e, shift, {
      const f4 vec = pair_vec(w, pos, w->predPos[e], shift);
      visc = fma4s(sub4(w->velInVisc[e], velocity), poly6(w, dotc(vec, vec)), visc);
    })
    w->vel[i] = add4(velocity, mul4s(visc, w->fluid.xsphViscosityCoeff));
  }
}

/* fld_updatePosition fluids.cl:444-450 ; cld_updatePosition clouds.cl:942-953 */
static void k_updatePosition(orc_world* w)
{
  if (w->model == RTP_MODEL_FLUIDS)
  {
    memcpy(w->pos, w->predPos, w->N * sizeof(f4));
    return;
  }
  const rtp_cloud_params* c = &w->cloud;
  const float WY = w->absW[1];
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < w->N; ++i)
  {
    const f4 pp = w->predPos[i];
    f4 p = pp;
    p.x += (1 - canon_expf(-(pp.y + WY) * 0.2f)) * c->windCoeff * c->timeStep * (float)(c->dim - 2);
    p.z += (1 - canon_expf(-(pp.y + WY) * 0.3f)) * 0.7f * c->windCoeff * c->timeStep;
    w->pos[i] = p;
  }
}

/* fld_fillFluidColor fluids.cl:458-479 */
static void k_fillFluidColor(orc_world* w)
{
  const f4 blue = mk4(0.0f, 0.1f, 1.0f, 0.5f), lightBlue = mk4(0.7f, 0.7f, 1.0f, 0.5f), darkBlue = mk4(0.0f, 0.0f, 0.8f, 0.5f);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < w->N; ++i)
  {
    const float constraint = (1.0f - w->density[i] / w->fluid.restDensity);
    f4 color = blue;
    if (constraint > 0.0f)
      color = add4(color, div4s(mul4s(sub4(lightBlue, blue), constraint), 0.35f));
    else if (constraint < 0.0f)
      color = add4(color, div4s(mul4s(sub4(blue, darkBlue), constraint), 0.35f));
    w->col[i] = color;
  }
}

/* fillColorFloat utils.cl:65-78 */
static void k_fillColorFloat(orc_world* w)
{
  size_t bytes;
  const float* q = (const float*)orc_field_ptr(w, w->dispField, &bytes);
  if (!q || bytes != 4 * w->M)
    return;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < w->N; ++i)
  {
    float val = (q[i] - w->dispMin) / (w->dispMax - w->dispMin);
    val *= (val < 0.0f) ? 0.0f : 1.0f; /* step(0, val) */
    val *= (1.0f < val) ? 0.0f : 1.0f; /* step(val, 1) */
    w->col[i] = mk4(val, val, val, val);
  }
}

/* clouds.cl:72-99 */
static inline float externalHeatSource(const orc_world* w, float alt) { return clampf(canon_expf(-(alt + w->absW[1]) / 3.0f), 0.0f, 1.0f); }
static inline float environmentTemp(const orc_world* w, float alt) { return -3.5f * (alt + w->absW[1]) + 293.0f; }
static inline float saturationVaporDensity(float T) { return (float)(217 * canon_exp(19.5 - 4303.4 / ((double)T - 29.5)) / (double)T); }

/* cld_initTemperature clouds.cl:116-122 + cld_initVaporDensity :127-135, both over M (Clouds.cpp:495-497) */
int orc_init_clouds_fields(orc_world* w)
{
  if (w->model != RTP_MODEL_CLOUDS)
    return RTP_ERR_STATE;
  for (size_t i = 0; i < w->M; ++i)
  {
    w->temp[i] = environmentTemp(w, w->pos[i].y);
    w->vaporDens[i] = w->cloud.initVaporDensityCoeff * saturationVaporDensity(w->temp[i]);
  }
  return RTP_OK;
}

/* Clouds.cpp:514-529: copyBuffer(temp->tempIn); cld_heatFromGround :159-168; cld_computeBuoyancy :173-184;
 * cld_applyAdiabaticCooling :189-198 (writes p_tempIn); cld_generateCloud :207-221; copies; cld_applyPhaseTransition
 * :226-238; cld_applyLatentHeat :243-252 */
static void k_cloudsThermo(orc_world* w)
{
  const rtp_cloud_params* c = &w->cloud;
  memcpy(w->tempIn, w->temp, w->M * sizeof(float));
  memcpy(w->vaporDensIn, w->vaporDens, w->M * sizeof(float));
  memcpy(w->cloudDensIn, w->cloudDens, w->M * sizeof(float));
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < w->N; ++i)
  {
    const float y = w->pos[i].y;
    /* heatFromGround */
    float temp = fminf(w->tempIn[i] + externalHeatSource(w, y) * c->groundHeatCoeff * c->timeStep, 313.0f);
    /* buoyancy */
    const float envTemp = environmentTemp(w, y);
    w->buoyancy[i] = c->buoyancyCoeff * (temp - envTemp) / envTemp - c->gravCoeff * ABS_GRAVITY_ACC_Y * w->cloudDens[i];
    /* adiabatic cooling -> p_tempIn */
    const float tempIn = fmaxf(temp - c->adiabaticLapseRate * w->vel[i].y * c->timeStep, 223.0f);
    w->tempIn[i] = tempIn;
    /* cloud generation */
    const float gen = c->phaseTransitionRate * (w->vaporDens[i] - saturationVaporDensity(tempIn));
    w->cloudGen[i] = gen;
    /* phase transition */
    w->cloudDens[i] = fmaxf(w->cloudDensIn[i] + gen * c->timeStep, 0.0f);
    w->vaporDens[i] = fmaxf(w->vaporDensIn[i] - gen * c->timeStep, 0.0f);
    /* latent heat */
    temp = tempIn + fmaxf(c->latentHeatCoeff * gen * c->timeStep, 0.0f);
    w->temp[i] = temp;
  }
}

/* cld_computeLaplacianTemp clouds.cl:508-569 (on p_pos, Clouds.cpp:253) */
static void k_laplacianTemp(orc_world* w)
{
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t i = 0; i < w->N; ++i)
  {
    const f4 pos = w->pos[i];
    const float temp = w->temp[i];
    const i3 ci = cell3D(w, pos);
    float lap = 0.0f;
    FOR_EACH_NEIGHBOUR(w, ci, e, shift, {
      const f4 vec = pair_vec(w, pos, w->pos[e], shift);
      const float sq = dotc(vec, vec);
      /* dot(vec, grad) = c * sq ; x / d = x * (1/d) */
      lap = fmaf((temp - w->temp[e]) * (spikyCoef(w, sq) * sq), 1.0f / (sq + FLOAT_EPS), lap);
    })
    w->lapTemp[i] = lap / w->cloud.restDensity;
  }
}

/* cld_computeConstraintFactorTemp clouds.cl:575-648 */
static void k_constraintFactorTemp(orc_world* w)
{
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t i = 0; i < w->N; ++i)
  {
    const f4 pos = w->pos[i];
    const float lap = w->lapTemp[i];
    const i3 ci = cell3D(w, pos);
    float sumGradCi = 0.0f, sumSqGradC = 0.0f;
    FOR_EACH_NEIGHBOUR(w, ci, e, shift, {
      const f4 vec = pair_vec(w, pos, w->pos[e], shift);
      const float sq = dotc(vec, vec);
      const float dT = (spikyCoef(w, sq) * sq) * (1.0f / fmaf(sq, w->cloud.restDensity, FLOAT_EPS));
      sumGradCi += dT;
      sumSqGradC = fmaf(dT, dT, sumSqGradC);
    })
    sumSqGradC += sumGradCi * sumGradCi;
    w->constFactorTemp[i] = -lap / (sumSqGradC + w->cloud.relaxCFM);
  }
}

/* cld_computeConstraintCorrectionTemp clouds.cl:654-722 */
static void k_constraintCorrectionTemp(orc_world* w)
{
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t i = 0; i < w->N; ++i)
  {
    const f4 pos = w->pos[i];
    const float lambdaI = w->constFactorTemp[i];
    const i3 ci = cell3D(w, pos);
    float corr = 0.0f;
    FOR_EACH_NEIGHBOUR(w, ci, e, shift, {
      const f4 vec = pair_vec(w, pos, w->pos[e], shift);
      const float sq = dotc(vec, vec);
      const float dT = (spikyCoef(w, sq) * sq) * (1.0f / fmaf(sq, w->cloud.restDensity, FLOAT_EPS));
      corr = fmaf(lambdaI + w->constFactorTemp[e], dT, corr);
    })
    w->corrTemp[i] = corr;
  }
}

/* cld_correctTemperature clouds.cl:931-937 */
static void k_correctTemperature(orc_world* w)
{
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < w->N; ++i)
    w->temp[i] += 0.3f * w->corrTemp[i];
}

/* ------------------------------------------------------------------ stage dispatch and full steps */

int orc_run_stage(orc_world* w, int stage)
{
  const int fluidLike = w->model != RTP_MODEL_BOIDS;
  switch (stage)
  {
  case ORC_FILL_CELL_IDS: k_fillCellIDs(w, fluidLike ? w->predPos : w->pos); break;
  case ORC_SORT_BY_CELL: sort_and_permute(w, w->cellID, w->perm, 0); break;
  case ORC_BUILD_CELL_TABLE: k_buildCellTable(w); break;
  case ORC_BD_RULES: if (w->dim == 2) k_bd_rules2D(w); else k_bd_rules3D(w); break;
  case ORC_BD_TARGET: if (w->targetActive) k_bd_target(w); break;
  case ORC_BD_UPDATE_VEL: k_bd_updateVel(w); break;
  case ORC_BD_UPDATE_POS: k_bd_updatePos(w); break;
  case ORC_PREDICT_POS: k_predictPosition(w); break;
  case ORC_APPLY_BOUNDARY: k_applyBoundary(w); break;
  case ORC_DENSITY: k_density(w); break;
  case ORC_CONSTRAINT_FACTOR: k_constraintFactor(w); break;
  case ORC_CONSTRAINT_CORRECTION: k_constraintCorrection(w); break;
  case ORC_CORRECT_POS: k_correctPosition(w); break;
  case ORC_UPDATE_VEL: k_updateVel(w); break;
  case ORC_VORTICITY: k_vorticity(w); break;
  case ORC_VORTICITY_CONFINEMENT: k_vorticityConfinement(w); break;
  case ORC_XSPH: k_xsph(w); break;
  case ORC_UPDATE_POS: k_updatePosition(w); break;
  case ORC_CLD_THERMO: k_cloudsThermo(w); break;
  case ORC_CLD_LAPLACIAN_TEMP: k_laplacianTemp(w); break;
  case ORC_CLD_CONSTRAINT_FACTOR_TEMP: k_constraintFactorTemp(w); break;
  case ORC_CLD_CONSTRAINT_CORRECTION_TEMP: k_constraintCorrectionTemp(w); break;
  case ORC_CLD_CORRECT_TEMP: k_correctTemperature(w); break;
  case ORC_RENDER_AUX:
    k_gridDetector(w);
    if (w->model == RTP_MODEL_FLUIDS)
      k_fillFluidColor(w);
    else if (w->model == RTP_MODEL_CLOUDS)
      k_fillColorFloat(w);
    break;
  case ORC_CAMERA_SORT:
    k_fillCameraDist(w);
    sort_and_permute(w, w->cameraDist, w->cameraPerm, 1);
    break;
  default: return RTP_ERR_INVALID;
  }
  return RTP_OK;
}

/* Boids::update Boids.cpp:323-384 */
static void step_boids(orc_world* w)
{
  orc_run_stage(w, ORC_FILL_CELL_IDS);
  orc_run_stage(w, ORC_SORT_BY_CELL);
  orc_run_stage(w, ORC_BUILD_CELL_TABLE);
  orc_run_stage(w, ORC_BD_RULES);
  orc_run_stage(w, ORC_BD_TARGET);
  orc_run_stage(w, ORC_BD_UPDATE_VEL);
  orc_run_stage(w, ORC_BD_UPDATE_POS);
}

/* Fluids::update Fluids.cpp:400-471 */
static void step_fluids(orc_world* w)
{
  orc_run_stage(w, ORC_PREDICT_POS);
  orc_run_stage(w, ORC_FILL_CELL_IDS);
  orc_run_stage(w, ORC_SORT_BY_CELL);
  orc_run_stage(w, ORC_BUILD_CELL_TABLE);
  for (int it = 0; it < w->jacobi; ++it)
  {
    orc_run_stage(w, ORC_APPLY_BOUNDARY);
    orc_run_stage(w, ORC_DENSITY);
    orc_run_stage(w, ORC_CONSTRAINT_FACTOR);
    orc_run_stage(w, ORC_CONSTRAINT_CORRECTION);
    orc_run_stage(w, ORC_CORRECT_POS);
  }
  orc_run_stage(w, ORC_UPDATE_VEL);
  if (w->fluid.isVorticityConfEnabled)
  {
    orc_run_stage(w, ORC_VORTICITY);
    orc_run_stage(w, ORC_VORTICITY_CONFINEMENT);
    orc_run_stage(w, ORC_XSPH);
  }
  orc_run_stage(w, ORC_UPDATE_POS);
}

/* Clouds::update Clouds.cpp:503-627 */
static void step_clouds(orc_world* w)
{
  orc_run_stage(w, ORC_CLD_THERMO);
  orc_run_stage(w, ORC_PREDICT_POS);
  orc_run_stage(w, ORC_APPLY_BOUNDARY);
  orc_run_stage(w, ORC_FILL_CELL_IDS);
  orc_run_stage(w, ORC_SORT_BY_CELL);
  orc_run_stage(w, ORC_BUILD_CELL_TABLE);
  if (w->cloud.isTempSmoothingEnabled)
  {
    orc_run_stage(w, ORC_CLD_LAPLACIAN_TEMP);
    orc_run_stage(w, ORC_CLD_CONSTRAINT_FACTOR_TEMP);
    orc_run_stage(w, ORC_CLD_CONSTRAINT_CORRECTION_TEMP);
    orc_run_stage(w, ORC_CLD_CORRECT_TEMP);
  }
  for (int it = 0; it < w->jacobi; ++it)
  {
    orc_run_stage(w, ORC_DENSITY);
    orc_run_stage(w, ORC_CONSTRAINT_FACTOR);
    orc_run_stage(w, ORC_CONSTRAINT_CORRECTION);
    orc_run_stage(w, ORC_CORRECT_POS);
    orc_run_stage(w, ORC_APPLY_BOUNDARY);
  }
  orc_run_stage(w, ORC_UPDATE_VEL);
  if (w->fluid.isVorticityConfEnabled)
  {
    orc_run_stage(w, ORC_VORTICITY);
    orc_run_stage(w, ORC_VORTICITY_CONFINEMENT);
    orc_run_stage(w, ORC_XSPH);
  }
  orc_run_stage(w, ORC_UPDATE_POS);
}

int orc_step(orc_world* w, unsigned flags, const float cam[3])
{
  orc_set_camera(w, cam);
  if (flags & RTP_STEP_PHYSICS)
  {
    switch (w->model)
    {
    case RTP_MODEL_BOIDS: step_boids(w); break;
    case RTP_MODEL_FLUIDS: step_fluids(w); break;
    case RTP_MODEL_CLOUDS: step_clouds(w); break;
    default: return RTP_ERR_INVALID;
    }
    if (flags & RTP_STEP_RENDER_AUX)
    {
      k_gridDetector(w);
      if (w->model == RTP_MODEL_FLUIDS)
        k_fillFluidColor(w);
    }
  }
  /* clouds colouring runs even on pause (Clouds.cpp:610-617) */
  if ((flags & RTP_STEP_RENDER_AUX) && w->model == RTP_MODEL_CLOUDS)
    k_fillColorFloat(w);
  if (flags & RTP_STEP_CAMERA_SORT)
    orc_run_stage(w, ORC_CAMERA_SORT);
  return RTP_OK;
}

/* ------------------------------------------------------------------ initial conditions (utils/Geometry.cpp) */

/* GenerateBoxGrid, Distribution::Uniform, utils/Geometry.cpp:198-227 */
int64_t orc_gen_box_grid(float* out, const int res[3], const float start[3], const float end[3])
{
  const float vx = end[0] - start[0], vy = end[1] - start[1], vz = end[2] - start[2];
  const float sx = vx / res[0], sy = vy / res[1], sz = vz / res[2];
  int64_t n = 0;
  for (int ix = 0; ix < res[0]; ++ix)
    for (int iy = 0; iy < res[1]; ++iy)
      for (int iz = 0; iz < res[2]; ++iz)
      {
        out[4 * n + 0] = start[0] + ix * sx;
        out[4 * n + 1] = start[1] + iy * sy;
        out[4 * n + 2] = start[2] + iz * sz;
        out[4 * n + 3] = 0.0f;
        ++n;
      }
  return n;
}

/* GenerateSphereGrid utils/Geometry.cpp:243-272 (Math::length utils/Math.hpp:1590 = sqrt(dot)) */
int64_t orc_gen_sphere_grid(float* out, const int res[3], const float start[3], const float end[3])
{
  const float PI_F = 3.1415927f;
  const float vx = end[0] - start[0], vy = end[1] - start[1], vz = end[2] - start[2];
  const float cx = start[0] + vx / 2.0f, cy = start[1] + vy / 2.0f, cz = start[2] + vz / 2.0f;
  const float radius = sqrtf(vx * vx + vy * vy + vz * vz) / 2.0f;
  const float phiSpacing = PI_F / res[0];
  const float thetaSpacing = 2.0f * PI_F / res[1];
  const float radiusSpacing = radius / res[2];
  int64_t n = 0;
  for (int iphi = 0; iphi < res[0]; ++iphi)
    for (int itheta = 0; itheta < res[1]; ++itheta)
      for (int ir = 0; ir < res[2]; ++ir)
      {
        out[4 * n + 0] = cx + ((ir + 1) * radiusSpacing) * cosf(itheta * thetaSpacing) * sinf(iphi * phiSpacing);
        out[4 * n + 1] = cy + ((ir + 1) * radiusSpacing) * sinf(itheta * thetaSpacing) * sinf(iphi * phiSpacing);
        out[4 * n + 2] = cz + ((ir + 1) * radiusSpacing) * cosf(iphi * phiSpacing);
        out[4 * n + 3] = 0.0f;
        ++n;
      }
  return n;
}

/* GenerateBoxGrid, Distribution::Random, utils/Geometry.cpp:229-239: glibc rand() in x, y, z call order */
int64_t orc_gen_random_box(float* out, int64_t n, const float start[3], const float end[3], int seed)
{
  const float vx = end[0] - start[0], vy = end[1] - start[1], vz = end[2] - start[2];
  if (seed >= 0)
    srand((unsigned)seed);
  for (int64_t i = 0; i < n; ++i)
  {
    out[4 * i + 0] = (float)rand() / (float)RAND_MAX * vx + start[0];
    out[4 * i + 1] = (float)rand() / (float)RAND_MAX * vy + start[1];
    out[4 * i + 2] = (float)rand() / (float)RAND_MAX * vz + start[2];
    out[4 * i + 3] = 0.0f;
  }
  return n;
}
