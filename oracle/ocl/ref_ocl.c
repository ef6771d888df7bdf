/*
 * ref_ocl.c -- the reference's UNMODIFIED OpenCL C kernels (physics/ocl/kernels/{define,sph,fluids,utils,grid}.cl) run on
 * the GPU box's own OpenCL device (NVIDIA's libnvidia-opencl.so.1 behind the CUDA toolkit's ICD loader), with the
 * reference's build options (Context.cpp:242-285: "-cl-denorms-are-zero -cl-fast-relaxed-math" + the -D constants of
 * Fluids.cpp:104-119). TEST INFRASTRUCTURE ONLY: a third checker beside oracle/rtp_oracle.c (restatement) and oracle/_ref/
 * libref_kernels.so (the same .cl files compiled for the CPU through a shim) -- it is the only one in which the OpenCL
 * BUILT-INS (fast_length, pow, step, normalize ...) and the relaxed-math contraction are a real driver's, not ours.
 *
 * What is restated here is only the HOST side of the fluids model: which kernel runs when with which buffers
 * (Fluids::createKernels Fluids.cpp:149-193, Fluids::update :400-457, physics part). RadixSort::sort (physics/utils/
 * RadixSort.cpp:122-190) is replaced by its contract, a stable sort by key + gather of {p_pos, p_vel, p_predPos}, done on
 * the host: the per-kernel times reported by rocl_kernel_times() therefore cover the model kernels, not the sort.
 *
 * No OpenCL headers exist in this image: the few entry points are declared by hand and resolved with dlopen(). The kernel
 * sources are embedded at BUILD time from where they lie under /root/reference (oracle/ocl/Makefile generates a temporary
 * header and deletes it); nothing of the reference is stored in this repository. Output: oracle/_ref/libref_ocl.so.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/rtp_cuda.h"
#include "ref_cl_sources.h" /* generated: static const char* const ref_cl_sources[5] */

typedef void* cl_obj;
typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef uint64_t cl_ulong;
#define CL_DEVICE_TYPE_GPU (1ul << 2)
#define CL_MEM_READ_WRITE (1ul << 0)
#define CL_QUEUE_PROFILING_ENABLE (1ul << 1)
#define CL_PROGRAM_BUILD_LOG 0x1183
#define CL_PROFILING_COMMAND_START 0x1282
#define CL_PROFILING_COMMAND_END 0x1283
#define CL_DEVICE_NAME 0x102B

static struct
{
  void* lib;
  cl_int (*GetPlatformIDs)(cl_uint, cl_obj*, cl_uint*);
  cl_int (*GetDeviceIDs)(cl_obj, cl_ulong, cl_uint, cl_obj*, cl_uint*);
  cl_int (*GetDeviceInfo)(cl_obj, cl_uint, size_t, void*, size_t*);
  cl_obj (*CreateContext)(const intptr_t*, cl_uint, const cl_obj*, void*, void*, cl_int*);
  cl_obj (*CreateCommandQueue)(cl_obj, cl_obj, cl_ulong, cl_int*);
  cl_obj (*CreateProgramWithSource)(cl_obj, cl_uint, const char**, const size_t*, cl_int*);
  cl_int (*BuildProgram)(cl_obj, cl_uint, const cl_obj*, const char*, void*, void*);
  cl_int (*GetProgramBuildInfo)(cl_obj, cl_obj, cl_uint, size_t, void*, size_t*);
  cl_obj (*CreateKernel)(cl_obj, const char*, cl_int*);
  cl_int (*SetKernelArg)(cl_obj, cl_uint, size_t, const void*);
  cl_obj (*CreateBuffer)(cl_obj, cl_ulong, size_t, void*, cl_int*);
  cl_int (*EnqueueWriteBuffer)(cl_obj, cl_obj, cl_uint, size_t, size_t, const void*, cl_uint, const cl_obj*, cl_obj*);
  cl_int (*EnqueueReadBuffer)(cl_obj, cl_obj, cl_uint, size_t, size_t, void*, cl_uint, const cl_obj*, cl_obj*);
  cl_int (*EnqueueCopyBuffer)(cl_obj, cl_obj, cl_obj, size_t, size_t, size_t, cl_uint, const cl_obj*, cl_obj*);
  cl_int (*EnqueueNDRangeKernel)(cl_obj, cl_obj, cl_uint, const size_t*, const size_t*, const size_t*, cl_uint, const cl_obj*, cl_obj*);
  cl_int (*Finish)(cl_obj);
  cl_int (*GetEventProfilingInfo)(cl_obj, cl_uint, size_t, void*, size_t*);
  cl_int (*ReleaseEvent)(cl_obj);
  cl_int (*ReleaseMemObject)(cl_obj);
  cl_int (*ReleaseKernel)(cl_obj);
  cl_int (*ReleaseProgram)(cl_obj);
  cl_int (*ReleaseCommandQueue)(cl_obj);
  cl_int (*ReleaseContext)(cl_obj);
} cl;

static int load_cl(char* err, size_t cap)
{
  if (cl.lib)
    return 0;
  /* the GPU boxes ship NVIDIA's OpenCL driver but no /etc/OpenCL/vendors entry: point the ICD loader at it */
  setenv("OCL_ICD_FILENAMES", "libnvidia-opencl.so.1", 0);
  const char* names[] = { "libOpenCL.so.1", "/usr/local/cuda/lib64/libOpenCL.so.1", "libOpenCL.so" };
  for (int k = 0; k < 3 && !cl.lib; ++k)
    cl.lib = dlopen(names[k], RTLD_NOW | RTLD_LOCAL);
  if (!cl.lib)
  {
    snprintf(err, cap, "no OpenCL ICD loader: %s", dlerror());
    return -1;
  }
#define SYM(f)                                         \
  if (!(*(void**)(&cl.f) = dlsym(cl.lib, "cl" #f)))    \
  {                                                    \
    snprintf(err, cap, "missing symbol cl" #f);        \
    return -1;                                         \
  }
  SYM(GetPlatformIDs) SYM(GetDeviceIDs) SYM(GetDeviceInfo) SYM(CreateContext) SYM(CreateCommandQueue) SYM(CreateProgramWithSource)
  SYM(BuildProgram) SYM(GetProgramBuildInfo) SYM(CreateKernel) SYM(SetKernelArg) SYM(CreateBuffer) SYM(EnqueueWriteBuffer)
  SYM(EnqueueReadBuffer) SYM(EnqueueCopyBuffer) SYM(EnqueueNDRangeKernel) SYM(Finish) SYM(GetEventProfilingInfo) SYM(ReleaseEvent)
  SYM(ReleaseMemObject) SYM(ReleaseKernel) SYM(ReleaseProgram) SYM(ReleaseCommandQueue) SYM(ReleaseContext)
#undef SYM
  return 0;
}

enum
{
  B_POS, B_VEL, B_PRED, B_CORR, B_VELVISC, B_VORT, B_COL, B_DENSITY, B_LAMBDA, B_CELLID, B_TABLE, B_COUNT
};
static const char* const kBufNames[B_COUNT] = { "p_pos", "p_vel", "p_predPos", "p_corrPos", "p_velInViscosity", "p_vort", "p_col", "p_density",
  "p_constFactor", "p_cellID", "c_startEndPartID" };
enum
{
  K_RESET_CELL_ID, K_FILL_CELL_ID, K_RESET_TABLE, K_FILL_START, K_FILL_END, K_ADJUST_END, K_PREDICT, K_BOUNDARY, K_DENSITY, K_LAMBDA,
  K_CORRECTION, K_CORRECT_POS, K_UPDATE_VEL, K_VORTICITY, K_CONFINEMENT, K_XSPH, K_UPDATE_POS, K_COUNT
};
/* kernel name, buffers bound at creation (Fluids.cpp:149-193; -1 ends the list, -2 = the FluidParams argument) */
static const struct
{
  const char* name;
  int args[6];
} kKernels[K_COUNT] = {
  { "resetCellIDs", { B_CELLID, -1 } },
  { "fillCellIDs", { B_PRED, B_CELLID, -1 } },
  { "resetStartEndCell", { B_TABLE, -1 } },
  { "fillStartCell", { B_CELLID, B_TABLE, -1 } },
  { "fillEndCell", { B_CELLID, B_TABLE, -1 } },
  { "adjustEndCell", { B_TABLE, -1 } },
  { "fld_predictPosition", { B_POS, B_VEL, -2, B_PRED, -1 } },
  { "fld_applyBoundaryCondition", { B_PRED, -1 } },
  { "fld_computeDensity", { B_PRED, B_TABLE, -2, B_DENSITY, -1 } },
  { "fld_computeConstraintFactor", { B_PRED, B_DENSITY, B_TABLE, -2, B_LAMBDA, -1 } },
  { "fld_computeConstraintCorrection", { B_LAMBDA, B_TABLE, B_PRED, -2, B_CORR, -1 } },
  { "fld_correctPosition", { B_CORR, B_PRED, -1 } },
  { "fld_updateVel", { B_PRED, B_POS, -2, B_VEL, -1 } },
  { "fld_computeVorticity", { B_PRED, B_TABLE, B_VEL, -2, B_VORT, -1 } },
  { "fld_applyVorticityConfinement", { B_PRED, B_TABLE, B_VORT, -2, B_VEL, -1 } },
  { "fld_applyXsphViscosityCorrection", { B_PRED, B_TABLE, B_VELVISC, -2, B_VEL, -1 } },
  { "fld_updatePosition", { B_PRED, B_POS, -1 } },
};

typedef struct
{
  cl_obj ctx, queue, program, dev;
  cl_obj buf[B_COUNT];
  size_t bytes[B_COUNT];
  cl_obj kernel[K_COUNT];
  double us[K_COUNT];
  unsigned launches[K_COUNT];
  unsigned M, N, cells;
  int jacobi;
  rtp_fluid_params fp;
  char devName[128];
  char err[512];
} rocl;

/* utils/Utils.cpp:24-29 FloatToStr: fixed notation, 10 decimals, 'f' suffix */
static void fstr(char* out, size_t cap, float v) { snprintf(out, cap, "%.10ff", (double)v); }

const char* rocl_last_error(rocl* r) { return r ? r->err : "null"; }
const char* rocl_device_name(rocl* r) { return r ? r->devName : ""; }

void rocl_destroy(rocl* r)
{
  if (!r)
    return;
  for (int k = 0; k < K_COUNT; ++k)
    if (r->kernel[k])
      cl.ReleaseKernel(r->kernel[k]);
  for (int b = 0; b < B_COUNT; ++b)
    if (r->buf[b])
      cl.ReleaseMemObject(r->buf[b]);
  if (r->program)
    cl.ReleaseProgram(r->program);
  if (r->queue)
    cl.ReleaseCommandQueue(r->queue);
  if (r->ctx)
    cl.ReleaseContext(r->ctx);
  free(r);
}

static int set_params(rocl* r)
{
  for (int k = 0; k < K_COUNT; ++k)
    for (int a = 0; a < 6 && kKernels[k].args[a] != -1; ++a)
      if (kKernels[k].args[a] == -2 && cl.SetKernelArg(r->kernel[k], (cl_uint)a, sizeof r->fp, &r->fp) != 0)
        return -1;
  return 0;
}

rocl* rocl_create(unsigned M, unsigned N, const unsigned box[3], const unsigned grid[3], int jacobi, char* err, size_t errcap)
{
  if (load_cl(err, errcap))
    return NULL;
  rocl* r = (rocl*)calloc(1, sizeof *r);
  cl_int e = 0;
  cl_obj plats[8];
  cl_uint np = 0, nd = 0;
#define FAIL(...)                      \
  do                                   \
  {                                    \
    snprintf(err, errcap, __VA_ARGS__); \
    rocl_destroy(r);                   \
    return NULL;                       \
  } while (0)
  if (cl.GetPlatformIDs(8, plats, &np) != 0 || np == 0)
    FAIL("no OpenCL platform (clGetPlatformIDs)");
  for (cl_uint p = 0; p < np && !r->dev; ++p)
    if (cl.GetDeviceIDs(plats[p], CL_DEVICE_TYPE_GPU, 1, &r->dev, &nd) != 0 || nd == 0)
      r->dev = NULL;
  if (!r->dev)
    FAIL("no OpenCL GPU device");
  cl.GetDeviceInfo(r->dev, CL_DEVICE_NAME, sizeof r->devName, r->devName, NULL);
  r->ctx = cl.CreateContext(NULL, 1, &r->dev, NULL, NULL, &e);
  if (!r->ctx || e)
    FAIL("clCreateContext: %d", e);
  r->queue = cl.CreateCommandQueue(r->ctx, r->dev, CL_QUEUE_PROFILING_ENABLE, &e);
  if (!r->queue || e)
    FAIL("clCreateCommandQueue: %d", e);
  r->M = M, r->N = N, r->jacobi = jacobi;
  r->cells = grid[0] * grid[1] * grid[2];
  r->fp = (rtp_fluid_params) { 450.0f, 600.0f, 0.010f, 3, 1, 0.006f, 0.001f, 4, 1, 0.0004f, 0.0001f }; /* Fluids.hpp:17-32 */

  /* Fluids::createProgram, Fluids.cpp:96-125 */
  const float h = (float)box[0] / grid[0];
  char o[16][64], opts[2048];
  fstr(o[0], 64, h);
  fstr(o[1], 64, box[0] / 2.0f);
  fstr(o[2], 64, box[1] / 2.0f);
  fstr(o[3], 64, box[2] / 2.0f);
  fstr(o[4], 64, (float)box[0] / grid[0]);
  const float PI_F = 3.1415927f; /* utils/Math.hpp */
  fstr(o[5], 64, 315.0f / (64.0f * PI_F * powf(h, 9.f)));
  fstr(o[6], 64, 15.0f / (PI_F * powf(h, 6.f)));
  fstr(o[7], 64, 30.0f);
  snprintf(opts, sizeof opts,
      "-DEFFECT_RADIUS=%s -DABS_WALL_X=%s -DABS_WALL_Y=%s -DABS_WALL_Z=%s -DGRID_RES_X=%u -DGRID_RES_Y=%u -DGRID_RES_Z=%u "
      "-DGRID_CELL_SIZE_XYZ=%s -DGRID_NUM_CELLS=%u -DNUM_MAX_PARTS_IN_CELL=%u -DPOLY6_COEFF=%s -DSPIKY_COEFF=%s -DMAX_VEL=%s"
      " -cl-denorms-are-zero -cl-fast-relaxed-math",
      o[0], o[1], o[2], o[3], grid[0], grid[1], grid[2], o[4], r->cells, 100u /* Fluids.cpp:79 */, o[5], o[6], o[7]);
  r->program = cl.CreateProgramWithSource(r->ctx, 5, (const char**)ref_cl_sources, NULL, &e);
  if (!r->program || e)
    FAIL("clCreateProgramWithSource: %d", e);
  if ((e = cl.BuildProgram(r->program, 1, &r->dev, opts, NULL, NULL)) != 0)
  {
    char log[1500] = { 0 };
    cl.GetProgramBuildInfo(r->program, r->dev, CL_PROGRAM_BUILD_LOG, sizeof log - 1, log, NULL);
    FAIL("clBuildProgram: %d\n%s", e, log);
  }
  /* Fluids::createBuffers, Fluids.cpp:127-147 (the GL-shared p_pos / p_col are plain buffers here); p_cellID has one extra
     word: fillEndCell reads cellID[i + 1] of the last particle (grid.cl:129-131), which the reference leaves out of bounds
     when N == M -- the oracle defines that word as "no valid id" */
  const size_t f4 = 16u * M, f1 = 4u * M;
  const size_t sizes[B_COUNT] = { f4, f4, f4, f4, f4, f4, f4, f1, f1, f1 + 4, 8u * r->cells };
  for (int b = 0; b < B_COUNT; ++b)
  {
    r->bytes[b] = sizes[b];
    r->buf[b] = cl.CreateBuffer(r->ctx, CL_MEM_READ_WRITE, sizes[b], NULL, &e);
    if (!r->buf[b] || e)
      FAIL("clCreateBuffer %s: %d", kBufNames[b], e);
    void* z = calloc(1, sizes[b]);
    cl.EnqueueWriteBuffer(r->queue, r->buf[b], 1, 0, sizes[b], z, 0, NULL, NULL);
    free(z);
  }
  const uint32_t sentinel = 0xFFFFFFFFu;
  cl.EnqueueWriteBuffer(r->queue, r->buf[B_CELLID], 1, f1, 4, &sentinel, 0, NULL, NULL);
  for (int k = 0; k < K_COUNT; ++k)
  {
    r->kernel[k] = cl.CreateKernel(r->program, kKernels[k].name, &e);
    if (!r->kernel[k] || e)
      FAIL("clCreateKernel %s: %d", kKernels[k].name, e);
    for (int a = 0; a < 6 && kKernels[k].args[a] != -1; ++a)
      if (kKernels[k].args[a] >= 0 && (e = cl.SetKernelArg(r->kernel[k], (cl_uint)a, sizeof(cl_obj), &r->buf[kKernels[k].args[a]])) != 0)
        FAIL("clSetKernelArg %s[%d]: %d", kKernels[k].name, a, e);
  }
  if (set_params(r))
    FAIL("clSetKernelArg(FluidParams)");
#undef FAIL
  return r;
}

static int find_buf(const char* name)
{
  for (int b = 0; b < B_COUNT; ++b)
    if (!strcmp(name, kBufNames[b]))
      return b;
  return -1;
}
int rocl_upload(rocl* r, const char* name, const void* host, size_t bytes)
{
  const int b = find_buf(name);
  if (!r || b < 0 || bytes > r->bytes[b])
    return -1;
  return cl.EnqueueWriteBuffer(r->queue, r->buf[b], 1, 0, bytes, host, 0, NULL, NULL);
}
int rocl_download(rocl* r, const char* name, void* host, size_t bytes)
{
  const int b = find_buf(name);
  if (!r || b < 0 || bytes > r->bytes[b])
    return -1;
  return cl.EnqueueReadBuffer(r->queue, r->buf[b], 1, 0, bytes, host, 0, NULL, NULL);
}
int rocl_set_fluid_params(rocl* r, const rtp_fluid_params* fp, int jacobi)
{
  if (!r || !fp)
    return -1;
  r->fp = *fp;
  if (jacobi > 0)
    r->jacobi = jacobi;
  return set_params(r);
}

static int run(rocl* r, int k, size_t n)
{
  if (!n)
    return 0;
  cl_obj ev = NULL;
  cl_int e = cl.EnqueueNDRangeKernel(r->queue, r->kernel[k], 1, NULL, &n, NULL, 0, NULL, &ev);
  if (e)
  {
    snprintf(r->err, sizeof r->err, "clEnqueueNDRangeKernel %s: %d", kKernels[k].name, e);
    return -1;
  }
  cl.Finish(r->queue); /* like the reference in profiling mode (Context.cpp:692-705) */
  cl_ulong t0 = 0, t1 = 0;
  cl.GetEventProfilingInfo(ev, CL_PROFILING_COMMAND_START, sizeof t0, &t0, NULL);
  cl.GetEventProfilingInfo(ev, CL_PROFILING_COMMAND_END, sizeof t1, &t1, NULL);
  cl.ReleaseEvent(ev);
  r->us[k] += (double)(t1 - t0) * 1e-3;
  r->launches[k]++;
  return 0;
}

int rocl_reset_ids(rocl* r) { return r ? run(r, K_RESET_CELL_ID, r->M) : -1; } /* Fluids::reset, Fluids.cpp:213 */

/* RadixSort::sort("p_cellID", {p_pos, p_col, p_vel, p_predPos}) by its contract: stable ascending, payload gathered */
typedef struct
{
  uint32_t key, idx;
} kv;
static int cmp_kv(const void* a, const void* b)
{
  const kv *x = (const kv*)a, *y = (const kv*)b;
  return x->key < y->key ? -1 : (x->key > y->key ? 1 : (x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0)));
}
static int sort_by_cell(rocl* r, uint32_t* permOut)
{
  const unsigned M = r->M;
  kv* a = (kv*)malloc(sizeof(kv) * M);
  uint32_t* keys = (uint32_t*)malloc(4u * M);
  float* in = (float*)malloc(16u * M);
  float* out = (float*)malloc(16u * M);
  cl.EnqueueReadBuffer(r->queue, r->buf[B_CELLID], 1, 0, 4u * M, keys, 0, NULL, NULL);
  for (unsigned i = 0; i < M; ++i)
    a[i].key = keys[i], a[i].idx = i;
  qsort(a, M, sizeof(kv), cmp_kv);
  for (unsigned i = 0; i < M; ++i)
    keys[i] = a[i].key;
  cl.EnqueueWriteBuffer(r->queue, r->buf[B_CELLID], 1, 0, 4u * M, keys, 0, NULL, NULL);
  const int payload[3] = { B_POS, B_VEL, B_PRED };
  for (int p = 0; p < 3; ++p)
  {
    cl.EnqueueReadBuffer(r->queue, r->buf[payload[p]], 1, 0, 16u * M, in, 0, NULL, NULL);
    for (unsigned i = 0; i < M; ++i)
      memcpy(out + 4u * i, in + 4u * a[i].idx, 16);
    cl.EnqueueWriteBuffer(r->queue, r->buf[payload[p]], 1, 0, 16u * M, out, 0, NULL, NULL);
  }
  if (permOut)
    for (unsigned i = 0; i < M; ++i)
      permOut[i] = a[i].idx;
  free(a), free(keys), free(in), free(out);
  return 0;
}

/* the physics part of Fluids::update (Fluids.cpp:409-457); perm_out (M words, optional) receives the sort's permutation */
int rocl_step(rocl* r, uint32_t* perm_out)
{
  if (!r)
    return -1;
  const size_t N = r->N;
#define RUN(k, n)       \
  if (run(r, (k), (n))) \
  return -1
  RUN(K_PREDICT, N);
  RUN(K_FILL_CELL_ID, N);
  sort_by_cell(r, perm_out);
  RUN(K_RESET_TABLE, r->cells);
  RUN(K_FILL_START, N);
  RUN(K_FILL_END, N);
  RUN(K_ADJUST_END, r->cells); /* m_simplifiedMode is true (Fluids.cpp:78) */
  for (int it = 0; it < r->jacobi; ++it)
  {
    RUN(K_BOUNDARY, N);
    RUN(K_DENSITY, N);
    RUN(K_LAMBDA, N);
    RUN(K_CORRECTION, N);
    RUN(K_CORRECT_POS, N);
  }
  RUN(K_UPDATE_VEL, N);
  if (r->fp.isVorticityConfEnabled)
  {
    RUN(K_VORTICITY, N);
    RUN(K_CONFINEMENT, N);
    cl.EnqueueCopyBuffer(r->queue, r->buf[B_VEL], r->buf[B_VELVISC], 0, 0, 16u * r->M, 0, NULL, NULL); /* Fluids.cpp:451 */
    RUN(K_XSPH, N);
  }
  RUN(K_UPDATE_POS, N);
#undef RUN
  cl.Finish(r->queue);
  return 0;
}

/* accumulated device time (OpenCL profiling events) and launch count per kernel since creation / the last call with reset */
int rocl_kernel_times(rocl* r, const char** names, double* us, unsigned* launches, int cap, int reset)
{
  if (!r)
    return 0;
  int n = 0;
  for (int k = 0; k < K_COUNT && n < cap; ++k, ++n)
  {
    names[n] = kKernels[k].name;
    us[n] = r->us[k];
    launches[n] = r->launches[k];
    if (reset)
      r->us[k] = 0.0, r->launches[k] = 0;
  }
  return n;
}
