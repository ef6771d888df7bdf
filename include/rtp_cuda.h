/*
 * rtp_cuda.h -- C ABI of the B200-native (sm_100a) CUDA backend for the RealTimeParticles hot path.
 *
 * This library replaces, for the hot path only, the reference's OpenCL layer:
 *   - CL::Context (string-keyed programs/kernels/buffers)         physics/ocl/Context.hpp:21-96
 *   - RadixSort (vendored OCLRadixSort)                            physics/utils/RadixSort.hpp:37-65
 *   - the kernel sequences of Boids/Fluids/Clouds::update()        physics/ocl/Boids.cpp:323-384,
 *                                                                  physics/ocl/Fluids.cpp:400-471,
 *                                                                  physics/ocl/Clouds.cpp:503-627
 * It is bound from C++ by Physics::CUDA::{Boids,Fluids,Clouds} (realtimeparticles_b200/cpp/CudaModels.hpp),
 * which derive from the unmodified Physics::Model (physics/Model.hpp:81-216), and from Python by ctypes
 * (realtimeparticles_b200/_abi.py) for the headless bench harness and the parity tests.
 *
 * Conventions: plain pointers and sizes, no exceptions across the boundary. Every call returns RTP_OK (0)
 * or a negative rtp_status; rtp_last_error() gives the text. One handle owns one CUDA device + one stream
 * and, like a reference model, is not thread-safe. There is NO CPU fallback: without a usable CUDA device
 * rtp_create() fails with RTP_ERR_CUDA.
 */
#ifndef RTP_CUDA_H
#define RTP_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define RTP_API __declspec(dllexport)
#else
#define RTP_API __attribute__((visibility("default")))
#endif

#define RTP_ABI_VERSION 1

typedef enum rtp_status
{
  RTP_OK = 0,
  RTP_ERR_INVALID = -1, /* bad argument / size mismatch / unknown field */
  RTP_ERR_CUDA = -2, /* CUDA runtime error, no device */
  RTP_ERR_STATE = -3, /* call not valid for this model type or state */
  RTP_ERR_COMM = -4 /* multi-GPU exchange error */
} rtp_status;

/* Physics::ModelType, physics/Model.hpp:16-21 */
typedef enum rtp_model
{
  RTP_MODEL_BOIDS = 0,
  RTP_MODEL_FLUIDS = 1,
  RTP_MODEL_CLOUDS = 2
} rtp_model;

/* Physics::Boundary, physics/Model.hpp:38-44 (boids only, Boids.cpp:363-373) */
typedef enum rtp_boundary
{
  RTP_BOUNDARY_BOUNCING_WALL = 0,
  RTP_BOUNDARY_CYCLIC_WALL = 1
} rtp_boundary;

/* Named device buffers. Names are the reference's string keys (SURVEY Appendix A:
 * Boids.cpp:126-136, Fluids.cpp:132-147, Clouds.cpp:166-199). f4 = float4 (xyz, w=0), f = float, u = uint32. */
typedef enum rtp_field
{
  RTP_F_POS = 0, /* p_pos            f4[M] */
  RTP_F_COL = 1, /* p_col            f4[M] */
  RTP_F_VEL = 2, /* p_vel            f4[M] */
  RTP_F_ACC = 3, /* p_acc            f4[M]  boids */
  RTP_F_PRED_POS = 4, /* p_predPos        f4[M]  fluids, clouds */
  RTP_F_CORR_POS = 5, /* p_corrPos        f4[M]  fluids, clouds (written only with RTP_STEP_DEBUG_FIELDS) */
  RTP_F_VORT = 6, /* p_vort           f4[M] */
  RTP_F_TOT_CORR_POS = 7, /* p_totCorrPos     f4[M]  clouds */
  RTP_F_DENSITY = 8, /* p_density        f[M] */
  RTP_F_CONST_FACTOR = 9, /* p_constFactor / p_constFactorFld  f[M] */
  RTP_F_TEMP = 10, /* p_temp           f[M]   clouds */
  RTP_F_VAPOR_DENS = 11, /* p_vaporDens      f[M] */
  RTP_F_CLOUD_DENS = 12, /* p_cloudDens      f[M] */
  RTP_F_BUOYANCY = 13, /* p_buoyancy       f[M] */
  RTP_F_CLOUD_GEN = 14, /* p_cloudGen       f[M] */
  RTP_F_PART_ID = 15, /* p_partID         f[M] */
  RTP_F_LAPLACIAN_TEMP = 16, /* p_laplacianTemp  f[M] */
  RTP_F_CONST_FACTOR_TEMP = 17, /* p_constFactorTemp f[M] */
  RTP_F_CORR_TEMP = 18, /* p_corrTemp       f[M] */
  RTP_F_CELL_ID = 19, /* p_cellID         u[M]   (sorted keys after a step) */
  RTP_F_CAMERA_DIST = 20, /* p_cameraDist     u[M] */
  RTP_F_START_END_CELL = 21, /* c_startEndPartID uint2[C] */
  RTP_F_PERM = 22, /* RadixSortIndices u[M]   permutation of the last cell sort: sorted[i] = unsorted[perm[i]] */
  RTP_F_CAMERA_PERM = 23, /* permutation of the last camera sort, u[M] */
  RTP_F_PART_DETECTOR = 24, /* c_partDetector   float8[C] */
  RTP_F_COUNT_
} rtp_field;

/* Kernel parameter blocks: field order and sizes are those of the reference's POD structs, which travel by
 * value into its kernels. */

/* BoidsRuleKernelInputs, physics/ocl/Boids.hpp:14-20 (16 B) */
typedef struct rtp_boids_params
{
  float velocityScale; /* 0.5  */
  float alignmentScale; /* 1.6  */
  float separationScale; /* 1.6  */
  float cohesionScale; /* 1.45 */
} rtp_boids_params;

/* TargetKernelInputs, physics/ocl/Boids.hpp:22-26 (8 B) */
typedef struct rtp_target_params
{
  float targetRadiusEffect; /* 2.0 */
  int32_t targetSignEffect; /* +1 attract, -1 repulse */
} rtp_target_params;

/* FluidKernelInputs, physics/ocl/Fluids.hpp:17-32 == FluidParams, kernels/define.cl:13-25 (44 B) */
typedef struct rtp_fluid_params
{
  float restDensity; /* 450 */
  float relaxCFM; /* 600 */
  float timeStep; /* 0.010 */
  uint32_t dim; /* 3 */
  uint32_t isArtPressureEnabled; /* 1 */
  float artPressureRadius; /* 0.006 */
  float artPressureCoeff; /* 0.001 */
  uint32_t artPressureExp; /* 4 */
  uint32_t isVorticityConfEnabled; /* 1 */
  float vorticityConfCoeff; /* 0.0004 */
  float xsphViscosityCoeff; /* 0.0001 */
} rtp_fluid_params;

/* CloudKernelInputs, physics/ocl/Clouds.hpp:15-45 == CloudParams, kernels/clouds.cl:19-33 (52 B) */
typedef struct rtp_cloud_params
{
  uint32_t dim; /* 3 */
  float timeStep; /* 0.01 */
  float restDensity; /* 450 (copied from the Fluids block, Clouds.cpp:308) */
  float groundHeatCoeff; /* 10 */
  float buoyancyCoeff; /* 0.10 */
  float gravCoeff; /* 0.0005 */
  float adiabaticLapseRate; /* 5 */
  float phaseTransitionRate; /* 0.3485 */
  float latentHeatCoeff; /* 0.07 */
  uint32_t isTempSmoothingEnabled; /* 1 */
  float relaxCFM; /* 600 */
  float initVaporDensityCoeff; /* 0.75 */
  float windCoeff; /* 1.0 */
} rtp_cloud_params;

/* Construction parameters == the part of Physics::ModelParams (physics/Model.hpp:60-73) the backend needs,
 * plus the per-model cap m_maxNbPartsInCell (Boids.cpp:78 -> 3000, Fluids.cpp:79 / Clouds.cpp:107 -> 100). */
typedef struct rtp_config
{
  int32_t model; /* rtp_model */
  int32_t device; /* CUDA device ordinal */
  uint64_t max_particles; /* m_maxNbParticles (M) */
  uint64_t nb_particles; /* m_currNbParticles (N <= M) */
  uint32_t box[3]; /* m_boxSize (integers in the reference, Geometry.hpp:42-47) */
  uint32_t grid[3]; /* m_gridRes */
  uint32_t dim; /* 2 or 3 */
  uint32_t max_parts_in_cell; /* 0 = model default */
} rtp_config;

typedef struct rtp_handle rtp_handle;

/* rtp_step() flags. A reference update() == PHYSICS | RENDER_AUX | CAMERA_SORT; on pause it is
 * CAMERA_SORT (+ clouds colour) only (Fluids.cpp:409, :466-468). */
#define RTP_STEP_PHYSICS 0x1u /* the solver stages */
#define RTP_STEP_RENDER_AUX 0x2u /* grid detector + colour kernels (grid.cl:43-60, fluids.cl:458-479, utils.cl:65-78) */
#define RTP_STEP_CAMERA_SORT 0x4u /* fillCameraDist + second sort (utils.cl:43-52, Fluids.cpp:466-468) */
#define RTP_STEP_DEBUG_FIELDS 0x8u /* also materialise intermediates a fused kernel would not write (p_corrPos) */

/* ---- life cycle (replaces CL::Context::Get()/release(), createProgram/createBuffers/createKernels) ---- */
RTP_API int rtp_abi_version(void);
RTP_API int rtp_device_count(void);
RTP_API int rtp_create(const rtp_config* cfg, rtp_handle** out);
RTP_API void rtp_destroy(rtp_handle* h);
RTP_API const char* rtp_last_error(const rtp_handle* h); /* h may be NULL: error of the last failed rtp_create */

/* ---- buffers (replaces Context::loadBufferFromHost / unloadBufferFromDevice, Context.cpp:355-400) ---- */
RTP_API int rtp_field_bytes(const rtp_handle* h, int field, size_t* bytes);
RTP_API int rtp_upload(rtp_handle* h, int field, const void* host, size_t bytes);
RTP_API int rtp_download(rtp_handle* h, int field, void* host, size_t bytes);
/* Stream-ordered variants for PAGE-LOCKED host buffers (a frame loop that streams its state through the host): the copy is
 * enqueued on the handle's stream, in order with the steps around it, and the call returns; the device data is in place
 * for the next rtp_step by stream order, the host buffer is complete / reusable after rtp_sync (or any blocking call).
 * With pageable memory they behave like the blocking calls for the host buffer's lifetime (the runtime stages it). */
RTP_API int rtp_upload_async(rtp_handle* h, int field, const void* host, size_t bytes);
RTP_API int rtp_download_async(rtp_handle* h, int field, void* host, size_t bytes);
/* device pointer of a field for zero-copy consumers (torch, CUDA-GL interop); valid until the next rtp_step */
RTP_API int rtp_device_ptr(rtp_handle* h, int field, void** dptr);

/* ---- parameters (replaces setKernelArg(.., sizeof(struct), &struct), Fluids.cpp:253-271 etc.) ---- */
RTP_API int rtp_set_boids_params(rtp_handle* h, const rtp_boids_params* rules, const rtp_target_params* target,
    const float target_pos[4], int target_active);
RTP_API int rtp_set_fluid_params(rtp_handle* h, const rtp_fluid_params* fluid, int nb_jacobi_iters);
RTP_API int rtp_set_cloud_params(rtp_handle* h, const rtp_cloud_params* cloud);
RTP_API int rtp_set_boundary(rtp_handle* h, int boundary);
RTP_API int rtp_set_nb_particles(rtp_handle* h, uint64_t n);
RTP_API int rtp_set_dimension(rtp_handle* h, int dim);
/* clouds colouring: which float field and which [min,max] (Clouds.cpp:610-617) */
RTP_API int rtp_set_displayed_quantity(rtp_handle* h, int field, float min_val, float max_val);

/* ---- reset-time kernels ---- */
/* resetCellIDs + resetCameraDist over M (grid.cl:65-71, utils.cl:35; Fluids.cpp:214-215) */
RTP_API int rtp_reset_ids(rtp_handle* h);
/* cld_initTemperature + cld_initVaporDensity over M (clouds.cl:116-135; Clouds.cpp:495-497) */
RTP_API int rtp_init_clouds_fields(rtp_handle* h);

/* ---- the step (replaces the body of {Boids,Fluids,Clouds}::update()) ---- */
RTP_API int rtp_step(rtp_handle* h, unsigned flags, const float camera_pos[3]);
/* n back-to-back steps replayed from one CUDA graph (bench / headless runs); same result as n rtp_step calls */
RTP_API int rtp_step_n(rtp_handle* h, unsigned flags, const float camera_pos[3], int n);
RTP_API int rtp_sync(rtp_handle* h);
/* the handle's cudaStream_t, so callers can time with events on the launching stream or order their own work */
RTP_API int rtp_get_stream(rtp_handle* h, void** stream);

/* ---- stand-alone access to the neighbour-search primitives (parity tests, other callers) ---- */
/* stable ascending sort of n 32-bit keys (device pointers); perm_out[i] = index of the i-th smallest key.
 * Replaces RadixSort::sort(key, ..) (RadixSort.cpp:122-161). key_bits = number of significant low bits. */
RTP_API int rtp_sort_keys(rtp_handle* h, const uint32_t* d_keys_in, uint32_t* d_keys_out, uint32_t* d_perm_out,
    uint64_t n, int key_bits);
/* host-buffer convenience wrapper around rtp_sort_keys (H2D, sort, D2H) */
RTP_API int rtp_sort_keys_host(rtp_handle* h, const uint32_t* keys_in, uint32_t* keys_out, uint32_t* perm_out,
    uint64_t n, int key_bits);

/* device self-test: the kernels' range-check-free exact sqrt / reciprocal vs the IEEE intrinsics over every float in
 * [lo, hi] (lo > 0); returns the number of bit mismatches of each */
RTP_API int rtp_selftest_math(rtp_handle* h, float lo, float hi, uint64_t* sqrt_mismatches, uint64_t* rcp_mismatches);

/* ---- profiling (replaces Context::enableProfiler + per-kernel event logging, Context.cpp:692-705) ---- */
RTP_API int rtp_enable_profiling(rtp_handle* h, int enable);
/* fills up to cap entries with the stage names / milliseconds of the last profiled step; returns the count */
RTP_API int rtp_get_stage_times(rtp_handle* h, const char** names, float* ms, int cap);
/* number of kernel launches (memset/memcpy nodes excluded) issued by the last rtp_step / per step of rtp_step_n */
RTP_API int rtp_last_launch_count(const rtp_handle* h);
/* Diagnostics of the neighbour lists after a physics-only rtp_step (no reference counterpart; the lists are an internal
 * optimisation, sweep.cuh): out = { particles, margin lists overflowed, centre cell changed since the list build (those
 * particles take the warp-cooperative 27-cell path), 0, sum of list lengths, longest list, hit lists overflowed, particles moved beyond the validity bound }. */
RTP_API int rtp_list_stats(rtp_handle* h, unsigned long long out[8]);

/* ---- initial-condition generators (host side; semantics of utils/Geometry.cpp:198-272) ---- */
/* out: float4[res.x*res.y*res.z]; returns number of points or negative rtp_status */
RTP_API int64_t rtp_gen_box_grid(float* out_xyzw, const int res[3], const float start[3], const float end[3]);
RTP_API int64_t rtp_gen_sphere_grid(float* out_xyzw, const int res[3], const float start[3], const float end[3]);
/* planar lattices of the 2D presets (utils/Geometry.cpp:8-196): plane 0 = XY, 1 = XZ, 2 = YZ; out: float4[res[0]*res[1]];
 * rectangle: res points along the plane's two axes; circle: res[0] angles x res[1] radii around the centre of start..end */
RTP_API int64_t rtp_gen_rectangle_grid(float* out_xyzw, int plane, const int res[2], const float start[3], const float end[3]);
RTP_API int64_t rtp_gen_circle_grid(float* out_xyzw, int plane, const int res[2], const float start[3], const float end[3]);
/* glibc rand()-driven uniform fill in (x,y,z) call order; seed<0 keeps the process' current rand() state */
RTP_API int64_t rtp_gen_random_box(float* out_xyzw, int64_t n, const float start[3], const float end[3], int seed);
/* the float a reference kernel sees for a -D constant: parse(FloatToStr(v)) (utils/Utils.cpp:24-29) */
RTP_API float rtp_baked_constant(float v);

/* ---- OpenGL interop (replaces the cl::BufferGL wrapping of the shared VBOs, Context.cpp:517-541, and
 * acquireGLBuffers / releaseGLBuffers around every frame, Context.cpp:710-750). The render engine owns the VBOs
 * (render/Engine.cpp:70-87 p_pos / p_col float4[maxNbParticles], :300-304 c_partDetector 8 floats per cell) and hands their
 * names to the model (ModelParams, Model.hpp:67-70). After rtp_register_gl the VBO IS the field: rtp_step maps it
 * (cudaGraphicsMapResources on the handle's stream), the kernels write straight into it, and it is unmapped before rtp_step
 * returns, so OpenGL may draw as soon as the stream has drained (rtp_sync) -- no copy. rtp_upload / rtp_download of a
 * registered field map it for the copy. Must be called on the thread that owns the OpenGL context; fails with RTP_ERR_CUDA
 * (and leaves the handle as it was) when there is none. field: RTP_F_POS, RTP_F_COL or RTP_F_PART_DETECTOR. */
RTP_API int rtp_register_gl(rtp_handle* h, int field, unsigned int vbo);
RTP_API int rtp_unregister_gl(rtp_handle* h, int field);

/* ---- boids target trajectory (host side; replaces Physics::Target, physics/utils/Target.cpp:11-50 with its three
 * PerlinNoise channels, physics/utils/PerlinNoise.cpp:10-85). The reference moves the target once per frame on the CPU
 * (Boids.cpp:351-358) and hands the position to bd_addTargetRule; a host does the same with rtp_target_update() and
 * rtp_set_boids_params(). dim: 2 or 3. */
typedef struct rtp_target rtp_target;
RTP_API rtp_target* rtp_target_create(uint32_t box_size);
RTP_API void rtp_target_destroy(rtp_target* t);
RTP_API int rtp_target_update(rtp_target* t, int dim, float particles_velocity, float out_pos[3]);

/* ---- multi-GPU slab decomposition (new design; fluids model) ----
 * One handle per GPU over the GLOBAL box/grid (cell ids stay global); it holds the rank's owned particles followed by
 * ghost copies of its slab neighbours' boundary layers. The kernels are the single-GPU ones; the caller
 * (realtimeparticles_b200/sharded.py) runs the step stage by stage and refreshes the ghost entries of the fields a
 * stage produced before the next stage reads them. */
typedef enum rtp_shard_stage_id
{
  RTP_SHARD_PREDICT = 0, /* predict + cell ids of the first n_owned particles, table reset */
  RTP_SHARD_GHOST_KEYS = 1, /* cell ids of the ghosts appended at [n_owned, nb_particles) from their p_predPos */
  RTP_SHARD_SORT = 2, /* sort by cell, payload gather (+ first boundary clamp), cell table */
  RTP_SHARD_DENSITY_LAMBDA = 3, /* iteration `iter` */
  RTP_SHARD_CORRECTION = 4, /* iteration `iter`; `last` != 0 also integrates the velocity */
  RTP_SHARD_VORTICITY = 5,
  RTP_SHARD_CONFINEMENT = 6,
  RTP_SHARD_XSPH = 7, /* + updatePosition: state back in p_pos / p_vel, sorted order */
  RTP_SHARD_DROP_GHOSTS = 8, /* compact p_pos / p_vel to the owned particles (first n_owned), cell-sorted order kept */
  RTP_SHARD_PREDICT_FROM = 9 /* predict + cell ids of the owned rows [iter, n_owned) only (arrivals of a migration), no reset */
} rtp_shard_stage_id;

/* internal buffers a slab exchange touches (device pointers valid until the next RTP_SHARD_SORT) */
typedef enum rtp_shard_buffer_id
{
  RTP_SHARD_BUF_KEYS_IN = 0, /* u32[M]   unsorted cell ids written by PREDICT / GHOST_KEYS */
  RTP_SHARD_BUF_PRED_IN = 1, /* f4[M]    unsorted predicted positions */
  RTP_SHARD_BUF_PRED_CUR = 2, /* f4[M]   sorted predicted positions the next stage reads */
  RTP_SHARD_BUF_LAMBDA = 3, /* f[M] */
  RTP_SHARD_BUF_VEL_SORTED = 4, /* f4[M] velocity after the last correction (input of the vorticity sweep) */
  RTP_SHARD_BUF_VORT_NORM = 5, /* f[M] */
  RTP_SHARD_BUF_VEL_CONFINED = 6, /* f4[M] velocity after vorticity confinement (input of the XSPH sweep) */
  RTP_SHARD_BUF_LIST_BUILD_POS = 7, /* f4[M] positions the neighbour lists were built from */
  RTP_SHARD_BUF_LIST_INVALID = 8, /* u32[2][16] per-epoch "lists invalid" flags: raised by own particles / by ghosts */
  RTP_SHARD_BUF_POS = 9, /* f4[M] p_pos (unsorted between steps: migration) */
  RTP_SHARD_BUF_VEL = 10, /* f4[M] p_vel */
  RTP_SHARD_BUF_ROW_BOUNDS = 11 /* u32[4] rtp_shard_set_interior: first interior row, one past the last, rows holding a
                                   particle, error flag (a launch by row phase was sized too small: results are invalid) */
} rtp_shard_buffer_id;

/* number of leading (unsorted) particles this rank owns; the rest of nb_particles are ghosts */
RTP_API int rtp_shard_set_owned(rtp_handle* h, uint64_t n_owned);
RTP_API int rtp_shard_stage(rtp_handle* h, int stage, int iter, int last);
RTP_API int rtp_shard_buffer(rtp_handle* h, int which, void** dptr, size_t* bytes);
/* The exchange kernels of the slab data plane (the transport between ranks -- NCCL / gloo send-recv -- stays with the
 * caller). All pointers are DEVICE pointers, everything is enqueued on the handle's stream, nothing synchronises.
 *  pack:   d_out[k] = buffer[d_idx[k]], k < n; an index 0xFFFFFFFF ("no particle": fixed-capacity exchanges are padded)
 *          packs a +inf position (f4 rows) or 0 (scalar rows);  unpack: buffer[d_idx[k]] = d_in[k], padding skipped.
 *  clear_rows: p_pos[d_idx[k]] = +inf, p_vel[d_idx[k]] = 0 (and the prediction / cell id RTP_SHARD_PREDICT gives such a row):
 *          the row holds no particle any more (it migrated); such rows sort
 *          behind every particle, appear in no cell range and are skipped by the sweeps of a sharded handle.
 *  inverse_perm: d_inv[perm[i]] = i for the nb_particles cell-sorted rows (where did unsorted row j go).
 *  check_ghosts: raise the "lists invalid" flag of next_epoch when a ghost row (sorted indices d_sorted_idx) has been moved
 *          further than the list validity bound from its position at the list build (its owner moves it, not this rank). */
RTP_API int rtp_shard_pack(rtp_handle* h, int buffer, const uint32_t* d_idx, uint64_t n, void* d_out);
RTP_API int rtp_shard_unpack(rtp_handle* h, int buffer, const uint32_t* d_idx, uint64_t n, const void* d_in);
RTP_API int rtp_shard_clear_rows(rtp_handle* h, const uint32_t* d_idx, uint64_t n);
/* classify: for the first n (unsorted) rows, after RTP_SHARD_PREDICT: d_below[i] = the row holds a particle (finite p_pos)
 *          whose predicted cell x-layer is < layer_below, d_above[i] = ... >= layer_from (bytes 0 / 1): the leavers of a
 *          migration, the face layers of a halo; the caller compacts the masks (order-preserving). */
RTP_API int rtp_shard_classify(rtp_handle* h, uint64_t n, uint32_t layer_below, uint32_t layer_from, uint8_t* d_below, uint8_t* d_above);
RTP_API int rtp_shard_inverse_perm(rtp_handle* h, uint32_t* d_inv);
RTP_API int rtp_shard_check_ghosts(rtp_handle* h, const uint32_t* d_sorted_idx, uint64_t n, int next_epoch);
/* neighbour-list validity: radius^2 a particle may move from its list-build position (see sweep.cuh) */
RTP_API float rtp_shard_list_dmax_sq(const rtp_handle* h);

/* ---- overlap of a ghost refresh with the sweeps of the interior rows ----
 * The cells [cell_lo, cell_hi) (x-layers two or more from both faces of the slab) hold the INTERIOR particles: no ghost among
 * their neighbours, and nobody's ghost. Every RTP_SHARD_SORT finds their (contiguous) range of sorted rows on the device; a
 * neighbour sweep can then run in two launches -- rtp_shard_stage_rows(..., RTP_ROWS_INTERIOR) first, RTP_ROWS_BOUNDARY
 * second -- and only the second has to wait for the refresh of the previous stage:
 *     stage k BOUNDARY | fork, pack, <transport on the exchange stream>, unpack, check_ghosts, done | stage k+1 INTERIOR |
 *     join | stage k+1 BOUNDARY | ...
 * Between fork and done, pack / unpack / check_ghosts are enqueued on the exchange stream (a second, high-priority stream
 * of the handle; the caller runs its transport on it too); join makes the compute stream wait for the last `done`.
 * A step runs ALL its sweeps by row phase or none (the straggler queues of the sweeps are split by row class).
 * max_boundary_rows bounds the rows outside the interior (ghost rows + the owned rows of the face layers: the caller's
 * exchange capacities): it sizes the BOUNDARY launches; "no particle" rows are not visited by launches by row phase. */
/* ghost layers per slab face the interior range of rtp_shard_set_interior is understood with (sharded.py: GHOST_LAYERS) */
#define RTP_SHARD_GHOST_LAYERS 2

typedef enum rtp_rows_id
{
  RTP_ROWS_ALL = 0,
  RTP_ROWS_BOUNDARY = 1, /* thread blocks holding a row outside the interior range (ghost rows included) */
  RTP_ROWS_INTERIOR = 2 /* thread blocks entirely inside it */
} rtp_rows_id;
RTP_API int rtp_shard_set_interior(rtp_handle* h, uint32_t cell_lo, uint32_t cell_hi, uint64_t max_boundary_rows);
RTP_API int rtp_shard_stage_rows(rtp_handle* h, int stage, int iter, int last, int rows);
RTP_API int rtp_shard_exchange_stream(rtp_handle* h, void** stream);
RTP_API int rtp_shard_exchange_fork(rtp_handle* h);
RTP_API int rtp_shard_exchange_done(rtp_handle* h);
RTP_API int rtp_shard_exchange_join(rtp_handle* h);

/* ---- the slab decomposition driven from inside the library (csrc/slab_group.cu) ----
 * One host thread, `nslabs` slabs of the GLOBAL box along x, slab r on CUDA device dev_ids[r] (ids may repeat: several
 * slabs on one GPU -- how the 1-GPU tests run it). Each slab is an rtp handle of `slab_capacity` rows in the static row
 * layout described above rtp_shard_stage_id / in realtimeparticles_b200/sharded.py: own region, arrival slots, one ghost
 * region of ghost_cap rows per face (0: slab_capacity / 8; migrate_cap 0: ghost_cap / 4). A step is the same sequence
 * sharded.py runs -- predict, migrate, halo, joint sort, 2I+3 ghost refreshes, compaction -- and the transport is a PULL over
 * peer memory: the receiving slab copies its neighbour's packed rows with cudaMemcpyAsync (NVLink between GPUs), ordered
 * by events between the slabs' streams. overlap != 0: the refreshes travel on the slabs' exchange streams while the next
 * sweep runs its interior rows (rtp_shard_stage_rows). rtp_slab_group_step only ENQUEUES; nothing synchronises with the
 * host until rtp_slab_group_sync / _check / _download. Bit-identical to the Python orchestration (tests/test_sharded.py). */
typedef struct rtp_slab_group rtp_slab_group;
RTP_API int rtp_slab_group_create(rtp_slab_group** out, int nslabs, const int* dev_ids, uint64_t slab_capacity, const uint32_t box[3],
    const uint32_t grid[3], uint64_t ghost_cap, uint64_t migrate_cap, int overlap);
RTP_API void rtp_slab_group_destroy(rtp_slab_group* g);
RTP_API const char* rtp_slab_group_last_error(const rtp_slab_group* g); /* g may be NULL: error of the last failed create */
RTP_API int rtp_slab_group_size(const rtp_slab_group* g);
/* the rtp handle of one slab (fields in cell-sorted order incl. ghosts: diagnostics, rtp_download of p_density ...) */
RTP_API int rtp_slab_group_handle(rtp_slab_group* g, int slab, rtp_handle** h);
RTP_API int rtp_slab_group_set_fluid_params(rtp_slab_group* g, const rtp_fluid_params* fluid, int nb_jacobi_iters);
/* the full initial state (host, float4 rows); every slab keeps the particles whose cell x-layer it owns */
RTP_API int rtp_slab_group_upload(rtp_slab_group* g, const float* pos_xyzw, const float* vel_xyzw, uint64_t n);
RTP_API int rtp_slab_group_step(rtp_slab_group* g, int nsteps);
RTP_API int rtp_slab_group_sync(rtp_slab_group* g);
/* synchronises and reads the device-side capacity flags: RTP_ERR_COMM when a ghost region, a migration message or the
 * arrival slots overflowed since the last check (results are then invalid); migrated_total: particles that changed slab */
RTP_API int rtp_slab_group_check(rtp_slab_group* g, uint64_t* migrated_total);
/* all particles, slab by slab (inside a slab: cell-sorted); returns their number or a negative rtp_status;
 * per_slab (optional): nslabs counts */
RTP_API int64_t rtp_slab_group_download(rtp_slab_group* g, float* pos_xyzw, float* vel_xyzw, uint64_t capacity_rows, uint64_t* per_slab);

#ifdef __cplusplus
}
#endif
#endif /* RTP_CUDA_H */
