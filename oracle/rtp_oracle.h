/*
 * rtp_oracle.h -- CPU restatement of the RealTimeParticles hot path. TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (realtimeparticles_b200/, include/rtp_cuda.h) never links, imports or calls it.
 *
 * PARITY PIN: the reference ships no tests, golden vectors or fixtures for this path and its
 * arithmetic lives in OpenCL C that cannot be JIT-compiled here (no OpenCL ICD / headers). The restatement is
 * therefore pinned two ways: (1) against oracle/_ref/libref_kernels.so, which compiles the reference's own
 * unmodified .cl kernel sources (from /root/reference, through the OpenCL-C shim oracle/ref/ocl_shim.hpp) and
 * executes them on the CPU -- see oracle/ref/README.md and tests/test_oracle_vs_ref.py; (2) against fixtures of
 * the reference's own compiled utils/Geometry.cpp + utils/Utils.cpp (initial states, baked -D constants).
 * What stays unpinned: the OpenCL built-ins (pow, exp, fast_length, dot ...) are implemented by whichever OpenCL
 * driver runs the reference; the canonical IEEE-fp32 choices made here are listed in DESIGN.md section "Oracle".
 */
#ifndef RTP_ORACLE_H
#define RTP_ORACLE_H

#include "../include/rtp_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_world orc_world;

/* individual stages == individual reference kernel launches (names = reference kernel names) */
typedef enum orc_stage
{
  /* grid.cl / RadixSort */
  ORC_FILL_CELL_IDS = 0, /* fillCellIDs on p_pos (boids) or p_predPos (fluids, clouds) */
  ORC_SORT_BY_CELL, /* RadixSort::sort("p_cellID", model payload list) */
  ORC_BUILD_CELL_TABLE, /* resetStartEndCell, fillStartCell, fillEndCell, adjustEndCell */
  /* boids.cl */
  ORC_BD_RULES, /* bd_applyBoidsRulesWithGrid3D / 2D by dimension */
  ORC_BD_TARGET, /* bd_addTargetRule (only if target active) */
  ORC_BD_UPDATE_VEL,
  ORC_BD_UPDATE_POS, /* wall or periodic by boundary */
  /* fluids.cl (clouds.cl variants when model == clouds) */
  ORC_PREDICT_POS,
  ORC_APPLY_BOUNDARY, /* fld_applyBoundaryCondition / cld_applyMixedBoundaryConditions on p_predPos */
  ORC_DENSITY,
  ORC_CONSTRAINT_FACTOR,
  ORC_CONSTRAINT_CORRECTION,
  ORC_CORRECT_POS, /* fluids: predPos += corrPos ; clouds: predPos += corrPos and totCorrPos += corrPos */
  ORC_UPDATE_VEL,
  ORC_VORTICITY,
  ORC_VORTICITY_CONFINEMENT,
  ORC_XSPH, /* copy p_vel -> p_velInViscosity, then xsph kernel */
  ORC_UPDATE_POS,
  /* clouds.cl thermodynamics */
  ORC_CLD_THERMO, /* copy+heatFromGround, buoyancy, adiabaticCooling, generateCloud, copies+phaseTransition, latentHeat */
  ORC_CLD_LAPLACIAN_TEMP,
  ORC_CLD_CONSTRAINT_FACTOR_TEMP,
  ORC_CLD_CONSTRAINT_CORRECTION_TEMP,
  ORC_CLD_CORRECT_TEMP,
  /* render side */
  ORC_RENDER_AUX, /* grid detector + colour */
  ORC_CAMERA_SORT, /* fillCameraDist + RadixSort::sort("p_cameraDist", ...) */
  ORC_STAGE_COUNT_
} orc_stage;

int orc_create(const rtp_config* cfg, orc_world** out);
void orc_destroy(orc_world* w);

/* direct pointer to a named buffer (rtp_field) so numpy can wrap it; bytes = its size */
void* orc_field_ptr(orc_world* w, int field, size_t* bytes);

int orc_set_boids_params(orc_world* w, const rtp_boids_params* rules, const rtp_target_params* target,
    const float target_pos[4], int target_active);
int orc_set_fluid_params(orc_world* w, const rtp_fluid_params* fluid, int nb_jacobi_iters);
int orc_set_cloud_params(orc_world* w, const rtp_cloud_params* cloud);
int orc_set_boundary(orc_world* w, int boundary);
int orc_set_nb_particles(orc_world* w, uint64_t n);
int orc_set_dimension(orc_world* w, int dim);
int orc_set_displayed_quantity(orc_world* w, int field, float min_val, float max_val);
void orc_set_camera(orc_world* w, const float cam[3]);

int orc_reset_ids(orc_world* w);
int orc_init_clouds_fields(orc_world* w);

int orc_run_stage(orc_world* w, int stage);
/* full {Boids,Fluids,Clouds}::update() with the same flags as rtp_step */
int orc_step(orc_world* w, unsigned flags, const float cam[3]);

/* baked constants as the reference kernels see them */
float orc_constant(const orc_world* w, const char* name);
float orc_baked_constant(float v);
int orc_max_threads(void);
void orc_set_threads(int n);

/* stand-alone stable sort (semantics of RadixSort::sort over n keys) */
void orc_sort_keys(const uint32_t* keys_in, uint32_t* keys_out, uint32_t* perm_out, uint64_t n);

/* initial-condition generators restating utils/Geometry.cpp:198-272 */
int64_t orc_gen_box_grid(float* out_xyzw, const int res[3], const float start[3], const float end[3]);
int64_t orc_gen_sphere_grid(float* out_xyzw, const int res[3], const float start[3], const float end[3]);
int64_t orc_gen_random_box(float* out_xyzw, int64_t n, const float start[3], const float end[3], int seed);

#ifdef __cplusplus
}
#endif
#endif
