/*
 * grav_sim_abi.h -- the slice of grav_sim's C ABI that the acceleration path is called through.
 *
 * These are layout-compatible declarations of the reference's public types and the symbols
 * our shim (gravity-simulator_b200/csrc/grav_sim_shim.c) exports in place of the reference's
 * src/acceleration.c, src/acceleration_barnes_hut.c and src/linear_octree.c.  Struct field
 * order and types are interface facts taken from the reference headers cited per item; when
 * the shim is compiled inside the reference tree, define GRAV_SIM_USE_REFERENCE_HEADERS and
 * the reference's own headers are included instead.
 */
#ifndef GRAV_SIM_ABI_H
#define GRAV_SIM_ABI_H

#ifdef GRAV_SIM_USE_REFERENCE_HEADERS
#include "acceleration.h"
#include "error.h"
#include "linear_octree.h"
#include "system.h"
#else

#include <stdbool.h>
#include <stdint.h>

/* error codes, src/error.h:14-21 */
#define GRAV_SUCCESS 0
#define GRAV_FAILURE 1
#define GRAV_VALUE_ERROR 2
#define GRAV_POINTER_ERROR 3
#define GRAV_MEMORY_ERROR 4
#define GRAV_TRACEBACK_NOT_INITIALIZED -1

/* src/error.h:31-36; returned by value (24 bytes, sret on SysV x86-64) */
typedef struct ErrorStatus {
    int return_code;
    char *traceback;
    int traceback_code_;
} ErrorStatus;

/* src/system.h:12-20 */
typedef struct System {
    int num_particles;
    int *particle_ids;
    double *x; /* AoS [3N] */
    double *v; /* AoS [3N] */
    double *m; /* [N]      */
    double G;
} System;

/* src/acceleration.h:16-26 */
#define ACCELERATION_METHOD_PAIRWISE 1
#define ACCELERATION_METHOD_MASSLESS 2
#define ACCELERATION_METHOD_BARNES_HUT 3
typedef struct AccelerationParam {
    int method;
    double opening_angle;
    double softening_length;
    int max_num_particles_per_leaf;
} AccelerationParam;

/* src/linear_octree.h:16-57 */
#define MORTON_MAX_LEVEL 21
typedef struct LinearOctree {
    double box_width;
    int num_internal_nodes; /* actually the total node count, leaves included */
    int64_t *particle_morton_indices_deepest_level;
    int *sorted_indices;
    int *tree_num_particles;
    int *tree_num_internal_children;
    int *tree_first_particle_sorted_idx;
    int *tree_first_internal_children_idx;
    double *tree_mass;
    double *tree_center_of_mass_x;
    double *tree_center_of_mass_y;
    double *tree_center_of_mass_z;
} LinearOctree;

/* provided by the reference's src/error.c when linked into libgrav_sim; the stand-alone shim
 * library carries weak equivalents */
ErrorStatus make_success_error_status(void);
ErrorStatus raise_error(const char *error_file, const int error_line, const char *error_func,
                        const int error_code, const char *error_msg);

/* ---- symbols the shim exports (signatures: src/acceleration.h:33-103, src/linear_octree.h:64-101) */
AccelerationParam get_new_acceleration_param(void);
ErrorStatus finalize_acceleration_param(AccelerationParam *acceleration_param);
ErrorStatus acceleration(double *a, const System *system, const AccelerationParam *acceleration_param);
ErrorStatus acceleration_barnes_hut(double *a, const System *system, const AccelerationParam *acceleration_param);
ErrorStatus benchmark_acceleration(const System *system, const AccelerationParam *acceleration_params,
                                   const int num_acceleration_params, const int *num_times_acceleration_param);
LinearOctree get_new_linear_octree(void);
ErrorStatus construct_octree(LinearOctree *octree, const System *system, const AccelerationParam *acceleration_param,
                             const double *box_center, const double box_width);
void free_linear_octree(LinearOctree *octree);
bool linear_octree_check_if_included(const int64_t morton_index_i, const int64_t morton_index_j, const int level);

#endif /* GRAV_SIM_USE_REFERENCE_HEADERS */

/* WHFast side door: the bodies of the two static kernels in src/integrator_whfast.c:839-957 and
 * :959-1264 forward here (one-line patch shown in INTEGRATION.md). */
ErrorStatus grav_b200_shim_whfast_acceleration(double *a, const System *system, const double *jacobi_x,
                                               const double *eta, const AccelerationParam *acceleration_param);

#endif /* GRAV_SIM_ABI_H */
