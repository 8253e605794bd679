/*
 * grav_sim_shim.c -- reference-facing host side of the B200 acceleration path (plain C).
 *
 * Exports, with the reference's exact signatures and error behaviour, every public symbol of
 * the three reference files it stands in for:
 *     src/acceleration.c            get_new_acceleration_param, finalize_acceleration_param,
 *                                   acceleration, benchmark_acceleration
 *     src/acceleration_barnes_hut.c acceleration_barnes_hut
 *     src/linear_octree.c           get_new_linear_octree, construct_octree, free_linear_octree,
 *                                   linear_octree_check_if_included
 * and forwards the arithmetic to libgrav_b200.so through the C ABI in include/grav_b200.h.
 * There is no CPU implementation in here: without a usable GPU every call returns
 * GRAV_FAILURE with the device error in the traceback.
 *
 * Built two ways: (a) stand-alone as libgrav_sim_b200.so (tests bind it with ctypes); the weak
 * definitions at the bottom then supply ErrorStatus helpers; (b) compiled into the reference's
 * libgrav_sim in place of the three files above, where src/error.c's strong
 * definitions win.
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "grav_b200.h"
#include "grav_sim_abi.h"

#define SHIM_RAISE(code, msg) raise_error(__FILE__, __LINE__, __func__, (code), (msg))

static ErrorStatus status_from_rc(const int rc, const char *file, const int line, const char *func)
{
    if (rc == GRAV_B200_OK)
    {
        return make_success_error_status();
    }
    int code = GRAV_FAILURE;
    if (rc == GRAV_B200_EINVAL)
    {
        code = GRAV_VALUE_ERROR;
    }
    else if (rc == GRAV_B200_ENOMEM)
    {
        code = GRAV_MEMORY_ERROR;
    }
    return raise_error(file, line, func, code, grav_b200_last_error());
}
#define SHIM_STATUS(rc) status_from_rc((rc), __FILE__, __LINE__, __func__)

static ErrorStatus raise_fmt(const char *file, const int line, const char *func, const int code,
                             const char *fmt, const double dval, const int ival, const int use_int)
{
    char msg[160];
    if (use_int)
    {
        snprintf(msg, sizeof(msg), fmt, ival);
    }
    else
    {
        snprintf(msg, sizeof(msg), fmt, dval);
    }
    return raise_error(file, line, func, code, msg);
}

/* ---- parameters (reference: src/acceleration.c:62-129) ------------------------------------ */

AccelerationParam get_new_acceleration_param(void)
{
    AccelerationParam p;
    memset(&p, 0, sizeof(p));
    p.method = ACCELERATION_METHOD_PAIRWISE;
    p.opening_angle = 1.0;
    p.softening_length = 0.0;
    p.max_num_particles_per_leaf = -1;
    return p;
}

static int method_is_known(const int method)
{
    return method == ACCELERATION_METHOD_PAIRWISE || method == ACCELERATION_METHOD_MASSLESS ||
           method == ACCELERATION_METHOD_BARNES_HUT;
}

ErrorStatus finalize_acceleration_param(AccelerationParam *acceleration_param)
{
    AccelerationParam *p = acceleration_param;
    if (!method_is_known(p->method))
    {
        return raise_fmt(__FILE__, __LINE__, __func__, GRAV_VALUE_ERROR,
                         "Unknown acceleration method. Got: %d", 0.0, p->method, 1);
    }
    if (p->softening_length < 0.0)
    {
        return raise_fmt(__FILE__, __LINE__, __func__, GRAV_VALUE_ERROR,
                         "Softening length is negative. Got: %.3g", p->softening_length, 0, 0);
    }
    if (p->method == ACCELERATION_METHOD_BARNES_HUT)
    {
        if (p->opening_angle < 0.0)
        {
            return raise_fmt(__FILE__, __LINE__, __func__, GRAV_VALUE_ERROR,
                             "Opening angle is negative. Got: %.3g", p->opening_angle, 0, 0);
        }
        if (p->max_num_particles_per_leaf == -1)
        {
            p->max_num_particles_per_leaf = 1;
        }
        else if (p->max_num_particles_per_leaf < 1)
        {
            return raise_fmt(__FILE__, __LINE__, __func__, GRAV_VALUE_ERROR,
                             "Maximum number of particles per leaf must be positive. Got: %d", 0.0,
                             p->max_num_particles_per_leaf, 1);
        }
    }
    return make_success_error_status();
}

/* ---- dispatch (reference: src/acceleration.c:131-154) -------------------------------------- */

ErrorStatus acceleration(double *restrict a, const System *restrict system,
                         const AccelerationParam *restrict acceleration_param)
{
    const AccelerationParam *p = acceleration_param;
    switch (p->method)
    {
        case ACCELERATION_METHOD_PAIRWISE:
            return SHIM_STATUS(grav_b200_acceleration_pairwise(a, system->num_particles, system->x, system->m,
                                                               system->G, p->softening_length));
        case ACCELERATION_METHOD_MASSLESS:
            return SHIM_STATUS(grav_b200_acceleration_massless(a, system->num_particles, system->x, system->m,
                                                               system->G, p->softening_length));
        case ACCELERATION_METHOD_BARNES_HUT:
            return acceleration_barnes_hut(a, system, p);
        default:
            return raise_fmt(__FILE__, __LINE__, __func__, GRAV_VALUE_ERROR,
                             "Unknown acceleration method. Got: %d", 0.0, p->method, 1);
    }
}

/* reference: src/acceleration_barnes_hut.c:33-76 (tree is built and dropped inside the call) */
ErrorStatus acceleration_barnes_hut(double *restrict a, const System *restrict system,
                                    const AccelerationParam *restrict acceleration_param)
{
    return SHIM_STATUS(grav_b200_acceleration_barnes_hut(
        a, system->num_particles, system->x, system->m, system->G, acceleration_param->softening_length,
        acceleration_param->opening_angle, acceleration_param->max_num_particles_per_leaf));
}

/* reference: whfast_acceleration dispatch, src/integrator_whfast.c:817-837 */
ErrorStatus grav_b200_shim_whfast_acceleration(double *restrict a, const System *system,
                                               const double *restrict jacobi_x, const double *restrict eta,
                                               const AccelerationParam *acceleration_param)
{
    switch (acceleration_param->method)
    {
        case ACCELERATION_METHOD_PAIRWISE:
            return SHIM_STATUS(grav_b200_whfast_acceleration_pairwise(a, system->num_particles, system->x, system->m,
                                                                      system->G, jacobi_x, eta,
                                                                      acceleration_param->softening_length));
        case ACCELERATION_METHOD_MASSLESS:
            return SHIM_STATUS(grav_b200_whfast_acceleration_massless(a, system->num_particles, system->x, system->m,
                                                                      system->G, jacobi_x, eta,
                                                                      acceleration_param->softening_length));
        default:
            return SHIM_RAISE(GRAV_VALUE_ERROR,
                              "Invalid acceleration method for WHFast integrator. Only pairwise and massless "
                              "methods are supported.");
    }
}

/* ---- linear octree (reference: src/linear_octree.c:90-103, 825-985) ------------------------ */

LinearOctree get_new_linear_octree(void)
{
    LinearOctree t;
    memset(&t, 0, sizeof(t)); /* the reference leaves box_width / counts / one pointer unset; NULL is a superset */
    return t;
}

ErrorStatus construct_octree(LinearOctree *restrict octree, const System *restrict system,
                             const AccelerationParam *restrict acceleration_param,
                             const double *restrict box_center, const double box_width)
{
    if (!octree)
    {
        return SHIM_RAISE(GRAV_POINTER_ERROR, "Octree pointer is NULL");
    }
    if (!system)
    {
        return SHIM_RAISE(GRAV_POINTER_ERROR, "System pointer is NULL");
    }
    if (!acceleration_param)
    {
        return SHIM_RAISE(GRAV_POINTER_ERROR, "Acceleration parameter pointer is NULL");
    }
    *octree = get_new_linear_octree();
    const int rc = grav_b200_construct_octree(
        system->num_particles, system->x, system->m, acceleration_param->max_num_particles_per_leaf, box_center,
        box_width, &octree->box_width, &octree->num_internal_nodes, &octree->particle_morton_indices_deepest_level,
        &octree->sorted_indices, &octree->tree_num_particles, &octree->tree_num_internal_children,
        &octree->tree_first_particle_sorted_idx, &octree->tree_first_internal_children_idx, &octree->tree_mass,
        &octree->tree_center_of_mass_x, &octree->tree_center_of_mass_y, &octree->tree_center_of_mass_z);
    if (rc != GRAV_B200_OK)
    {
        free_linear_octree(octree);
        *octree = get_new_linear_octree();
    }
    return SHIM_STATUS(rc);
}

void free_linear_octree(LinearOctree *restrict octree)
{
    free(octree->particle_morton_indices_deepest_level);
    free(octree->sorted_indices);
    free(octree->tree_num_particles);
    free(octree->tree_num_internal_children);
    free(octree->tree_first_particle_sorted_idx);
    free(octree->tree_first_internal_children_idx);
    free(octree->tree_mass);
    free(octree->tree_center_of_mass_x);
    free(octree->tree_center_of_mass_y);
    free(octree->tree_center_of_mass_z);
}

bool linear_octree_check_if_included(const int64_t morton_index_i, const int64_t morton_index_j, const int level)
{
    const int shift = 3 * (MORTON_MAX_LEVEL - level);
    return (morton_index_i >> shift) == (morton_index_j >> shift);
}

/* ---- benchmark harness (reference: src/acceleration.c:369-507) ----------------------------- */

static double now_seconds(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static const char *method_name(const int method)
{
    switch (method)
    {
        case ACCELERATION_METHOD_PAIRWISE:
            return "Pairwise";
        case ACCELERATION_METHOD_MASSLESS:
            return "Massless";
        case ACCELERATION_METHOD_BARNES_HUT:
            return "Barnes-Hut";
        default:
            return NULL;
    }
}

ErrorStatus benchmark_acceleration(const System *restrict system, const AccelerationParam *acceleration_params,
                                   const int num_acceleration_params, const int *restrict num_times_acceleration_param)
{
    const size_t len = (size_t)system->num_particles * 3;
    double *first = malloc(len * sizeof(double));
    double *cur = malloc(len * sizeof(double));
    if (!first || !cur)
    {
        free(first);
        free(cur);
        return SHIM_RAISE(GRAV_MEMORY_ERROR, "Failed to allocate memory for acceleration arrays");
    }
    ErrorStatus status = make_success_error_status();

    fputs("Benchmarking acceleration...\n", stdout);
    for (int t = 0; t < num_acceleration_params; t++)
    {
        const AccelerationParam *param = &acceleration_params[t];
        const int reps = num_times_acceleration_param[t];
        if (reps <= 0)
        {
            printf("Test %d:    Skipped since num_times: %d <= 0\n\n", t, reps);
            continue;
        }
        if (!method_name(param->method))
        {
            status = raise_fmt(__FILE__, __LINE__, __func__, GRAV_VALUE_ERROR, "Unknown acceleration method. Got: %d",
                               0.0, param->method, 1);
            break;
        }

        /* the very first evaluation of the whole benchmark is the comparison vector */
        double sum = 0.0, sum_sq = 0.0, mae = 0.0;
        for (int r = 0; r < reps; r++)
        {
            double *out = (t == 0 && r == 0) ? first : cur;
            const double t0 = now_seconds();
            status = acceleration(out, system, param);
            const double dt = now_seconds() - t0;
            if (status.return_code != GRAV_SUCCESS)
            {
                goto done;
            }
            sum += dt;
            sum_sq += dt * dt;
            if (t != 0 && r == 0)
            {
                for (size_t k = 0; k < len; k++)
                {
                    mae += fabs(first[k] - cur[k]);
                }
                mae /= system->num_particles;
            }
        }
        const double mean = sum / reps;
        double var = 0.0;
        if (reps > 1)
        {
            var = (sum_sq - reps * mean * mean) / (reps - 1);
            if (var < 0.0)
            {
                var = 0.0;
            }
        }
        printf("Test %d:    Method: %s\n", t, method_name(param->method));
        printf("    Number of times: %d\n", reps);
        printf("    Avg time: %.3g (+- %.3g) s\n", mean, sqrt(var));
        printf("    MAE: %.3g\n\n", mae);
    }
done:
    free(first);
    free(cur);
    return status;
}

/* ---- stand-alone ErrorStatus helpers (weak: src/error.c overrides them inside libgrav_sim) -- */

__attribute__((weak)) ErrorStatus make_success_error_status(void)
{
    ErrorStatus s;
    s.return_code = GRAV_SUCCESS;
    s.traceback = NULL;
    s.traceback_code_ = GRAV_TRACEBACK_NOT_INITIALIZED;
    return s;
}

__attribute__((weak)) ErrorStatus raise_error(const char *error_file, const int error_line, const char *error_func,
                                              const int error_code, const char *error_msg)
{
    ErrorStatus s;
    s.return_code = error_code;
    s.traceback_code_ = 0;
    const size_t cap = strlen(error_file) + strlen(error_func) + strlen(error_msg) + 64;
    s.traceback = malloc(cap);
    if (s.traceback)
    {
        snprintf(s.traceback, cap, "%s:%d in %s: %s", error_file, error_line, error_func, error_msg);
    }
    else
    {
        s.traceback_code_ = 1;
    }
    return s;
}
