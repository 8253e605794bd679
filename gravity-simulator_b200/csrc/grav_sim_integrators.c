/*
 * grav_sim_integrators.c -- device-resident time loops for the fixed-step integrators whose every sub-step is on
 * the hot path: leapfrog (config 2 and 4), WHFast (config 3) and Euler / Euler-Cromer / RK4.  Plain C,
 * reference-facing.
 *
 * Also here: the hook for the O(N^2) energy diagnostic (compute_energy, src/utils.c:27-59; compute_energy_python,
 * src/python_interface.c:195-243), which at config 2/3 sizes costs more than many integration steps on the host.
 *
 * Compiled ONLY inside the reference tree (it uses the reference's IntegratorParam / OutputParam /
 * SimulationStatus / Settings, output_snapshot() and the progress bar as they are -- INTEGRATION.md section 1),
 * together with a three-line hook at the top of the reference's own leapfrog(), euler(), euler_cromer(), rk4()
 * (src/integrator.c:894, :281, :456, :630) and whfast() (src/integrator_whfast.c:200):
 *
 *     { ErrorStatus es_; if (grav_b200_shim_whfast(&es_, system, integrator_param, acceleration_param,
 *                                                  output_param, simulation_status, settings, tf)) return es_; }
 *
 * Each hook returns 0 ("not handled") when GRAV_B200_RESIDENT=0 is set or the configuration is not one it
 * covers, and the reference's own loop then runs with acceleration() forwarding to the GPU per call.  When it
 * handles the run, the particle state is uploaded once, every step (drift, force, kick -- for WHFast also the
 * distance sort, eta, the Kepler solve and both coordinate transforms) runs on the device, and the host arrays
 * in `system` are refreshed only where the reference itself looks at them: before output_snapshot() and at the
 * end.  Bookkeeping (t, num_steps, dt overshoot, output schedule, is_exit polling, progress bar) is the
 * reference's, statement for statement: src/integrator.c:894-1121, src/integrator_whfast.c:200-407.
 */
#include <math.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "acceleration.h"
#include "common.h"
#include "error.h"
#include "output.h"
#include "progress_bar.h"
#include "settings.h"
#include "system.h"

#include "grav_b200.h"

#define RESIDENT_MAX_QUEUED_STEPS 64   /* steps enqueued between two looks at is_exit / the progress bar */

static int resident_enabled(void)
{
    const char *e = getenv("GRAV_B200_RESIDENT");
    return !(e && e[0] == '0');
}

static int resident_device(void)
{
    const char *e = getenv("GRAV_B200_DEVICE");
    return e ? atoi(e) : 0;
}

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
#define RC_STATUS(rc) status_from_rc((rc), __FILE__, __LINE__, __func__)
#define TRY_RC(call)                              \
    do                                            \
    {                                             \
        const int rc_ = (call);                   \
        if (rc_ != GRAV_B200_OK)                  \
        {                                         \
            error_status = RC_STATUS(rc_);        \
            goto done;                            \
        }                                         \
    } while (0)
#define TRY_STATUS(call)                                   \
    do                                                     \
    {                                                      \
        error_status = WRAP_TRACEBACK(call);               \
        if (error_status.return_code != GRAV_SUCCESS)      \
        {                                                  \
            goto done;                                     \
        }                                                  \
    } while (0)

/* ---- WHFast (src/integrator_whfast.c:200-407) ------------------------------------------------------------- */

static int whfast_download(grav_b200_ctx *ctx, System *system, const int snapshot)
{
    int n = 0;
    const int rc = grav_b200_ctx_whfast_get_state(ctx, snapshot, &n, system->particle_ids, system->x, system->v, system->m);
    if (rc == GRAV_B200_OK)
    {
        system->num_particles = n;     /* whfast_drift may have removed particles (:607-671) */
    }
    return rc;
}

int grav_b200_shim_whfast(ErrorStatus *out, System *system, IntegratorParam *integrator_param,
                          AccelerationParam *acceleration_param, OutputParam *output_param,
                          SimulationStatus *simulation_status, Settings *settings, const double tf)
{
    if (!resident_enabled())
    {
        return 0;
    }
    if (acceleration_param->method != ACCELERATION_METHOD_PAIRWISE &&
        acceleration_param->method != ACCELERATION_METHOD_MASSLESS)
    {
        return 0;   /* the reference's dispatcher raises its own error (:817-837) */
    }

    ErrorStatus error_status = make_success_error_status();
    grav_b200_ctx *ctx = NULL;
    double dt = integrator_param->dt;
    const bool is_output = (output_param->method != OUTPUT_METHOD_DISABLED);
    const double output_interval = output_param->output_interval;
    double next_output_time = output_interval;
    const bool enable_progress_bar = settings->enable_progress_bar;
    ProgressBarParam progress_bar_param;
    int64 queued = 0;
    double queued_dt = dt;

    TRY_RC(grav_b200_ctx_create(&ctx, resident_device(), 0, 1, NULL));
    TRY_RC(grav_b200_ctx_set_system(ctx, system->num_particles, system->x, system->v, system->m, system->G));
    /* sort by distance, eta, Jacobi coordinates, first acceleration, half kick (:241-273) */
    TRY_RC(grav_b200_ctx_whfast_begin(ctx, system->particle_ids, acceleration_param->method,
                                      acceleration_param->softening_length, dt,
                                      integrator_param->whfast_remove_invalid_particles));
    TRY_RC(whfast_download(ctx, system, 0));     /* the distance-sorted system, as the initial output sees it */
    if (is_output && output_param->output_initial)
    {
        TRY_STATUS(output_snapshot(output_param, system, integrator_param, acceleration_param, simulation_status, settings));
    }

    const int64 total_num_steps = (int64) ceil(tf / dt);
    if (enable_progress_bar)
    {
        TRY_STATUS(start_progress_bar(&progress_bar_param, total_num_steps));
    }
    simulation_status->t = 0.0;
    simulation_status->dt = dt;
    simulation_status->num_steps = 0;
    while (simulation_status->num_steps < total_num_steps)
    {
        if (simulation_status->t + dt > tf)      /* dt overshoot (:293-297) */
        {
            dt = tf - simulation_status->t;
        }
        simulation_status->dt = dt;
        if (queued > 0 && dt != queued_dt)
        {
            TRY_RC(grav_b200_ctx_whfast_steps(ctx, queued_dt, queued));
            queued = 0;
        }
        queued_dt = dt;
        queued++;

        (simulation_status->num_steps)++;
        simulation_status->t = (simulation_status->num_steps) * dt;

        const bool output_due = is_output && simulation_status->t >= next_output_time;
        const bool leaving = *(settings->is_exit_ptr) || simulation_status->num_steps == total_num_steps;
        if (output_due || leaving || queued >= RESIDENT_MAX_QUEUED_STEPS)
        {
            TRY_RC(grav_b200_ctx_whfast_steps(ctx, queued_dt, queued));
            queued = 0;
        }
        if (output_due)                          /* (:346-366) */
        {
            TRY_RC(whfast_download(ctx, system, 1));
            TRY_STATUS(output_snapshot(output_param, system, integrator_param, acceleration_param, simulation_status, settings));
            next_output_time = (output_param->output_count_) * output_interval;
        }
        if (enable_progress_bar)
        {
            update_progress_bar(&progress_bar_param, simulation_status->num_steps, false);
        }
        if (*(settings->is_exit_ptr))
        {
            break;
        }
    }
    if (enable_progress_bar)
    {
        update_progress_bar(&progress_bar_param, simulation_status->num_steps, true);
    }
    /* the reference leaves system->x / v as its last jacobi_to_cartesian wrote them; so does the device state */
    TRY_RC(whfast_download(ctx, system, 0));

done:
    if (ctx)
    {
        grav_b200_ctx_destroy(ctx);
    }
    *out = error_status;
    return 1;
}

/* ---- leapfrog (src/integrator.c:894-1121) ------------------------------------------------------------------ */

int grav_b200_shim_leapfrog(ErrorStatus *out, System *system, IntegratorParam *integrator_param,
                            AccelerationParam *acceleration_param, OutputParam *output_param,
                            SimulationStatus *simulation_status, Settings *settings, const double tf)
{
    if (!resident_enabled())
    {
        return 0;
    }
    const int method = acceleration_param->method;
    if (method != ACCELERATION_METHOD_PAIRWISE && method != ACCELERATION_METHOD_MASSLESS &&
        method != ACCELERATION_METHOD_BARNES_HUT)
    {
        return 0;
    }

    ErrorStatus error_status = make_success_error_status();
    grav_b200_ctx *ctx = NULL;
    double dt = integrator_param->dt;
    const bool is_output = (output_param->method != OUTPUT_METHOD_DISABLED);
    const double output_interval = output_param->output_interval;
    double next_output_time = output_interval;
    const bool enable_progress_bar = settings->enable_progress_bar;
    ProgressBarParam progress_bar_param;
    int64 queued = 0;
    double queued_dt = dt;

    if (is_output && output_param->output_initial)       /* (:944-960) */
    {
        TRY_STATUS(output_snapshot(output_param, system, integrator_param, acceleration_param, simulation_status, settings));
    }
    TRY_RC(grav_b200_ctx_create_auto(&ctx));   /* GRAV_B200_DEVICE, or a team of GRAV_B200_DEVICES GPUs driven from this thread */
    TRY_RC(grav_b200_ctx_set_system(ctx, system->num_particles, system->x, system->v, system->m, system->G));
    /* a(x0) and v_1/2 (:963-982) */
    TRY_RC(grav_b200_ctx_leapfrog_begin(ctx, method, acceleration_param->softening_length, acceleration_param->opening_angle,
                                        acceleration_param->max_num_particles_per_leaf, dt));

    const int64 total_num_steps = (int64) ceil(tf / dt);
    if (enable_progress_bar)
    {
        TRY_STATUS(start_progress_bar(&progress_bar_param, total_num_steps));
    }
    simulation_status->t = 0.0;
    simulation_status->dt = dt;
    simulation_status->num_steps = 0;
    while (simulation_status->num_steps < total_num_steps)
    {
        if (simulation_status->t + dt > tf)
        {
            dt = tf - simulation_status->t;
        }
        simulation_status->dt = dt;
        if (queued > 0 && dt != queued_dt)
        {
            TRY_RC(grav_b200_ctx_leapfrog_steps(ctx, queued_dt, queued));
            queued = 0;
        }
        queued_dt = dt;
        queued++;

        (simulation_status->num_steps)++;
        simulation_status->t = (simulation_status->num_steps) * dt;

        const bool output_due = is_output && simulation_status->t >= next_output_time;
        const bool leaving = *(settings->is_exit_ptr) || simulation_status->num_steps == total_num_steps;
        if (output_due || leaving || queued >= RESIDENT_MAX_QUEUED_STEPS)
        {
            TRY_RC(grav_b200_ctx_leapfrog_steps(ctx, queued_dt, queued));
            queued = 0;
        }
        if (output_due)      /* v_1 from v_1+1/2 for the snapshot, state untouched (:1045-1073) */
        {
            TRY_RC(grav_b200_ctx_get_positions(ctx, system->x));
            TRY_RC(grav_b200_ctx_get_velocities(ctx, system->v));
            TRY_STATUS(output_snapshot(output_param, system, integrator_param, acceleration_param, simulation_status, settings));
            next_output_time = (output_param->output_count_) * output_interval;
        }
        if (enable_progress_bar)
        {
            update_progress_bar(&progress_bar_param, simulation_status->num_steps, false);
        }
        if (*(settings->is_exit_ptr))
        {
            break;
        }
    }
    /* synchronise v_1+1/2 to v_1 (:1088-1094) and hand the state back */
    TRY_RC(grav_b200_ctx_leapfrog_end(ctx));
    TRY_RC(grav_b200_ctx_get_positions(ctx, system->x));
    TRY_RC(grav_b200_ctx_get_velocities(ctx, system->v));
    if (enable_progress_bar)
    {
        update_progress_bar(&progress_bar_param, simulation_status->num_steps, true);
    }

done:
    if (ctx)
    {
        grav_b200_ctx_destroy(ctx);
    }
    *out = error_status;
    return 1;
}


/* ---- Euler, Euler-Cromer, RK4 (src/integrator.c:281-454, :456-628, :630-892) --------------------------------- */

static int fixed_step_resident(ErrorStatus *out, const int integrator, System *system, IntegratorParam *integrator_param,
                               AccelerationParam *acceleration_param, OutputParam *output_param,
                               SimulationStatus *simulation_status, Settings *settings, const double tf)
{
    if (!resident_enabled())
    {
        return 0;
    }
    const int method = acceleration_param->method;
    if (method != ACCELERATION_METHOD_PAIRWISE && method != ACCELERATION_METHOD_MASSLESS &&
        method != ACCELERATION_METHOD_BARNES_HUT)
    {
        return 0;
    }

    ErrorStatus error_status = make_success_error_status();
    grav_b200_ctx *ctx = NULL;
    double dt = integrator_param->dt;
    const bool is_output = (output_param->method != OUTPUT_METHOD_DISABLED);
    const double output_interval = output_param->output_interval;
    double next_output_time = output_interval;
    const bool enable_progress_bar = settings->enable_progress_bar;
    const bool check_overshoot = (integrator != GRAV_B200_INTEGRATOR_RK4);   /* rk4() has no overshoot check (:736-741) */
    ProgressBarParam progress_bar_param;
    int64 queued = 0;
    double queued_dt = dt;

    if (is_output && output_param->output_initial)
    {
        TRY_STATUS(output_snapshot(output_param, system, integrator_param, acceleration_param, simulation_status, settings));
    }
    TRY_RC(grav_b200_ctx_create_auto(&ctx));   /* GRAV_B200_DEVICE, or a team of GRAV_B200_DEVICES GPUs driven from this thread */
    TRY_RC(grav_b200_ctx_set_system(ctx, system->num_particles, system->x, system->v, system->m, system->G));
    TRY_RC(grav_b200_ctx_fixed_begin(ctx, integrator, method, acceleration_param->softening_length,
                                     acceleration_param->opening_angle, acceleration_param->max_num_particles_per_leaf));

    const int64 total_num_steps = (int64) ceil(tf / dt);
    if (enable_progress_bar)
    {
        TRY_STATUS(start_progress_bar(&progress_bar_param, total_num_steps));
    }
    simulation_status->t = 0.0;
    simulation_status->dt = dt;
    simulation_status->num_steps = 0;
    while (simulation_status->num_steps < total_num_steps)
    {
        if (check_overshoot)
        {
            if (simulation_status->t + dt > tf)
            {
                dt = tf - simulation_status->t;
            }
            simulation_status->dt = dt;
        }
        if (queued > 0 && dt != queued_dt)
        {
            TRY_RC(grav_b200_ctx_fixed_steps(ctx, queued_dt, queued));
            queued = 0;
        }
        queued_dt = dt;
        queued++;

        (simulation_status->num_steps)++;
        simulation_status->t = (simulation_status->num_steps) * dt;

        const bool output_due = is_output && simulation_status->t >= next_output_time;
        const bool leaving = *(settings->is_exit_ptr) || simulation_status->num_steps == total_num_steps;
        if (output_due || leaving || queued >= RESIDENT_MAX_QUEUED_STEPS)
        {
            TRY_RC(grav_b200_ctx_fixed_steps(ctx, queued_dt, queued));
            queued = 0;
        }
        if (output_due)
        {
            TRY_RC(grav_b200_ctx_get_positions(ctx, system->x));
            TRY_RC(grav_b200_ctx_get_velocities(ctx, system->v));
            TRY_STATUS(output_snapshot(output_param, system, integrator_param, acceleration_param, simulation_status, settings));
            next_output_time = (output_param->output_count_) * output_interval;
        }
        if (enable_progress_bar)
        {
            update_progress_bar(&progress_bar_param, simulation_status->num_steps, false);
        }
        if (*(settings->is_exit_ptr))
        {
            break;
        }
    }
    TRY_RC(grav_b200_ctx_get_positions(ctx, system->x));
    TRY_RC(grav_b200_ctx_get_velocities(ctx, system->v));
    if (enable_progress_bar)
    {
        update_progress_bar(&progress_bar_param, simulation_status->num_steps, true);
    }

done:
    if (ctx)
    {
        grav_b200_ctx_destroy(ctx);
    }
    *out = error_status;
    return 1;
}

int grav_b200_shim_euler(ErrorStatus *out, System *system, IntegratorParam *integrator_param,
                         AccelerationParam *acceleration_param, OutputParam *output_param,
                         SimulationStatus *simulation_status, Settings *settings, const double tf)
{
    return fixed_step_resident(out, GRAV_B200_INTEGRATOR_EULER, system, integrator_param, acceleration_param, output_param,
                               simulation_status, settings, tf);
}

int grav_b200_shim_euler_cromer(ErrorStatus *out, System *system, IntegratorParam *integrator_param,
                                AccelerationParam *acceleration_param, OutputParam *output_param,
                                SimulationStatus *simulation_status, Settings *settings, const double tf)
{
    return fixed_step_resident(out, GRAV_B200_INTEGRATOR_EULER_CROMER, system, integrator_param, acceleration_param,
                               output_param, simulation_status, settings, tf);
}

int grav_b200_shim_rk4(ErrorStatus *out, System *system, IntegratorParam *integrator_param,
                       AccelerationParam *acceleration_param, OutputParam *output_param,
                       SimulationStatus *simulation_status, Settings *settings, const double tf)
{
    return fixed_step_resident(out, GRAV_B200_INTEGRATOR_RK4, system, integrator_param, acceleration_param, output_param,
                               simulation_status, settings, tf);
}

/* ---- energy diagnostic (src/utils.c:27-59, src/python_interface.c:195-243) ------------------------------------ */

#define ENERGY_MIN_PARTICLES 1024   /* below this the host pair loop takes microseconds; the hook declines */

/* Hook at the top of compute_energy():  { double e_; if (grav_b200_shim_compute_energy(&e_, system)) return e_; } */
int grav_b200_shim_compute_energy(double *energy, const System *system)
{
    if (!resident_enabled() || system->num_particles < ENERGY_MIN_PARTICLES)
    {
        return 0;
    }
    const int rc = grav_b200_compute_energy(energy, system->num_particles, system->x, system->v, system->m, system->G);
    if (rc != GRAV_B200_OK)
    {
        /* compute_energy() has no error channel: fail loudly instead of silently taking another path */
        fprintf(stderr, "compute_energy (B200): %s\n", grav_b200_last_error());
        *energy = NAN;
    }
    return 1;
}

/* Hook at the top of compute_energy_python(): snapshots are packed [m, x, y, z, vx, vy, vz] per particle. */
int grav_b200_shim_compute_energy_python(double *energy, const double G, const double *sol_state, const int num_snapshots,
                                         const int num_particles)
{
    if (!resident_enabled() || num_particles < ENERGY_MIN_PARTICLES)
    {
        return 0;
    }
    double *buf = malloc(sizeof(double) * 7 * (size_t) num_particles);
    if (!buf)
    {
        return 0;
    }
    double *x = buf, *v = buf + 3 * (size_t) num_particles, *m = buf + 6 * (size_t) num_particles;
    for (int s = 0; s < num_snapshots; s++)
    {
        const double *snap = sol_state + (size_t) s * 7 * (size_t) num_particles;
        for (int i = 0; i < num_particles; i++)
        {
            m[i] = snap[i * 7 + 0];
            for (int k = 0; k < 3; k++)
            {
                x[i * 3 + k] = snap[i * 7 + 1 + k];
                v[i * 3 + k] = snap[i * 7 + 4 + k];
            }
        }
        if (grav_b200_compute_energy(&energy[s], num_particles, x, v, m, G) != GRAV_B200_OK)
        {
            fprintf(stderr, "compute_energy_python (B200): %s\n", grav_b200_last_error());
            energy[s] = NAN;
        }
    }
    free(buf);
    return 1;
}
