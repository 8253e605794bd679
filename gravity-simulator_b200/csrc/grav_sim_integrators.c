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

/* ---- the time loop shared by every resident integrator ------------------------------------------------------
 * Bookkeeping of src/integrator.c:996-1086 (leapfrog; euler :351-441, euler_cromer :525-614, rk4 :736-880) and
 * src/integrator_whfast.c:286-383, which are the same statements: dt overshoot, t = num_steps * dt, output when t passes
 * the next output time, progress bar, is_exit.  Steps are queued on the device and flushed when the host has to look
 * (output due, leaving, RESIDENT_MAX_QUEUED_STEPS reached, dt changed). */
typedef struct ResidentOps
{
    int (*steps)(grav_b200_ctx *ctx, double dt, int64_t num_steps);
    int (*download)(grav_b200_ctx *ctx, System *system, int snapshot);   /* host arrays of `system` <- device state */
    int (*finish)(grav_b200_ctx *ctx);                                    /* after the last step, before the final download; may be NULL */
    bool check_overshoot;                                                 /* rk4() has no overshoot check (:736-741) */
} ResidentOps;

static int download_xv(grav_b200_ctx *ctx, System *system, const int snapshot)
{
    (void) snapshot;   /* while a leapfrog runs get_velocities() already applies the snapshot convention (:1045-1073) */
    const int rc = grav_b200_ctx_get_positions(ctx, system->x);
    return rc != GRAV_B200_OK ? rc : grav_b200_ctx_get_velocities(ctx, system->v);
}

static ErrorStatus resident_time_loop(grav_b200_ctx *ctx, const ResidentOps *ops, System *system, IntegratorParam *integrator_param,
                                      AccelerationParam *acceleration_param, OutputParam *output_param,
                                      SimulationStatus *simulation_status, Settings *settings, const double tf)
{
    ErrorStatus error_status = make_success_error_status();
    double dt = integrator_param->dt;
    const bool is_output = (output_param->method != OUTPUT_METHOD_DISABLED);
    const double output_interval = output_param->output_interval;
    double next_output_time = output_interval;
    const bool enable_progress_bar = settings->enable_progress_bar;
    ProgressBarParam progress_bar_param;
    int64 queued = 0;
    double queued_dt = dt;

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
        if (ops->check_overshoot)
        {
            if (simulation_status->t + dt > tf)
            {
                dt = tf - simulation_status->t;
            }
            simulation_status->dt = dt;
        }
        if (queued > 0 && dt != queued_dt)
        {
            TRY_RC(ops->steps(ctx, queued_dt, queued));
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
            TRY_RC(ops->steps(ctx, queued_dt, queued));
            queued = 0;
        }
        if (output_due)
        {
            TRY_RC(ops->download(ctx, system, 1));
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
    if (ops->finish)
    {
        TRY_RC(ops->finish(ctx));
    }
    TRY_RC(ops->download(ctx, system, 0));
    if (enable_progress_bar)
    {
        update_progress_bar(&progress_bar_param, simulation_status->num_steps, true);
    }
done:
    return error_status;
}

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
    /* the reference leaves system->x / v as its last jacobi_to_cartesian wrote them; so does the final download (snapshot = 0);
     * outputs between steps use the -dt/2 snapshot convention (:346-366) */
    const ResidentOps ops = {grav_b200_ctx_whfast_steps, whfast_download, NULL, true};

    TRY_RC(grav_b200_ctx_create(&ctx, resident_device(), 0, 1, NULL));   /* one GPU by design: DESIGN.md section 5 */
    TRY_RC(grav_b200_ctx_set_system(ctx, system->num_particles, system->x, system->v, system->m, system->G));
    /* sort by distance, eta, Jacobi coordinates, first acceleration, half kick (:241-273) */
    TRY_RC(grav_b200_ctx_whfast_begin(ctx, system->particle_ids, acceleration_param->method,
                                      acceleration_param->softening_length, integrator_param->dt,
                                      integrator_param->whfast_remove_invalid_particles));
    TRY_RC(grav_b200_ctx_whfast_set_verbose(ctx, settings->verbose));   /* the removal message of whfast_drift (:609-623) */
    TRY_RC(whfast_download(ctx, system, 0));     /* the distance-sorted system, as the initial output sees it */
    if (output_param->method != OUTPUT_METHOD_DISABLED && output_param->output_initial)
    {
        TRY_STATUS(output_snapshot(output_param, system, integrator_param, acceleration_param, simulation_status, settings));
    }
    error_status = resident_time_loop(ctx, &ops, system, integrator_param, acceleration_param, output_param, simulation_status,
                                      settings, tf);

done:
    if (ctx)
    {
        grav_b200_ctx_destroy(ctx);
    }
    *out = error_status;
    return 1;
}

/* ---- leapfrog (src/integrator.c:894-1121), Euler, Euler-Cromer, RK4 (:281-454, :456-628, :630-892) ----------- */

/* integrator < 0: leapfrog */
static int xv_resident(ErrorStatus *out, const int integrator, System *system, IntegratorParam *integrator_param,
                       AccelerationParam *acceleration_param, OutputParam *output_param, SimulationStatus *simulation_status,
                       Settings *settings, const double tf)
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
    const bool is_leapfrog = integrator < 0;
    /* leapfrog: v_1 from v_1+1/2 for snapshots with the state untouched (:1045-1073), and for good at the end (:1088-1094) */
    const ResidentOps ops = {is_leapfrog ? grav_b200_ctx_leapfrog_steps : grav_b200_ctx_fixed_steps, download_xv,
                             is_leapfrog ? grav_b200_ctx_leapfrog_end : NULL, integrator != GRAV_B200_INTEGRATOR_RK4};

    if (output_param->method != OUTPUT_METHOD_DISABLED && output_param->output_initial)       /* (:944-960) */
    {
        TRY_STATUS(output_snapshot(output_param, system, integrator_param, acceleration_param, simulation_status, settings));
    }
    TRY_RC(grav_b200_ctx_create_auto(&ctx));   /* GRAV_B200_DEVICE, or a team of GRAV_B200_DEVICES GPUs driven from this thread */
    TRY_RC(grav_b200_ctx_set_system(ctx, system->num_particles, system->x, system->v, system->m, system->G));
    if (is_leapfrog)
    {
        /* a(x0) and v_1/2 (:963-982) */
        TRY_RC(grav_b200_ctx_leapfrog_begin(ctx, method, acceleration_param->softening_length, acceleration_param->opening_angle,
                                            acceleration_param->max_num_particles_per_leaf, integrator_param->dt));
    }
    else
    {
        TRY_RC(grav_b200_ctx_fixed_begin(ctx, integrator, method, acceleration_param->softening_length,
                                         acceleration_param->opening_angle, acceleration_param->max_num_particles_per_leaf));
    }
    error_status = resident_time_loop(ctx, &ops, system, integrator_param, acceleration_param, output_param, simulation_status,
                                      settings, tf);

done:
    if (ctx)
    {
        grav_b200_ctx_destroy(ctx);
    }
    *out = error_status;
    return 1;
}

int grav_b200_shim_leapfrog(ErrorStatus *out, System *system, IntegratorParam *integrator_param,
                            AccelerationParam *acceleration_param, OutputParam *output_param,
                            SimulationStatus *simulation_status, Settings *settings, const double tf)
{
    return xv_resident(out, -1, system, integrator_param, acceleration_param, output_param, simulation_status, settings, tf);
}

int grav_b200_shim_euler(ErrorStatus *out, System *system, IntegratorParam *integrator_param,
                         AccelerationParam *acceleration_param, OutputParam *output_param,
                         SimulationStatus *simulation_status, Settings *settings, const double tf)
{
    return xv_resident(out, GRAV_B200_INTEGRATOR_EULER, system, integrator_param, acceleration_param, output_param,
                               simulation_status, settings, tf);
}

int grav_b200_shim_euler_cromer(ErrorStatus *out, System *system, IntegratorParam *integrator_param,
                                AccelerationParam *acceleration_param, OutputParam *output_param,
                                SimulationStatus *simulation_status, Settings *settings, const double tf)
{
    return xv_resident(out, GRAV_B200_INTEGRATOR_EULER_CROMER, system, integrator_param, acceleration_param,
                               output_param, simulation_status, settings, tf);
}

int grav_b200_shim_rk4(ErrorStatus *out, System *system, IntegratorParam *integrator_param,
                       AccelerationParam *acceleration_param, OutputParam *output_param,
                       SimulationStatus *simulation_status, Settings *settings, const double tf)
{
    return xv_resident(out, GRAV_B200_INTEGRATOR_RK4, system, integrator_param, acceleration_param, output_param,
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
