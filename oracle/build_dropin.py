"""Build oracle/_ref/libgrav_sim_dropin.so: the reference's libgrav_sim with its acceleration path replaced
by ours.  TEST INFRASTRUCTURE (end-to-end integration tests through the reference's own integrators and
Python-facing entry points); needs /root/reference, so it only runs in the build container -- the GPU box
uses the prebuilt file.

Composition (exactly what INTEGRATION.md tells a maintainer to do):
  * every reference source file EXCEPT src/acceleration.c, src/acceleration_barnes_hut.c, src/linear_octree.c,
    compiled from where it lies, unmodified, with the project's flags;
  * gravity-simulator_b200/csrc/grav_sim_shim.c compiled against the reference's own headers
    (-DGRAV_SIM_USE_REFERENCE_HEADERS) in their place;
  * src/integrator_whfast.c with a ONE-LINE patch: its static dispatcher whfast_acceleration() forwards to
    grav_b200_shim_whfast_acceleration().  The patched text lives only in a temporary directory.
  * gravity-simulator_b200/csrc/grav_sim_integrators.c (device-resident leapfrog and WHFast time loops) plus a
    three-line hook at the top of the reference's leapfrog(), euler(), euler_cromer(), rk4() (src/integrator.c) and
    whfast(): the hook runs the
    resident loop and returns, or declines (GRAV_B200_RESIDENT=0) and lets the reference's own loop run.
  * the same kind of hook at the top of compute_energy() (src/utils.c) and compute_energy_python()
    (src/python_interface.c): the O(N^2) potential sum runs on the GPU for N >= 1024.
  * linked against libgrav_b200.so.
"""
import re
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
PKG = ROOT / "gravity-simulator_b200"
OUT = HERE / "_ref" / "libgrav_sim_dropin.so"

KEEP = ["cosmology.c", "error.c", "grav_sim.c", "integrator_rk_embedded.c", "integrator_ias15.c",
        "math_functions.c", "output.c", "progress_bar.c", "settings.c", "system.c"]
FLAGS = ["-std=gnu99", "-O3", "-fPIC", "-fopenmp", "-DUSE_OPENMP", '-DVERSION_INFO="0.0.4-b200"', f"-I{REF}/src", f"-I{REF}/pcg", "-w"]


def simple_hook(src, definition_regex, text):
    m = re.search(definition_regex, src)
    if not m:
        raise SystemExit(f"could not locate {definition_regex!r} in the reference")
    return src[:m.end()] + "\n" + text + src[m.end():]


def resident_hook(src, definition_regex, hook_fn):
    """Insert the three-line hook right after the opening brace of a time-loop function's DEFINITION."""
    m = re.search(definition_regex, src)
    if not m:
        raise SystemExit(f"could not locate the definition for {hook_fn} in the reference")
    hook = (f"\n    extern int {hook_fn}(ErrorStatus *, System *, IntegratorParam *, AccelerationParam *, OutputParam *,"
            " SimulationStatus *, Settings *, const double);\n"
            f"    {{ ErrorStatus es_; if ({hook_fn}(&es_, system, integrator_param, acceleration_param, output_param,"
            " simulation_status, settings, tf)) return es_; }\n")
    return src[:m.end()] + hook + src[m.end():]


def main():
    if not (REF / "src").is_dir():
        print(f"{REF} absent: keeping prebuilt {OUT.name} (if any)")
        return
    if not (PKG / "libgrav_b200.so").exists():
        raise SystemExit("build gravity-simulator_b200/libgrav_b200.so first")
    OUT.parent.mkdir(exist_ok=True)
    src = (REF / "src" / "integrator_whfast.c").read_text()
    # the definition (not the prototype) of the static dispatcher: ") {" follows the parameter list
    m = re.search(r"IN_FILE ErrorStatus whfast_acceleration\(\s*double \*restrict a,[^)]*\)\s*\{", src)
    if not m:
        raise SystemExit("could not locate whfast_acceleration() in the reference")
    hook = ("\n    extern ErrorStatus grav_b200_shim_whfast_acceleration(double *restrict, const System *, const double *restrict,"
            " const double *restrict, const AccelerationParam *);\n"
            "    return grav_b200_shim_whfast_acceleration(a, system, jacobi_x, eta, acceleration_param);\n")
    patched = src[:m.end()] + hook + src[m.end():]
    patched = resident_hook(patched, r"WIN32DLL_API ErrorStatus whfast\(\s*System \*system,[^)]*\)\s*\{", "grav_b200_shim_whfast")
    integ = (REF / "src" / "integrator.c").read_text()
    for fn in ("leapfrog", "euler", "euler_cromer", "rk4"):
        integ = resident_hook(integ, r"IN_FILE ErrorStatus " + fn + r"\(\s*System \*system,[^)]*\)\s*\{", "grav_b200_shim_" + fn)
    utils = simple_hook((REF / "src" / "utils.c").read_text(),
                        r"WIN32DLL_API double compute_energy\(const System \*restrict system\)\s*\{",
                        "    extern int grav_b200_shim_compute_energy(double *, const System *);\n"
                        "    { double e_; if (grav_b200_shim_compute_energy(&e_, system)) return e_; }\n")
    pyif = simple_hook((REF / "src" / "python_interface.c").read_text(),
                       r"WIN32DLL_API void compute_energy_python\([^)]*\)\s*\{",
                       "    extern int grav_b200_shim_compute_energy_python(double *, const double, const double *, const int, const int);\n"
                       "    if (grav_b200_shim_compute_energy_python(energy, G, sol_state, num_snapshots, num_particles)) return;\n")
    with tempfile.TemporaryDirectory() as tmp:
        pw = Path(tmp) / "integrator_whfast_patched.c"
        pw.write_text(patched)
        pi = Path(tmp) / "integrator_patched.c"
        pi.write_text(integ)
        pu = Path(tmp) / "utils_patched.c"
        pu.write_text(utils)
        pp = Path(tmp) / "python_interface_patched.c"
        pp.write_text(pyif)
        cmd = (["/usr/bin/gcc"] + FLAGS + ["-shared", "-o", str(OUT)] + [str(REF / "src" / f) for f in KEEP]
               + [str(pw), str(pi), str(pu), str(pp), str(REF / "pcg" / "pcg_basic.c")]
               + ["-DGRAV_SIM_USE_REFERENCE_HEADERS", f"-I{ROOT}/include", str(PKG / "csrc" / "grav_sim_shim.c"),
                  str(PKG / "csrc" / "grav_sim_integrators.c")]
               + [f"-L{PKG}", "-lgrav_b200", f"-Wl,-rpath,{PKG}", "-Wl,-rpath,$ORIGIN/../../gravity-simulator_b200", "-lm", "-lrt"])
        subprocess.run(cmd, check=True)
    print(f"built {OUT}")


if __name__ == "__main__":
    main()
