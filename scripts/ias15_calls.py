"""Config 1 through the drop-in: wall time, number of acceleration() calls and time per call (launch path counts one
kernel launch per call).  GRAV_B200_SMALL_MAILBOX=0/1 selects the small-system path."""
import os
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_package
gb = load_package()
from gravity_simulator_b200 import ics
from oracle.bind import launch_simulation, DROPIN_SO, REF_SO
yrs = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
x, v, m, G = ics.solar_system()
kw = dict(tf=yrs * 365.24, integrator="ias15", tolerance=1e-9, method="pairwise")
launch_simulation(DROPIN_SO, x, v, m, G, **dict(kw, tf=365.24))
n0 = gb.kernel_launch_count()
t0 = time.perf_counter(); launch_simulation(DROPIN_SO, x, v, m, G, **kw); t1 = time.perf_counter()
calls = gb.kernel_launch_count() - n0
t2 = time.perf_counter(); launch_simulation(REF_SO, x, v, m, G, **kw); t3 = time.perf_counter()
print(f"mailbox={os.environ.get('GRAV_B200_SMALL_MAILBOX', '1')}: drop-in {t1 - t0:.3f} s, reference {t3 - t2:.3f} s for {yrs} yr; "
      f"kernel launches {calls}" + (f" -> {(t1 - t0) / calls * 1e6:.1f} us per call" if calls > 1000 else ""))
