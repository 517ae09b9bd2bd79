"""Developer probe: kernel-family timings on BASELINE config 4 (not a test)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from openimucameracalibrator_b200 import _capi as capi, calibrator, synthetic as syn
F = int(os.environ.get('PROBE_FLAGS', capi.FLAG_SPLINE | capi.FLAG_T_I_C))
cfgs = [int(a) for a in sys.argv[1:]] or [4]
for c in cfgs:
    ds = syn.make_dataset(syn.CONFIGS[c])
    g = capi.CApi(calibrator.load_library(), "icc_", 0); capi.load_dataset(g, ds)
    g.time_evaluations(3, F, 1)
    vis, imu, cost, jac = g.time_evaluations(20, F, 2), g.time_evaluations(20, F, 3), g.time_evaluations(20, F, 0), g.time_evaluations(20, F, 1)
    s = g.lm_iterations(3, F)
    print(f"cfg{c}: vision {vis*1e3:.1f} us  imu {imu*1e3:.1f} us  cost {cost*1e3:.1f} us  jac(all+memset) {jac*1e3:.1f} us | solve {1e6*s.seconds_linear_solve/s.iterations:.1f} us/iter  jac_in_lm {1e6*s.seconds_jacobian/max(1,s.jacobian_evaluations):.1f} us  lm3 wall {s.seconds_total*1e3:.2f} ms")
