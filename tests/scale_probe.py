"""Developer probe: a sequence several times longer than BASELINE config 4 (robustness of the plan / shared-memory sizing)."""
import sys, os, time, dataclasses
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from openimucameracalibrator_b200 import _capi as capi, calibrator, synthetic as syn
mult = int(sys.argv[1]) if len(sys.argv) > 1 else 5
flags = int(sys.argv[2]) if len(sys.argv) > 2 else (capi.FLAG_SPLINE | capi.FLAG_T_I_C)
cfg = dataclasses.replace(syn.CONFIGS[4], n_frames=3000 * mult, name=f"cfg4_x{mult}")
t0 = time.time(); ds = syn.make_dataset(cfg); print(f"dataset {time.time() - t0:.1f} s: frames {cfg.n_frames}, imu {ds['imu_t'].size}")
g = capi.CApi(calibrator.load_library(), "icc_", 0); capi.load_dataset(g, ds, known_gravity=not (flags & 16))
print("tangent", g.num_tangent(flags), "residuals", g.num_residuals())
s = g.lm_iterations(3, flags)
print(f"3 LM iterations: successful {s.successful_steps}, cost {s.initial_cost:.4e} -> {s.final_cost:.4e}, solve {1e6 * s.seconds_linear_solve / s.iterations:.1f} us/iter, jac {1e6 * s.seconds_jacobian / max(1, s.jacobian_evaluations):.1f} us, total {1e3 * s.seconds_total:.2f} ms")
s2 = g.optimize(50, flags)
print(f"optimize: iterations {s2.iterations} termination {s2.termination} final cost {s2.final_cost:.4e} reproj {s2.mean_reproj_error:.4f} px")
T = g.get_T_i_c(); print("T_i_c err", np.abs(T - ds["truth"]["T_i_c"]).max() if np.dot(T[:4], ds["truth"]["T_i_c"][:4]) > 0 else np.abs(T[:4] + ds["truth"]["T_i_c"][:4]).max())
