"""Probe (not a test): does the solver take the POINTS border of the BASELINE boards?  python tests/points_probe.py"""
import dataclasses, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from helpers import F_STAGE1
from openimucameracalibrator_b200 import _capi as capi, calibrator, synthetic as syn
lib = calibrator.load_library()
for cfg_id, nf in ((2, 60), (2, 300), (4, 300), (4, 3000)):
    cfg = dataclasses.replace(syn.CONFIGS[cfg_id], n_frames=nf)
    ds = syn.make_dataset(cfg)
    g = capi.CApi(lib, "icc_", 0); capi.load_dataset(g, ds)
    flags = F_STAGE1 | capi.FLAG_POINTS
    try:
        t = time.time(); s = g.optimize(50, flags); dt = time.time() - t
        print(f"cfg{cfg_id} x{nf}: nb={6 + 3 * len(ds['board_xyzw'])} iterations {s.iterations} cost {s.initial_cost:.6e} -> {s.final_cost:.6e} term {s.termination} "
              f"jac {1e3 * s.seconds_jacobian / max(1, s.jacobian_evaluations):.3f} ms solve {1e3 * s.seconds_linear_solve / max(1, s.iterations):.3f} ms total {dt:.3f} s", flush=True)
    except capi.IccError as e:
        print(f"cfg{cfg_id} x{nf}: {e}", flush=True)
    g.close()
