"""Probe (not a test): run-to-run spread of the tight-tolerance optimum (atomics reorder the sums).  python tests/tight_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import F_STAGE1
from oracle_api import new_oracle
from openimucameracalibrator_b200 import _capi as capi, calibrator, synthetic as syn
ds = syn.make_dataset(syn.tiny_config(n_frames=40))
o = new_oracle(); capi.load_dataset(o, ds); o.set_solver_options(function_tolerance=1e-14, parameter_tolerance=1e-14)
so = o.optimize(60, F_STAGE1)
print("oracle", so.iterations, so.successful_steps, so.termination, repr(so.final_cost))
for rep in range(12):
    g = capi.CApi(calibrator.load_library(), "icc_", 0); capi.load_dataset(g, ds); g.set_solver_options(function_tolerance=1e-14, parameter_tolerance=1e-14)
    sg = g.optimize(60, F_STAGE1)
    print("gpu   ", sg.iterations, sg.successful_steps, sg.termination, repr(sg.final_cost), f"{abs(sg.final_cost - so.final_cost) / so.final_cost:.2e}", f"{np.abs(g.get_T_i_c() - o.get_T_i_c()).max():.2e}")
    g.close()
