#!/usr/bin/env python
"""Workload for `ncu`: load BASELINE config 4 and run two LM iterations (nothing else launches kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openimucameracalibrator_b200 import _capi as capi, calibrator, synthetic as syn
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ds = syn.make_dataset(syn.CONFIGS[cfg])
g = capi.CApi(calibrator.load_library(), "icc_", 0); capi.load_dataset(g, ds)
s = g.lm_iterations(2, capi.FLAG_SPLINE | capi.FLAG_T_I_C)
print("iterations", s.iterations, "successful", s.successful_steps, "final cost", s.final_cost)
