"""Developer probe (not a test): thread scaling of the oracle Jacobian evaluation.  usage: oracle_scaling_probe.py <config> <threads>..."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from openimucameracalibrator_b200 import _capi as capi, synthetic as syn
from oracle_api import new_oracle
import numpy as np
F = capi.FLAG_SPLINE | capi.FLAG_T_I_C
cfg = int(sys.argv[1]); 
ds = syn.make_dataset(syn.CONFIGS[cfg])
for nt in [int(a) for a in sys.argv[2:]] or [0]:
    o = new_oracle(nt) if nt else new_oracle()
    capi.load_dataset(o, ds)
    nres = sum(o.num_residuals())
    o.time_evaluations(1, F, 1)
    ms = o.time_evaluations(2, F, 1)
    print(f"cfg{cfg} threads {nt or 'all'}: {ms:.1f} ms/eval  {nres/ms/1e3:.3f} M res/s", flush=True)
