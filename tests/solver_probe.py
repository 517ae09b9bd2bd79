"""Developer probe: one LM step with different solver chunkings must give the same step (not a test).
usage: solver_probe.py <config> <chunks...>   (spawns one process per chunk count; ICC_SOLVER_CHUNKS is read once per process)"""
import sys, os, subprocess, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
if len(sys.argv) >= 2 and sys.argv[1] == "--child":
    from openimucameracalibrator_b200 import _capi as capi, calibrator, synthetic as syn
    c = int(sys.argv[2])
    ds = syn.make_dataset(syn.CONFIGS[c]) if c > 0 else syn.make_dataset(syn.tiny_config())
    g = capi.CApi(calibrator.load_library(), "icc_", 0); capi.load_dataset(g, ds, known_gravity=not (int(os.environ.get('PROBE_FLAGS', 0)) & 16))
    s = g.lm_iterations(1, int(os.environ.get('PROBE_FLAGS', capi.FLAG_SPLINE | capi.FLAG_T_I_C)))
    so3, r3, _, _ = g.get_knots()
    np.savez(sys.argv[3], so3=so3, r3=r3, T=g.get_T_i_c(), succ=s.successful_steps, cost=s.final_cost)
    sys.exit(0)
cfg = int(sys.argv[1]); chunks = [int(a) for a in sys.argv[2:]]
ref = None
for ch in chunks:
    out = f"/tmp/probe_{cfg}_{ch}.npz"
    env = dict(os.environ, ICC_SOLVER_CHUNKS=str(ch))
    subprocess.run([sys.executable, __file__, "--child", str(cfg), out], check=True, env=env)
    d = np.load(out)
    if ref is None: ref = d; print(f"cfg{cfg} chunks {ch}: reference, successful {int(d['succ'])} cost {float(d['cost']):.12e}"); continue
    e1 = np.abs(d["so3"] - ref["so3"]).max(); e2 = np.abs(d["r3"] - ref["r3"]).max(); e3 = np.abs(d["T"] - ref["T"]).max()
    print(f"cfg{cfg} chunks {ch}: successful {int(d['succ'])} cost {float(d['cost']):.12e}  max|dso3| {e1:.2e} max|dr3| {e2:.2e} max|dT| {e3:.2e}")
