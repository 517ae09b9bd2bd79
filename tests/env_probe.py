"""Developer probe (not a test): kernel-family timings of one BASELINE config under environment switches.
usage: env_probe.py <config> [NAME=VALUE ...]   (each run in a fresh process)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from openimucameracalibrator_b200 import _capi as capi, calibrator, synthetic as syn
    F = capi.FLAG_SPLINE | capi.FLAG_T_I_C
    c = int(sys.argv[2])
    ds = syn.make_dataset(syn.CONFIGS[c])
    g = capi.CApi(calibrator.load_library(), "icc_", 0); capi.load_dataset(g, ds)
    g.time_evaluations(3, F, 1)
    vis, imu, jac, cost = (1e3 * g.time_evaluations(20, F, m) for m in (2, 3, 1, 0))
    s = g.lm_iterations(3, F)
    print(f"cfg{c} {' '.join(sys.argv[3:]) or 'default':40s} vision {vis:7.1f} imu {imu:6.1f} jac {jac:7.1f} cost {cost:6.1f} us | solve {1e6*s.seconds_linear_solve/s.iterations:6.1f} us/iter lm3 wall {s.seconds_total*1e3:.2f} ms", flush=True)
else:
    cfg = sys.argv[1]
    for variant in " ".join(sys.argv[2:]).split("--") if len(sys.argv) > 2 else [""]:
        env = dict(os.environ)
        kv = variant.split()
        for x in kv:
            k, v = x.split("="); env[k] = v
        subprocess.run([sys.executable, __file__, "--child", cfg] + kv, env=env)
