"""The nested-dissection solver (level-0 chunks -> block cyclic reduction -> root) must produce the same LM step as the plain
sequential band factorisation (one chunk), for narrow and wide borders.  The chunk count is a developer switch read once per
process (ICC_SOLVER_CHUNKS), so each plan runs in its own subprocess (tests/solver_probe.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(cfg, chunks, flags, out):
    env = dict(os.environ, ICC_SOLVER_CHUNKS=str(chunks), PROBE_FLAGS=str(flags))
    subprocess.run([sys.executable, os.path.join(HERE, "solver_probe.py"), "--child", str(cfg), out], check=True, env=env)
    return np.load(out)


@pytest.mark.parametrize("cfg,flags", [(2, 66), (3, 66), (3, 66 | 4 | 16)], ids=["cfg2-stage1", "cfg3-stage1", "cfg3-biases-gravity"])
def test_chunk_plans_agree_with_sequential_factorisation(tmp_path, cfg, flags):
    ref = _run(cfg, 1, flags, str(tmp_path / "p1.npz"))
    assert int(ref["succ"]) == 1
    for chunks in (2, 5, 16, 31):
        d = _run(cfg, chunks, flags, str(tmp_path / f"p{chunks}.npz"))
        assert int(d["succ"]) == 1
        assert abs(float(d["cost"]) - float(ref["cost"])) <= 1e-9 * float(ref["cost"])
        for k in ("so3", "r3", "T"):
            assert np.abs(d[k] - ref[k]).max() < 1e-10, (chunks, k)
