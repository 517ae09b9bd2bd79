"""N > 1 host logic on CPU: two gloo ranks each assemble their time-slice shard, the all-reduced partial normal equations
(J^T J, J^T r, cost) equal the single-process ones.  The per-shard evaluator here is the oracle — the CUDA
path runs the same partition + reduction on the GPU box (tests/test_gpu_multi.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import F_STAGE1
    from openimucameracalibrator_b200 import _capi as capi
    from openimucameracalibrator_b200 import synthetic as syn
    from openimucameracalibrator_b200.distributed import shard_bounds
    from oracle_api import new_oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ds = syn.make_dataset(syn.tiny_config())
    o = new_oracle(1); capi.load_dataset(o, ds, shard=(rank, world))
    c, r, g, H = o.evaluate(F_STAGE1, hessian=True)
    lo, hi = shard_bounds(ds, rank, world)
    assert sum(o.num_residuals()) == hi - lo
    packed = torch.from_numpy(np.concatenate([H.ravel(), g, [c]]))
    dist.all_reduce(packed)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), packed.numpy())
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_full_problem(tmp_path, oracle_factory):
    from helpers import F_STAGE1, rel
    from openimucameracalibrator_b200 import _capi as capi
    from openimucameracalibrator_b200 import synthetic as syn
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    red = np.load(tmp_path / "reduced.npy")
    ds = syn.make_dataset(syn.tiny_config())
    o = oracle_factory(1); capi.load_dataset(o, ds)
    c, r, g, H = o.evaluate(F_STAGE1, hessian=True)
    n = g.size
    assert rel(red[: n * n].reshape(n, n), H) < 1e-12
    assert rel(red[n * n: n * n + n], g) < 1e-12
    assert abs(red[-1] - c) <= 1e-12 * c
