"""The oracle's forward-mode duals x local-parameterisation Jacobians (the restatement of Ceres autodiff) against central
finite differences of its own cost, for every parameter block type and every camera model."""
import numpy as np
import pytest

from helpers import F_STAGE1, F_ALL, TangentWalker
from openimucameracalibrator_b200 import _capi as capi
from openimucameracalibrator_b200 import synthetic as syn


def _check(o, flags, seed, tol=2e-6):
    _, _, g, _ = o.evaluate(flags)
    w = TangentWalker(o, flags)
    rng = np.random.default_rng(seed)
    for _ in range(3):
        d = rng.normal(size=w.n)
        # the line-delay gradient is ~1e7 larger than the rest: probe it with a proportionally smaller perturbation
        eps = 1e-7
        if (flags & capi.FLAG_CAM_LINE_DELAY):
            d[o.num_tangent(flags & ~(capi.FLAG_IMU_BIASES)) - 1] *= 1e-4 if flags & capi.FLAG_SPLINE else 1.0
        fd = (w.cost(eps * d) - w.cost(-eps * d)) / (2 * eps)
        an = float(g @ d)
        assert abs(fd - an) <= tol * max(abs(fd), abs(an)), (fd, an)
    w.restore()


@pytest.mark.parametrize("k", range(7))
def test_gradient_all_blocks_each_model(oracle_factory, k):
    cfg = syn.config5(k); cfg.n_frames = 12
    ds = syn.make_dataset(cfg)
    o = oracle_factory(); capi.load_dataset(o, ds, known_gravity=False)
    _check(o, F_ALL, seed=k)


def test_gradient_line_delay_only(oracle_factory):
    ds = syn.make_dataset(syn.tiny_config())
    o = oracle_factory(); capi.load_dataset(o, ds)
    flags = capi.FLAG_CAM_LINE_DELAY
    _, _, g, _ = o.evaluate(flags)
    w = TangentWalker(o, flags)
    eps = 1e-9
    fd = (w.cost(np.array([eps])) - w.cost(np.array([-eps]))) / (2 * eps)
    assert abs(fd - g[0]) <= 1e-6 * abs(fd)


def test_gradient_with_board_points_as_parameters(oracle_factory):
    """SplineOptimFlags::POINTS (impl.h:136-152): board points become 4-vector blocks with ceres::HomogeneousVectorParameterization(4);
    the oracle's Jet Jacobian x Householder local Jacobian against central differences through the NumPy Plus."""
    ds = syn.make_dataset(syn.tiny_config(n_frames=10))
    o = oracle_factory(); capi.load_dataset(o, ds)
    flags = F_STAGE1 | capi.FLAG_POINTS
    n = o.num_tangent(flags)
    assert n == o.num_tangent(F_STAGE1) + 3 * len(ds["board_xyzw"])
    _, _, g, _ = o.evaluate(flags)
    w = TangentWalker(o, flags)
    rng = np.random.default_rng(5)
    for trial in range(4):
        d = np.zeros(n)
        if trial < 3:
            d[n - 3 * len(ds["board_xyzw"]):] = rng.normal(size=3 * len(ds["board_xyzw"]))     # point directions only
        else:
            d = rng.normal(size=n)
        eps = 1e-7
        fd = (w.cost(eps * d) - w.cost(-eps * d)) / (2 * eps)
        assert abs(fd - g @ d) <= 2e-5 * max(abs(fd), abs(g @ d)), (trial, fd, g @ d)
    w.restore()
