"""The oracle's Ceres-style LM behaves like the reference's documented runs: converges in a handful of iterations to
noise-level reprojection error (Readme.md:43-51 reports 0.6-0.9 px on real data; synthetic corner noise here is 0.2 px)."""
import numpy as np

from helpers import F_STAGE1, F_STAGE2
from openimucameracalibrator_b200 import _capi as capi
from openimucameracalibrator_b200 import synthetic as syn


def test_config1_converges(oracle_factory):
    ds = syn.make_dataset(syn.CONFIGS[1])
    o = oracle_factory(); capi.load_dataset(o, ds)
    assert o.num_knots()[:2] == (45, 45)
    assert o.num_residuals() == (4200, 1182, 1182)
    assert o.num_tangent(F_STAGE1) == 276
    s = o.optimize(50, F_STAGE1)
    assert s.termination in (1, 2) and 2 <= s.iterations <= 15
    assert s.final_cost < 1e-2 * s.initial_cost
    assert 0.15 < s.mean_reproj_error < 0.4                    # ~ sigma * sqrt(pi/2) for sigma = 0.2 px
    q = o.get_T_i_c()[:4]; qt = ds["truth"]["T_i_c"][:4]
    assert np.degrees(2 * np.arccos(min(1.0, abs(float(q @ qt))))) < 0.6   # started 1 deg off; 2 s of data, function_tolerance 1e-4
    s2 = o.optimize(10, F_STAGE2)
    assert s2.num_tangent == 1
    assert abs(o.get_line_delay() - ds["truth"]["line_delay"]) < 0.2 * ds["truth"]["line_delay"]


def test_lm_iterations_schedule(oracle_factory):
    """n iterations = n solves + n cost evaluations + one Jacobian evaluation per accepted step (lazy rebuild)."""
    ds = syn.make_dataset(syn.tiny_config())
    o = oracle_factory(); capi.load_dataset(o, ds)
    s = o.lm_iterations(3, F_STAGE1)
    assert s.iterations == 3 and s.cost_evaluations == 3
    assert s.jacobian_evaluations == min(3, 1 + s.successful_steps)
    assert s.final_cost < s.initial_cost


def test_global_shutter_zero_weights_vision(oracle_factory):
    """with a zero line delay the GS functor is wrapped in HuberLoss(0) => vision contributes nothing."""
    ds = dict(syn.make_dataset(syn.tiny_config()))
    ds["init_line_delay_s"] = 0.0
    o = oracle_factory(); capi.load_dataset(o, ds)
    assert o.num_residuals()[0] == 0
    assert o.num_tangent(F_STAGE1 | capi.FLAG_CAM_LINE_DELAY) == o.num_tangent(F_STAGE1)
    assert np.isfinite(o.mean_reprojection_error())
