"""Levenberg-Marquardt parity: the CUDA driver follows the same Ceres-style trust-region path as the oracle, so iteration
counts agree and the converged T_cam_imu / line delay agree far inside the north-star's 1e-4 relative bar."""
import numpy as np
import pytest

from helpers import F_ALL, F_STAGE1, F_STAGE2, rel
from openimucameracalibrator_b200 import _capi as capi
from openimucameracalibrator_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _run_both(oracle_factory, gpu_factory, ds, flags, iters=50, **kw):
    o = oracle_factory(); capi.load_dataset(o, ds, **kw)
    g = gpu_factory(); capi.load_dataset(g, ds, **kw)
    return o, g, o.optimize(iters, flags), g.optimize(iters, flags)


def test_config1_two_stage_calibration(oracle_factory, gpu_factory, eval_path):
    ds = syn.make_dataset(syn.CONFIGS[1])
    o, g, so, sg = _run_both(oracle_factory, gpu_factory, ds, F_STAGE1)
    assert sg.iterations == so.iterations and sg.termination == so.termination and sg.successful_steps == so.successful_steps
    assert abs(sg.final_cost - so.final_cost) <= 1e-9 * so.final_cost
    assert rel(g.get_T_i_c(), o.get_T_i_c()) < 1e-8            # north-star bar: 1e-4
    assert abs(sg.mean_reproj_error - so.mean_reproj_error) < 1e-9
    s2o, s2g = o.optimize(10, F_STAGE2), g.optimize(10, F_STAGE2)
    assert s2g.iterations == s2o.iterations
    assert abs(g.get_line_delay() - o.get_line_delay()) <= 1e-8 * abs(o.get_line_delay())
    so3o, r3o, _, _ = o.get_knots(); so3g, r3g, _, _ = g.get_knots()
    assert rel(so3g, so3o) < 1e-8 and rel(r3g, r3o) < 1e-8


def test_config2_fisheye(oracle_factory, gpu_factory, eval_path):
    ds = syn.make_dataset(syn.CONFIGS[2])
    o, g, so, sg = _run_both(oracle_factory, gpu_factory, ds, F_STAGE1)
    assert sg.num_tangent == 1236 and sg.num_residuals == so.num_residuals
    assert sg.iterations == so.iterations and sg.termination == so.termination
    assert rel(g.get_T_i_c(), o.get_T_i_c()) < 1e-8
    assert sg.gpu_launches >= 4 * sg.iterations


def test_biases_and_gravity_free(oracle_factory, gpu_factory):
    ds = syn.make_dataset(syn.tiny_config(n_frames=40, imu_rate_hz=200.0))
    o, g, so, sg = _run_both(oracle_factory, gpu_factory, ds, F_STAGE1 | capi.FLAG_IMU_BIASES | capi.FLAG_GRAVITY_DIR, known_gravity=False)
    assert sg.iterations == so.iterations
    assert rel(g.get_T_i_c(), o.get_T_i_c()) < 1e-7 and rel(g.get_gravity(), o.get_gravity()) < 1e-7
    for a, b in zip(g.get_knots(), o.get_knots()):
        assert rel(a, b) < 1e-7


def test_line_delay_stage_recovers_perturbed_init(oracle_factory, gpu_factory):
    """Config-3 shape (DoubleSphere + rolling-shutter line-delay stage), shortened: init line delay is 10 % off."""
    import copy
    cfg = copy.copy(syn.CONFIGS[3]); cfg.n_frames = 90
    ds = syn.make_dataset(cfg)
    o, g, so, sg = _run_both(oracle_factory, gpu_factory, ds, F_STAGE1)
    assert sg.iterations == so.iterations
    s2o, s2g = o.optimize(10, F_STAGE2), g.optimize(10, F_STAGE2)
    assert s2g.iterations == s2o.iterations and s2g.num_tangent == 1
    assert abs(g.get_line_delay() - o.get_line_delay()) <= 1e-7 * abs(o.get_line_delay())
    assert rel(g.get_T_i_c(), o.get_T_i_c()) < 1e-7


def test_tight_tolerance_optimum_agrees(oracle_factory, gpu_factory):
    """Compare far down the valley (tolerances far below the reference's 1e-4) — SURVEY §7 'hard parts'.  Sixty LM iterations of this problem
    still creep along flat directions, and the order of the GPU's floating-point atomics differs from run to run: measured over 12 runs
    (tests/tight_probe.py) the final cost scatters by 2e-9 .. 3e-8 relative around the oracle's and T_i_c by 6e-7 .. 4e-6 -- the bars below
    leave a decade above that scatter and stay a decade below the north star's 1e-4."""
    ds = syn.make_dataset(syn.tiny_config(n_frames=40))
    o = oracle_factory(); capi.load_dataset(o, ds); o.set_solver_options(function_tolerance=1e-14, parameter_tolerance=1e-14)
    g = gpu_factory(); capi.load_dataset(g, ds); g.set_solver_options(function_tolerance=1e-14, parameter_tolerance=1e-14)
    so, sg = o.optimize(60, F_STAGE1), g.optimize(60, F_STAGE1)
    assert sg.iterations == so.iterations
    assert abs(sg.final_cost - so.final_cost) <= 3e-7 * so.final_cost
    assert rel(g.get_T_i_c(), o.get_T_i_c()) < 3e-5


def test_lm_iteration_schedule_and_launch_count(gpu_factory):
    ds = syn.make_dataset(syn.tiny_config())
    g = gpu_factory(); capi.load_dataset(g, ds)
    s = g.lm_iterations(3, F_STAGE1)
    assert s.iterations == 3 and s.cost_evaluations == 3 and s.jacobian_evaluations == min(3, 1 + s.successful_steps)
    assert s.gpu_launches >= 3 * 4 and s.seconds_jacobian > 0 and s.seconds_linear_solve > 0


def test_very_wide_border_uses_hbm_border_blocks(oracle_factory, gpu_factory):
    """A 200 s sequence with IMU_BIASES | GRAVITY_DIR free: 2 x 23 bias knots -> 147 border columns.  The level-0 local border
    block and the root's border block no longer fit in shared memory and live in HBM (SolvePlan::cl_global / cs_global); the LM
    path must still follow the oracle."""
    cfg = syn.tiny_config(n_frames=400, fps=2.0, dt_so3_s=0.5, dt_r3_s=0.5, imu_rate_hz=50.0, seed=41)
    ds = syn.make_dataset(cfg)
    flags = F_STAGE1 | capi.FLAG_IMU_BIASES | capi.FLAG_GRAVITY_DIR
    o = oracle_factory(); capi.load_dataset(o, ds, known_gravity=False)
    g = gpu_factory(); capi.load_dataset(g, ds, known_gravity=False)
    nb = g.num_tangent(flags) - 3 * g.num_knots()[0] - 3 * g.num_knots()[1]
    assert nb >= 140, nb
    so, sg = o.lm_iterations(3, flags), g.lm_iterations(3, flags)
    assert sg.successful_steps == so.successful_steps >= 1
    assert abs(sg.final_cost - so.final_cost) <= 1e-7 * so.final_cost
    assert rel(g.get_T_i_c(), o.get_T_i_c()) < 1e-7 and rel(g.get_gravity(), o.get_gravity()) < 1e-7
