"""Camera-model restatements agree: oracle (templated C++, autodiff-able) vs the NumPy generator (independent code)."""
import ctypes

import numpy as np
import pytest

from openimucameracalibrator_b200 import camera_models as cm
from openimucameracalibrator_b200 import synthetic as syn
from oracle_api import oracle_lib

CASES = [(m, np.array(k, dtype=np.float64)) for m, k in syn._CFG5_MODELS[:6]] + [
    (cm.PINHOLE_RADIAL_TANGENTIAL, np.array([440.0, 1.01, 0.2, 480.0, 270.0, -0.1, 0.02, -0.003, 1e-3, -2e-3]))]


def _oracle_project(model, k, p, fov=1):
    lib = oracle_lib()
    px = np.zeros(2)
    dp = ctypes.POINTER(ctypes.c_double)
    ok = lib.icco_project(ctypes.c_int(model), k.ctypes.data_as(dp), np.ascontiguousarray(p).ctypes.data_as(dp), px.ctypes.data_as(dp), ctypes.c_int(fov))
    return bool(ok), px


@pytest.mark.parametrize("model,k", CASES, ids=[cm.MODEL_NAMES[m] for m, _ in CASES])
def test_projection_agrees_with_numpy(model, k):
    rng = np.random.default_rng(model)
    pts = np.concatenate([rng.uniform(-0.3, 0.3, size=(200, 2)), rng.uniform(0.3, 1.0, size=(200, 1))], axis=1)
    uv, valid = cm.project(model, k, pts)
    for i, p in enumerate(pts):
        ok, px = _oracle_project(model, k, p)
        assert ok == bool(valid[i])
        assert np.allclose(px, uv[i], rtol=1e-13, atol=1e-10)


def test_fov_is_not_dispatched_by_the_reference():
    # ceres_calib_split_residuals.h:366-389 has no FOV branch: success stays false
    ok, _ = _oracle_project(cm.FOV, np.array([437.0, 1.0, 489.0, 271.0, 0.9]), np.array([0.1, 0.05, 0.6]), fov=0)
    assert not ok


def test_unified_models_reject_points_outside_domain():
    ok, _ = _oracle_project(cm.DOUBLE_SPHERE, np.array([342.4, 1.0, 0.0, 472.6, 273.9, -0.215, 0.513]), np.array([0.1, 0.0, -5.0]))
    assert not ok
    ok, _ = _oracle_project(cm.EXTENDED_UNIFIED, np.array([438.0, 1.0, 0.0, 489.5, 272.0, 0.5115, 1.062]), np.array([0.1, 0.0, -5.0]))
    assert not ok


def test_division_undistortion_roundtrip():
    # undistorting the projected pixel with the division model recovers the pinhole pixel (self-consistency of the formula)
    k = np.array([437.1, 1.0, 489.1, 270.9, -1.44e-6])
    p = np.array([0.12, -0.07, 0.5])
    ok, px = _oracle_project(cm.DIVISION_UNDISTORTION, k, p)
    assert ok
    xd, yd = px[0] - k[2], px[1] - k[3]
    rd2 = xd * xd + yd * yd
    xu, yu = xd / (1 + k[4] * rd2), yd / (1 + k[4] * rd2)
    assert np.allclose([xu, yu], [k[0] * p[0] / p[2], k[0] * k[1] * p[1] / p[2]], rtol=1e-12)
