"""Pins the oracle against the only known-answer vectors the reference's path offers: the uniform B-spline blending
matrices of basalt_spline/spline_common.h:67-133."""
import ctypes

import numpy as np

from oracle_api import oracle_lib

M6 = np.array([[1, -5, 10, -10, 5, -1], [26, -50, 20, 20, -20, 5], [66, 0, -60, 0, 30, -10], [26, 50, 20, -20, -20, 10], [1, 5, 10, 10, 5, -5],
               [0, 0, 0, 0, 0, 1]], dtype=np.float64) / 120.0
MC6 = np.array([[120, 0, 0, 0, 0, 0], [119, 5, -10, 10, -5, 1], [93, 55, -30, -10, 15, -4], [27, 55, 30, -10, -15, 6], [1, 5, 10, 10, 5, -4],
                [0, 0, 0, 0, 0, 1]], dtype=np.float64) / 120.0
M3 = np.array([[1, -2, 1], [1, 2, -2], [0, 0, 1]], dtype=np.float64) / 2.0
MC3 = np.array([[2, 0, 0], [1, 2, -1], [0, 0, 1]], dtype=np.float64) / 2.0
BASE6 = np.array([[1, 1, 1, 1, 1, 1], [0, 1, 2, 3, 4, 5], [0, 0, 2, 6, 12, 20], [0, 0, 0, 6, 24, 60], [0, 0, 0, 0, 24, 120], [0, 0, 0, 0, 0, 120]], dtype=np.float64)


def _blend(n, cumulative):
    out = np.zeros((n, n))
    lib = oracle_lib()
    lib.icco_blending_matrix.restype = None
    lib.icco_blending_matrix(ctypes.c_int(n), ctypes.c_int(cumulative), out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out


def test_blending_matrices_order6():
    assert np.allclose(_blend(6, 0), M6, rtol=0, atol=1e-15)
    assert np.allclose(_blend(6, 1), MC6, rtol=0, atol=1e-15)


def test_blending_matrices_order3():
    assert np.allclose(_blend(3, 0), M3, rtol=0, atol=1e-15)
    assert np.allclose(_blend(3, 1), MC3, rtol=0, atol=1e-15)


def test_base_coefficients():
    out = np.zeros((6, 6))
    lib = oracle_lib()
    lib.icco_base_coefficients.restype = None
    lib.icco_base_coefficients(ctypes.c_int(6), out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    assert np.array_equal(out, BASE6)


def test_partition_of_unity():
    # rows of the non-cumulative matrix sum (over knots) to the monomial e_0: sum_i c_i(u) == 1 for every u
    assert np.allclose(M6.sum(axis=0), [1, 0, 0, 0, 0, 0], atol=1e-15)
    assert np.allclose(M3.sum(axis=0), [1, 0, 0], atol=1e-15)
