"""Test-side loader of the CPU oracle (oracle/libicc_oracle.so).  Only tests/, smoke() and bench.py's CPU legs use this."""
import ctypes
import os
import subprocess

from openimucameracalibrator_b200._capi import CApi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libicc_oracle.so")


def build_oracle(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("icc_oracle.cpp", "oracle_math.hpp")]
    if force or not os.path.exists(ORACLE_SO) or any(os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in srcs if os.path.exists(s)):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return ORACLE_SO


_lib = None


def oracle_lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_oracle())
    return _lib


def new_oracle(n_threads=0) -> CApi:
    return CApi(oracle_lib(), "icco_", n_threads)
