"""Page-locked caller buffers: icc_set_frames / icc_set_imu send the corner and reading arrays straight to the device by DMA (no host staging
copy) and use them in place.  The results must be the ones of the staged path bit for bit -- also on the paths that have to fetch the arrays
back to gather them (unsorted IMU stream, a sample dropped mid-stream, a non-contiguous frame selection) and in a residual shard."""
import numpy as np
import pytest
import torch

from helpers import F_STAGE1
from openimucameracalibrator_b200 import _capi as capi
from openimucameracalibrator_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
BIG = ("uv", "point_ids", "accel", "gyro", "imu_t")


def pinned(ds):
    out, keep = dict(ds), []
    for k in BIG:
        t = torch.from_numpy(np.ascontiguousarray(ds[k])).pin_memory(); keep.append(t)
        out[k] = t.numpy()
    out["_keep"] = keep
    return out


def variants():
    base = dict(syn.make_dataset(syn.tiny_config(n_frames=30)))
    yield "plain", base, None
    d = dict(base); perm = np.random.default_rng(0).permutation(len(d["imu_t"]))
    d["imu_t"], d["accel"], d["gyro"] = d["imu_t"][perm], d["accel"][perm], d["gyro"][perm]
    yield "shuffled_imu", d, None
    d = dict(base); t = d["imu_t"].copy(); t[50] = 1e6; d["imu_t"] = t
    yield "imu_hole", d, None
    d = dict(base); t = d["frame_t"].copy(); t[7] = t[0] - 10.0; d["frame_t"] = t     # a view before the others: the kept frames are not one contiguous run of corners? (time order differs)
    yield "frame_out_of_order", d, None
    yield "shard_1_of_3", base, (1, 3)


@pytest.mark.parametrize("name,ds,shard", list(variants()), ids=lambda v: v if isinstance(v, str) else "")
def test_pinned_inputs_match_staged_inputs(gpu_factory, name, ds, shard):
    a = gpu_factory(); capi.load_dataset(a, ds, shard=shard)
    dp = pinned(ds)
    b = gpu_factory(); capi.load_dataset(b, dp, shard=shard)
    assert a.num_residuals() == b.num_residuals() and a.num_knots() == b.num_knots()
    for x, y in zip(a.imu_used(), b.imu_used()):
        assert np.array_equal(x, y)
    assert np.array_equal(a.get_gravity(), b.get_gravity())
    ca, ra, ga, _ = a.evaluate(F_STAGE1)
    cb, rb, gb, _ = b.evaluate(F_STAGE1)
    assert np.array_equal(ra, rb)
    assert abs(ca - cb) <= 1e-13 * ca                                          # (atomics: summation order differs between runs)
    assert np.allclose(ga, gb, rtol=1e-12, atol=1e-9 * np.abs(ga).max())
    if shard is None:
        sa, sb = a.optimize(10, F_STAGE1), b.optimize(10, F_STAGE1)
        assert sa.iterations == sb.iterations and abs(sa.final_cost - sb.final_cost) <= 1e-9 * sa.final_cost


def test_pinned_buffers_are_not_referenced_after_set(gpu_factory):
    """The DMA has finished when set_frames / set_imu return: overwriting the caller's arrays afterwards must not change the problem."""
    ds = dict(syn.make_dataset(syn.tiny_config(n_frames=20)))
    a = gpu_factory(); capi.load_dataset(a, ds)
    dp = pinned(ds)
    b = gpu_factory()
    W, H = dp["image_size"]
    b.set_camera(dp["model"], dp["intrinsics"], W, H); b.set_board_points(dp["board_xyzw"])
    b.set_frames(dp["frame_t"], dp["corner_offsets"], dp["point_ids"], dp["uv"], dp["q_wc"], dp["p_wc"])
    b.set_imu(dp["imu_t"], dp["accel"], dp["gyro"])
    for k in ("uv", "accel", "gyro"):
        dp[k][...] = 0.0
    b.batch_init_spline(dp["T_i_c_init"], dp["dt_so3_s"], dp["dt_r3_s"], dp["std_so3"], dp["std_r3"], dp["time_offset_imu_to_cam_s"], dp["init_line_delay_s"],
                        acc_bias=dp["acc_bias"], gyr_bias=dp["gyr_bias"], dispatch_fov=False)
    b.set_known_gravity_dir(dp["gravity"])
    ca = a.evaluate(F_STAGE1, residuals=False, gradient=False)[0]; cb = b.evaluate(F_STAGE1, residuals=False, gradient=False)[0]
    assert abs(ca - cb) <= 1e-13 * ca
