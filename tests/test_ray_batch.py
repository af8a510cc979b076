"""rth_ray_batch (csrc/host/host_abi.cpp), the multi-threaded generator of the SURVEY 8d C4 ray batches, against its definition
scenes.ray_batch: origins, segments, t_max and tags bit-equal; unit directions equal up to the last bit of cos / sin (numpy's
float32 cos / sin are not glibc's)."""
import numpy as np


def test_ray_batch_twin(native_libs):
    from rustracer_b200 import host, scenes
    lo, hi = np.float32([-3, -2, -5]), np.float32([4, 6, 7])
    for any_hit in (False, True):
        a = scenes.ray_batch(70000, lo, hi, any_hit=any_hit, first=12345)
        b = host.ray_batch(70000, lo, hi, any_hit=any_hit, first=12345)
        assert np.array_equal(a[:, :4], b[:, :4])
        assert np.array_equal(a[:, 7].view(np.uint32), b[:, 7].view(np.uint32))
        if any_hit:
            assert np.array_equal(a[:, 4:7], b[:, 4:7])
        else:
            assert np.abs(a[:, 4:7] - b[:, 4:7]).max() <= 1.2e-7
            assert np.array_equal(a[:, 6], b[:, 6])
