"""The exact arithmetic of the device ray sampler (csrc/nfb_sampler.h), through its host-only C hook: np.cumsum of the
reference's importance map — a sequential float64 accumulation over H*W entries — reproduced bit for bit from the map's run
structure, with and without zeroed entries (the later rounds of np.random.choice(replace=False)).  No GPU."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def lib(built_lib):
    import nerf  # noqa: F401
    from nerf import _capi
    return _capi


def _cdf(capi, m, zeroed, ks):
    z = np.ascontiguousarray(np.sort(np.asarray(zeroed, dtype=np.int64)))
    ks = np.ascontiguousarray(np.asarray(ks, dtype=np.int64))
    out = np.empty(ks.size, dtype=np.float64)
    rc = capi.lib.nfb_host_map_cdf(C.byref(m), z.ctypes.data_as(C.POINTER(C.c_longlong)) if z.size else None, int(z.size),
                                   ks.ctypes.data_as(C.POINTER(C.c_longlong)), int(ks.size), out.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0
    return out


@pytest.mark.parametrize("H,W,bbox", [(512, 512, (128, 384, 150, 400)), (64, 96, (0, 64, 0, 96)), (96, 64, (10, 11, 5, 60)),
                                       (128, 128, (0, 50, 30, 128)), (100, 100, (40, 40, 10, 20))])
def test_map_cumsum_bit_exact(lib, H, W, bbox):
    from nerf import ray_sampler
    m, flat = ray_sampler.importance_map(H, W, bbox, 0.9)
    rng = np.random.default_rng(H * 1000 + W)
    ks = np.arange(H * W)
    ref = np.cumsum(flat)
    got = _cdf(lib, m, [], ks)
    assert np.array_equal(got, ref)  # every one of the H*W partial sums, bit for bit
    assert _cdf(lib, m, [], [-1])[0] == ref[-1]
    for n_zero in (1, 37, 2030):
        z = rng.choice(H * W, size=n_zero, replace=False)
        p = flat.copy()
        p[z] = 0
        assert np.array_equal(_cdf(lib, m, z, ks), np.cumsum(p))


def test_map_values_match_the_reference_expression(lib):
    from nerf import ray_sampler
    m, flat = ray_sampler.importance_map(64, 64, (10, 30, 20, 50), 0.9)
    assert flat[10 * 64 + 20] == m.q_in and flat[0] == m.q_out and abs(flat.sum() - 1.0) < 1e-12
    assert sorted(set(flat.tolist())) == sorted({m.q_in, m.q_out})
