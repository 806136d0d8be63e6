"""-m gpu: the FP32 mode of the CUDA path (pmx_fiber_desc.precision = PMX_F32) against the FP64 numpy oracle.

BASELINE.json north_star: "an FP32 mode is reported separately with <= 1e-5 tolerance" -- relative L2 error of the
output field against the reference's (FP64) result.  The step schedule is produced by the same device-side
step control from float data, so ncycle may differ by a step where a boundary is hit within float rounding;
it is compared with a tolerance of one step."""
import numpy as np
import pytest

import oracle.fiber_oracle as orc
import polmux_b200 as pmx
from common import base_fiber, make_tx, rel_l2

pytestmark = pytest.mark.gpu
TOL32 = 1e-5


@pytest.fixture(autouse=True, params=['scalar', 'vector'])
def disp_mode(request, monkeypatch):
    import importlib
    fmod = importlib.import_module('polmux_b200.fiber')
    monkeypatch.setattr(fmod, 'DISP_MODE', request.param)
    return request.param


def run32(nsymb, nt, fib, flag, nch=1, ftype='unique', seed=1000):
    gs = make_tx(nsymb, nt, nch, ftype=ftype)
    orc.fiber(gs, fib, flag, rng=np.random.Generator(np.random.PCG64(seed)))
    pmx.fiber(fib, flag, rng=np.random.Generator(np.random.PCG64(seed)), precision='f32')
    G = pmx.GSTATE
    return rel_l2(G.FIELDX, G.FIELDY, gs.FIELDX, gs.FIELDY), gs


@pytest.mark.parametrize('lg', [12, 14, 16, 18])
def test_fp32_linear_gvd(lg):
    err, gs = run32(1 << (lg - 4), 16, base_fiber(length=1e5), 'g---')
    assert err < TOL32
    assert pmx.FIBER_LAST['ncycle'] == 1


@pytest.mark.parametrize('lg,nplates', [(12, 10), (16, 200)])
def test_fp32_linear_pmd(lg, nplates):
    err, gs = run32(1 << (lg - 4), 16, base_fiber(dgd=0.5, nplates=nplates), 'gp--')
    assert err < TOL32


@pytest.mark.parametrize('lg,manakov', [(14, 'no'), (14, 'yes'), (16, 'yes'), (17, 'no')])
def test_fp32_nonlinear_pmd(lg, manakov):
    """C1-like: 100 km, 'gps-', 10 plates, CNLSE and Manakov (about 45 steps)."""
    fib = base_fiber(length=1e5, dgd=1.0, nplates=10, manakov=manakov)
    err, gs = run32(1 << (lg - 4), 16, fib, 'gps-')
    assert err < TOL32
    assert abs(pmx.FIBER_LAST['ncycle'] - gs.log['ncycle']) <= 1


def test_fp32_sepfields_spm():
    """two 'sepfields' columns, SPM only (no 'p'): per-column gamma."""
    fib = base_fiber(length=5e4)
    err, gs = run32(1 << 10, 16, fib, 'g-s-', nch=2, ftype='sepfields')
    assert err < TOL32


def test_fp32_ampliflat_gain_and_injected_noise():
    """ampliflat on an FP32 field: u*sqrt(G) + sigma*noise (ampliflat.m:78-148) against numpy, float rounding only."""
    from polmux_b200 import _lib
    ctx = _lib.default_context()
    n = 1 << 12
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((1, 1, n)) + 1j * rng.standard_normal((1, 1, n)))
    y = (rng.standard_normal((1, 1, n)) + 1j * rng.standard_normal((1, 1, n)))
    noise = (rng.standard_normal((1, 2, n)) + 1j * rng.standard_normal((1, 2, n)))
    f = _lib.DeviceField(ctx, n, 1, 1, precision=_lib.PMX_F32)
    f.upload(x, y)
    _lib.ampliflat_exec(ctx, f, 4.0, [0.25], noise=noise)
    gx, gy = f.download()
    np.testing.assert_allclose(gx[0, 0], 2.0 * x[0, 0] + 0.25 * noise[0, 0], rtol=0, atol=2e-6)
    np.testing.assert_allclose(gy[0, 0], 2.0 * y[0, 0] + 0.25 * noise[0, 1], rtol=0, atol=2e-6)


def test_fp32_qpsk_count_is_fp64_only():
    from polmux_b200 import _lib
    ctx = _lib.default_context()
    f = _lib.DeviceField(ctx, 1 << 12, 1, 1, precision=_lib.PMX_F32)
    with pytest.raises(_lib.PolmuxError):
        _lib.qpsk_count(ctx, f, np.zeros((2, 256), dtype=np.uint8), 256, 16, 0)
