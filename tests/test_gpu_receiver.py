"""The receiver front-end on the device (polmux_b200/receiver.py: filter plans, pmx_field_modulate, pmx_cohmix_exec)
against the interpreted receiver_cohmix.m (goldens of tests/golden/rx/) and, at sizes that take the three-pass path,
against the oracle; the filter plan alone against numpy."""
import numpy as np
import pytest

import oracle.receiver_oracle as rxo
import polmux_b200 as pmx
from polmux_b200 import _lib, receiver, synth
from test_receiver_oracle import CASES, load, oracle_state

pytestmark = pytest.mark.gpu


def product_state(z, m):
    pmx.reset_all(m['nsymb'], m['nt'], m['nch'])
    G = pmx.GSTATE
    G.SYMBOLRATE, G.POWER, G.LAMBDA = m['rate'], z['POWER'].ravel().copy(), synth.wdm_lambdas(m['nch'])
    G.FIELDX = z['FIELDX'].copy()
    G.FIELDY = z['FIELDY'].copy() if z['FIELDY'].size else None
    G.FIELDX_TX = z['FIELDX_TX'].copy()
    G.FIELDY_TX = z['FIELDY_TX'].copy() if z['FIELDY_TX'].size else None
    return G


@pytest.mark.parametrize('name', CASES)
def test_cuda_receiver_matches_reference_source(name):
    z, m, x = load(name)
    G = product_state(z, m)
    before = (np.array(G.FIELDX), None if G.FIELDY is None else np.array(G.FIELDY))
    iric, xo = receiver.receiver_cohmix(m['ich'], x, nargout=2)
    assert iric.shape == z['Iric'].shape
    assert np.linalg.norm(iric - z['Iric']) <= 1e-10 * np.linalg.norm(z['Iric'])
    assert abs(xo['avgebx'] / float(z['avgebx'][0]) - 1) < 1e-10
    if z['avgeby'].size:
        assert abs(xo['avgeby'] / float(z['avgeby'][0]) - 1) < 1e-10
    assert abs(xo['post_delay'] - float(z['post_delay'][0])) <= 1e-12 * max(1.0, abs(float(z['post_delay'][0])))
    # GSTATE is left unchanged (receiver_cohmix.m:60-61)
    assert np.array_equal(np.asarray(G.FIELDX), before[0])
    if before[1] is not None:
        assert np.array_equal(np.asarray(G.FIELDY), before[1])
    one = receiver.receiver_cohmix(m['ich'], x)
    assert np.array_equal(one, iric)


@pytest.mark.parametrize('lg,nch,ftype,ich', [(14, 3, 'unique', 1), (16, 1, 'unique', 1), (13, 2, 'sepfields', 2)])
def test_cuda_receiver_three_pass_sizes_against_oracle(lg, nch, ftype, ich):
    nt = 32
    nsymb = (1 << lg) // nt
    ex, ey, _, _ = synth.pdm_qpsk(nsymb, nt, nch)
    pmx.reset_all(nsymb, nt, nch)
    G = pmx.GSTATE
    G.SYMBOLRATE, G.LAMBDA, G.POWER = 28.0, synth.wdm_lambdas(nch, 1550.0, 0.8), np.full(nch, 2.0)
    pmx.create_field(ftype, ex, ey, {'power': 'average'})
    pmx.fiber(dict(synth.SMF, length=2e4, dgd=0.2, nplates=10, manakov='yes'), 'gps-',
              rng=np.random.Generator(np.random.PCG64(2)))
    pn = np.cumsum(0.01 * np.random.Generator(np.random.PCG64(3)).standard_normal(1 << lg))
    x = {'oftype': 'gauss', 'obw': 1.9, 'eftype': 'bessel5', 'ebw': 0.65, 'lopower': 1.0, 'lodetuning': 1.3e9,
         'lophasenoise': pn, 'dpost': -340.0, 'slopez': 0.0, 'lambda': 1550.0}

    class GS:
        pass
    gs = GS()
    for k in ('FN', 'LAMBDA', 'SYMBOLRATE', 'NSYMB', 'NT', 'NCH', 'POWER'):
        setattr(gs, k, getattr(G, k))
    gs.FIELDX, gs.FIELDY = np.array(G.FIELDX), np.array(G.FIELDY)
    ref, xr = rxo.receiver_cohmix(gs, ich, x)
    iric, xo = receiver.receiver_cohmix(ich, x, nargout=2)
    assert np.linalg.norm(iric - ref) <= 1e-10 * np.linalg.norm(ref)
    assert abs(xo['avgebx'] / xr['avgebx'] - 1) < 1e-10 and abs(xo['avgeby'] / xr['avgeby'] - 1) < 1e-10


@pytest.mark.parametrize('lg,nfc,batch,percol,prec', [(8, 1, 1, False, 'f64'), (12, 3, 2, True, 'f64'), (13, 2, 2, True, 'f64'),
                                                      (16, 1, 3, False, 'f64'), (20, 1, 1, False, 'f64'),
                                                      (11, 2, 1, True, 'f32'), (15, 2, 2, False, 'f32')])
def test_filter_plan_against_numpy(lg, nfc, batch, percol, prec):
    """u <- ifft(fft(u) .* H) for an arbitrary complex H: on-chip kernel (<= 2^12) and three-pass path, shared and
    per-column H, batches, both precisions (FP32: <= 1e-5, reported apart from the FP64 bar)"""
    n = 1 << lg
    g = np.random.Generator(np.random.PCG64(lg))
    ux = g.standard_normal((batch, nfc, n)) + 1j * g.standard_normal((batch, nfc, n))
    uy = g.standard_normal((batch, nfc, n)) + 1j * g.standard_normal((batch, nfc, n))
    h = (g.standard_normal((n, nfc if percol else 1)) + 1j * g.standard_normal((n, nfc if percol else 1)))
    ctx = _lib.default_context()
    pc = _lib.PMX_F64 if prec == 'f64' else _lib.PMX_F32
    fld = _lib.DeviceField(ctx, n, nfc, batch, precision=pc)
    fld.upload(ux, uy)
    flt = _lib.Filter(ctx, n, nfc, h, batch=batch, precision=pc)
    flt.execute(fld)
    flt.close()
    ox, oy = fld.download()
    fld.close()
    hh = h.T[None] if percol else h.T[None]            # [1][nfc or 1][n]
    rx_ = np.fft.ifft(np.fft.fft(ux, axis=2) * hh, axis=2)
    ry_ = np.fft.ifft(np.fft.fft(uy, axis=2) * hh, axis=2)
    err = np.sqrt((np.abs(ox - rx_) ** 2).sum() + (np.abs(oy - ry_) ** 2).sum()) / np.sqrt((np.abs(rx_) ** 2).sum() + (np.abs(ry_) ** 2).sum())
    assert err < (1e-12 if prec == 'f64' else 1e-5), err


def test_modulate_and_copy_cols():
    n, nfc = 1 << 13, 3
    g = np.random.Generator(np.random.PCG64(1))
    ux = g.standard_normal((1, nfc, n)) + 1j * g.standard_normal((1, nfc, n))
    uy = g.standard_normal((1, nfc, n)) + 1j * g.standard_normal((1, nfc, n))
    ctx = _lib.default_context()
    fld = _lib.DeviceField(ctx, n, nfc, 1)
    fld.upload(ux, uy)
    one = _lib.DeviceField(ctx, n, 1, 1)
    _lib.field_copy_cols(one, 0, fld, 2, 1)
    for m in (5, -37, n + 3):
        _lib.field_modulate(ctx, one, m)
    ox, oy = one.download()
    ph = np.exp(2j * np.pi * (((5 - 37 + n + 3) * np.arange(n)) % n) / n)
    np.testing.assert_allclose(ox[0, 0], ux[0, 2] * ph, atol=1e-13)
    np.testing.assert_allclose(oy[0, 0], uy[0, 2] * ph, atol=1e-13)
    with pytest.raises(_lib.PolmuxError):
        _lib.field_copy_cols(one, 0, fld, 3, 1)
    fld.close()
    one.close()
