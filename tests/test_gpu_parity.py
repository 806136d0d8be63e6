"""-m gpu: the CUDA path (through fiber() -> C ABI pmx_fiber_run) against the numpy oracle.

Tolerance (BASELINE.json north_star): relative L2 error of the output field <= 1e-10 in FP64;
ncycle and the per-step trunk schedule must be equal."""
import numpy as np
import pytest

import oracle.fiber_oracle as orc
import polmux_b200 as pmx
from common import base_fiber, make_tx, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(autouse=True, params=['scalar', 'vector'])
def disp_mode(request, monkeypatch):
    """every parity test runs in both dispersion modes of the C ABI"""
    import importlib
    fmod = importlib.import_module('polmux_b200.fiber')
    monkeypatch.setattr(fmod, 'DISP_MODE', request.param)
    return request.param


def run_both(nsymb, nt, fib, flag, nch=1, ftype='unique', seed=1000, pavg=2.0, rate=28.0):
    gs = make_tx(nsymb, nt, nch, rate=rate, pavg_mw=pavg, ftype=ftype)
    brf_o = orc.fiber(gs, fib, flag, rng=np.random.Generator(np.random.PCG64(seed)))
    brf_g = pmx.fiber(fib, flag, rng=np.random.Generator(np.random.PCG64(seed)), trace=True)
    G = pmx.GSTATE
    err = rel_l2(G.FIELDX, G.FIELDY, gs.FIELDX, gs.FIELDY)
    return err, gs, brf_o, brf_g


def check_schedule(gs):
    L = pmx.FIBER_LAST
    assert L['ncycle'] == gs.log['ncycle']
    sched = gs.log['schedule']
    assert list(L['trace_ntrunk']) == [s['ntrunk'] for s in sched]
    np.testing.assert_allclose(L["trace_dz"], [s["dz"] for s in sched], rtol=1e-10)  # dz = -log(1-dl)/alpha is ill-conditioned near the dzmax cap
    np.testing.assert_allclose(L['firstdz'], gs.log['firstdz'], rtol=1e-13)


@pytest.mark.parametrize('lg', [12, 13, 14, 16])
def test_linear_gvd(lg):
    """'g---': one step, identity Jones."""
    err, gs, _, _ = run_both(1 << (lg - 4), 16, base_fiber(length=1e5), 'g---')
    assert err < TOL
    assert pmx.FIBER_LAST['ncycle'] == 1


@pytest.mark.parametrize('lg,nplates', [(12, 10), (15, 20), (16, 200)])
def test_linear_pmd(lg, nplates):
    """'gp--': one step, nplates trunks."""
    err, gs, bo, bg = run_both(1 << (lg - 4), 16, base_fiber(dgd=0.5, nplates=nplates), 'gp--')
    assert err < TOL
    np.testing.assert_array_equal(bo['theta'], bg['theta'])
    check_schedule(gs)
    assert pmx.FIBER_LAST['ntot'] == nplates


@pytest.mark.parametrize('manakov', ['yes', 'no'])
@pytest.mark.parametrize('lg', [12, 14, 16])
def test_nonlinear_pmd(lg, manakov):
    """'gps-' Manakov / CNLSE with random plates: the C1/C2 code path at small size."""
    fib = base_fiber(length=1e5, dgd=1.0, nplates=10, manakov=manakov)
    err, gs, _, _ = run_both(1 << (lg - 4), 16, fib, 'gps-')
    assert err < TOL
    check_schedule(gs)


def test_nonlinear_100_plates():
    fib = base_fiber(length=8e4, dgd=0.1, nplates=100, manakov='yes')
    err, gs, _, _ = run_both(1 << 11, 16, fib, 'gps-')
    assert err < TOL
    check_schedule(gs)
    assert pmx.FIBER_LAST['ntot'] == 100


def test_two_pol_no_pmd():
    """two polarizations without 'p': identity plates, CNLSE nonlinear step ('g-s-')."""
    err, gs, _, _ = run_both(1 << 10, 16, base_fiber(), 'g-s-')
    assert err < TOL
    check_schedule(gs)


def test_sepfields_two_columns():
    """'sepfields' with nfc = 3 columns: per-column beta1/beta2/gamma, common step."""
    fib = base_fiber(length=5e4, dgd=0.3, nplates=20, manakov='yes')
    err, gs, _, _ = run_both(1 << 9, 16, fib, 'gps-', nch=3, ftype='sepfields')
    assert err < TOL
    check_schedule(gs)


def test_energy_ratio():
    """every sub-step is unitary except exp(-alpha dz): sum|u|^2 out/in = exp(-alpha L)."""
    fib = base_fiber(length=8e4, dgd=0.1, nplates=100, manakov='yes')
    gs = make_tx(1 << 10, 16)
    G = pmx.GSTATE
    e_in = np.sum(np.abs(G.FIELDX) ** 2 + np.abs(G.FIELDY) ** 2)
    pmx.fiber(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(7)))
    e_out = np.sum(np.abs(G.FIELDX) ** 2 + np.abs(G.FIELDY) ** 2)
    alphalin = np.log(10) * 1e-4 * fib['alphadB']
    assert abs(e_out / e_in / np.exp(-alphalin * fib['length']) - 1) < 1e-12


def test_plate_index_error():
    """(Lf=8e4, nplates=59) runs past the last plate in the reference; we report it."""
    fib = base_fiber(length=8e4, dgd=0.1, nplates=59)
    make_tx(1 << 8, 16)
    with pytest.raises(pmx.PolmuxError) as e:
        pmx.fiber(fib, 'gp--', rng=np.random.Generator(np.random.PCG64(1)))
    assert e.value.code == -4
