"""CPU: structural self-checks of the numpy oracle -- closed forms the
reference's own structure guarantees, plus the extended-precision arbiter."""
import math

import numpy as np
import pytest

import oracle.fiber_oracle as orc
from common import base_fiber, make_tx


def _rng(seed=1000):
    return np.random.Generator(np.random.PCG64(seed))


def test_linear_gvd_closed_form():
    """'g---' with two polarizations: ONE step, u_out = ifft(fft(u) exp(-i betat L)) exp(-alpha L/2)."""
    gs = make_tx(256, 16)
    x0, y0 = gs.FIELDX.copy(), gs.FIELDY.copy()
    fib = base_fiber(length=1e5)
    brf = orc.fiber(gs, fib, 'g---')
    assert gs.log['ncycle'] == 1
    alphalin = math.log(10) * 1e-4 * fib['alphadB']
    h = np.exp(-1j * brf['betat'] * fib['length']) * math.exp(-alphalin * fib['length'] / 2)
    ex = np.fft.ifft(np.fft.fft(x0, axis=0) * h, axis=0)
    ey = np.fft.ifft(np.fft.fft(y0, axis=0) * h, axis=0)
    assert orc.rel_l2(gs.FIELDX, gs.FIELDY, ex, ey) < 1e-12


def test_linear_pmd_single_step_trunks():
    """'gp--': one FFT/IFFT pair around the ordered product of all plates (lcorr each)."""
    gs = make_tx(256, 16)
    fib = base_fiber(length=8e4, dgd=0.5, nplates=20)
    orc.fiber(gs, fib, 'gp--', rng=_rng())
    sched = gs.log['schedule']
    assert len(sched) == 1 and sched[0]['ntrunk'] == 20
    np.testing.assert_allclose(sched[0]['dzb'], [4000.0] * 20)


def test_spm_exact_single_step():
    """'--s-' single scalar field: exact solution in one step (fiber.m:170-174)."""
    gs = make_tx(128, 16)
    gs.FIELDY = None
    u0 = gs.FIELDX.copy()
    fib = base_fiber(length=5e4)
    orc.fiber(gs, fib, '--s-')
    alphalin = math.log(10) * 1e-4 * fib['alphadB']
    leff = (1 - math.exp(-alphalin * fib['length'])) / alphalin
    gam = gs.log['gam'][0]
    ex = u0 * np.exp(-1j * gam * np.abs(u0) ** 2 * leff) * math.exp(-alphalin * fib['length'] / 2)
    assert np.linalg.norm(gs.FIELDX - ex) / np.linalg.norm(ex) < 1e-12
    assert gs.log['ncycle'] == 1


@pytest.mark.parametrize('manakov', ['yes', 'no'])
def test_energy_ratio(manakov):
    """every sub-step but exp(-alpha dz/2) is unitary: sum|u|^2 out/in = exp(-alpha L) to ~1e-13."""
    gs = make_tx(256, 16)
    e_in = np.sum(np.abs(gs.FIELDX) ** 2 + np.abs(gs.FIELDY) ** 2)
    fib = base_fiber(length=8e4, dgd=0.3, nplates=20, manakov=manakov)
    orc.fiber(gs, fib, 'gps-', rng=_rng())
    e_out = np.sum(np.abs(gs.FIELDX) ** 2 + np.abs(gs.FIELDY) ** 2)
    alphalin = math.log(10) * 1e-4 * fib['alphadB']
    assert abs(e_out / e_in / math.exp(-alphalin * fib['length']) - 1) < 1e-12


def test_inverse_pmd_round_trip():
    """fiber(x,'gp--') followed by inverse_pmd({brf}) restores the Tx field up to attenuation."""
    gs = make_tx(256, 16)
    x0, y0 = gs.FIELDX.copy(), gs.FIELDY.copy()
    fib = base_fiber(length=8e4, dgd=0.5, nplates=20)
    brf = orc.fiber(gs, fib, 'gp--', rng=_rng(7))
    orc.inverse_pmd(gs, [brf])
    att = math.exp(-math.log(10) * 1e-4 * fib['alphadB'] * fib['length'] / 2)
    assert orc.rel_l2(gs.FIELDX, gs.FIELDY, x0 * att, y0 * att) < 1e-10


@pytest.mark.parametrize('nplates,length', [(10, 1e5), (20, 8e4), (100, 8e4), (200, 8e4)])
def test_trunk_counter_invariants(nplates, length):
    """after matrix_ssfm ntot == nplates and sum(dzb) == Lf (fiber.m:529,545)."""
    gs = make_tx(128, 16)
    fib = base_fiber(length=length, dgd=0.2, nplates=nplates, manakov='yes')
    orc.fiber(gs, fib, 'gps-', rng=_rng(3))
    sched = gs.log['schedule']
    last = sched[-1]
    assert last['ntot'] + last['ntrunk'] - last['nmem'] == nplates
    assert abs(sum(sum(s['dzb']) for s in sched) - length) < 1e-6
    assert abs(sum(s['dz'] for s in sched) - length) < 1e-6
    assert len(sched) == gs.log['ncycle']


def test_manakov_is_cnlse_without_s3_rotation():
    """Manakov = CNLSE path minus the s3 rotation with gamma -> 8/9 gamma (fiber.m:499-504,841-851)."""
    n = 512
    g = np.random.default_rng(5)
    ux = (g.standard_normal((n, 1)) + 1j * g.standard_normal((n, 1)))
    uy = (g.standard_normal((n, 1)) + 1j * g.standard_normal((n, 1)))
    gam, dz = np.array([1.3e-6]), 1000.0
    mx, my = orc.matrix_nl_step(True, 0.0, gam, dz, ux.copy(), uy.copy(), 1, 1, 0)
    p = np.abs(ux) ** 2 + np.abs(uy) ** 2
    e = np.exp(-1j * gam[0] * dz * p)
    assert np.allclose(mx, ux * e, rtol=1e-13, atol=1e-13) and np.allclose(my, uy * e, rtol=1e-13, atol=1e-13)
    cx, cy = orc.matrix_nl_step(False, 0.0, gam, dz, ux.copy(), uy.copy(), 1, 1, 0)
    # the s3 rotation is orthogonal: power is preserved sample by sample
    np.testing.assert_allclose(np.abs(cx) ** 2 + np.abs(cy) ** 2, p, rtol=1e-12)


def test_checkstep_worked_trace():
    """Lf=1000, nplates=4, dz = 300,300,300, last 100."""
    lcorr, ntot, dz_miss = 250.0, 0, 0.0
    want = [([250.0, 50.0], 200.0, 0, 2), ([200.0, 100.0], 150.0, 1, 2), ([150.0, 150.0], 100.0, 1, 2)]
    for zprop, (dzb_w, miss_w, nmem_w, ntr_w) in zip((300.0, 600.0, 900.0), want):
        dzb, dz_miss, nmem, ntrunk = orc.checkstep(zprop, 300.0, lcorr, dz_miss, ntot)
        assert (dzb, dz_miss, nmem, ntrunk) == (dzb_w, miss_w, nmem_w, ntr_w)
        ntot += ntrunk - nmem
    dzb, dz_miss, nmem, ntrunk = orc.checkstep(1000.0, 100.0, lcorr, dz_miss, ntot)
    assert (dzb, dz_miss, nmem, ntrunk, ntot) == ([100.0], 0.0, 1, 1, 4)


def test_nextstep_edge_cases():
    """Pmax = 0 or phimax = Inf -> dzmax, with and without attenuation."""
    z = np.zeros((8, 1), dtype=complex)
    u = np.ones((8, 1), dtype=complex)
    for alpha in (0.0, 4.6e-5):
        assert orc.nextstep(2e4, 5e-3, np.array([1e-6]), alpha, z, z, True) == 2e4
        assert orc.nextstep(2e4, math.inf, np.array([1e-6]), alpha, u, u, True) == 2e4
    # small power: dl >= 1 branch
    assert orc.nextstep(2e4, 5e-3, np.array([1e-6]), 4.6e-5, u * 1e-3, u * 1e-3, True) == 2e4
    # the log branch
    dz = orc.nextstep(2e4, 5e-3, np.array([1.3e-6]), 4.6e-5, u, u, True)
    leff = 5e-3 / (1.3e-6 * 2)
    assert abs(dz - (-1 / 4.6e-5 * math.log(1 - 4.6e-5 * leff))) < 1e-9


def test_flag_table():
    """SURVEY A.2."""
    x = {'length': 1e5, 'dzmax': 2e4, 'dphimax': 5e-3}
    assert orc.parse_flag('gp--', 1, x) == ([1, 1, 0, 0], math.inf, 1e5)
    assert orc.parse_flag('--s-', 1, x) == ([0, 0, 1, 0], math.inf, 1e5)
    assert orc.parse_flag('--s-', 3, x) == ([0, 0, 1, 0], 5e-3, 2e4)
    assert orc.parse_flag('GPSX', 1, x) == ([1, 1, 1, 0], 5e-3, 2e4)
    assert orc.parse_flag('gpsx', 2, x) == ([1, 1, 1, 1], 5e-3, 2e4)
    with pytest.raises(ValueError):
        orc.parse_flag('g--x', 1, x)
    with pytest.raises(ValueError):
        orc.parse_flag('abcd', 1, x)


def test_longdouble_arbiter_agrees():
    """The same restatement in np.longdouble: the float64 oracle sits within ~1e-13 of it."""
    fib = base_fiber(length=5e4, dgd=0.3, nplates=10, manakov='no')
    gs = make_tx(128, 16)
    orc.fiber(gs, fib, 'gps-', rng=_rng(11))
    gl = make_tx(128, 16, real=np.longdouble)
    orc.fiber(gl, fib, 'gps-', rng=_rng(11))
    assert gs.log['ncycle'] == gl.log['ncycle']
    assert orc.rel_l2(gs.FIELDX, gs.FIELDY, gl.FIELDX, gl.FIELDY) < 1e-12


def test_xpm_vector_raises():
    gs = make_tx(64, 16, nch=2, ftype='sepfields')
    with pytest.raises(ValueError):
        orc.fiber(gs, base_fiber(dgd=0.1, nplates=10), 'gpsx', rng=_rng())


def test_scalar_path_matches_vector_path_with_zero_y():
    """single polarization 'g-s-' (scalar_ssfm) == two-pol CNLSE path fed with FIELDY = 0."""
    gs = make_tx(128, 16)
    gs.FIELDY = None
    fib = base_fiber(length=5e4)
    gv = make_tx(128, 16)
    gv.FIELDY = np.zeros_like(gv.FIELDX)
    orc.fiber(gs, fib, 'g-s-')
    orc.fiber(gv, fib, 'g-s-')
    assert gs.log['ncycle'] == gv.log['ncycle']
    assert np.linalg.norm(gs.FIELDX - gv.FIELDX) / np.linalg.norm(gv.FIELDX) < 1e-12
    assert np.max(np.abs(gv.FIELDY)) == 0.0
