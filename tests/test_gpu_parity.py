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


@pytest.mark.parametrize('nch,flag,man', [(24, 'gps-', 'yes'), (40, 'gps-', 'no'), (64, 'gp--', 'no')])
def test_sepfields_many_columns(nch, flag, man):
    """more separate channels than a dense WDM comb of the reference's scripts (up to the library's 64 columns): per-column
    constants, the step control's max over all the columns"""
    fib = base_fiber(length=2e4, dgd=0.3, nplates=10, manakov=man, slope=0.057)
    err, gs, _, _ = run_both(1 << 8, 16, fib, flag, nch=nch, ftype='sepfields', pavg=0.5)
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


def test_batch_of_realizations_matches_oracle_each(disp_mode):
    """A resident batch (5 realizations, own plate draw each; run as two realization groups on two streams,
    ragged step counts) against one oracle run per realization: field <= 1e-10, ncycle equal."""
    from polmux_b200 import _lib, mc
    from polmux_b200.fiber import fiber_setup, setup_to_desc
    nsymb, nt, batch = 1 << 10, 16, 5
    n = nsymb * nt
    fib = base_fiber(length=6e4, dgd=0.4, nplates=12, manakov='yes')
    gs0 = make_tx(nsymb, nt)
    G = pmx.GSTATE
    setup = fiber_setup(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(0)))
    draws = [mc.draw_plates(2000 + b, setup.nplates) for b in range(batch)]
    pl = [np.stack([d[i] for d in draws]) for i in range(3)]
    ctx = _lib.default_context()
    desc, keep = setup_to_desc(setup, batch=batch, plate_sets=batch, db0=pl[0], theta=pl[1], epsilon=pl[2])
    plan = _lib.Plan(ctx, desc, keep)
    tx = _lib.DeviceField(ctx, n, 1, 1)
    # realization b gets the Tx field scaled by (1 + b/10): different peak power -> different step schedule
    work = _lib.DeviceField(ctx, n, 1, batch)
    xs = np.stack([(1 + b / 10) * np.asarray(G.FIELDX).T for b in range(batch)])
    ys = np.stack([(1 + b / 10) * np.asarray(G.FIELDY).T for b in range(batch)])
    work.upload(xs, ys)
    res = plan.execute(work)
    gx, gy = work.download()
    ncycles = []
    for b in range(batch):
        gs = make_tx(nsymb, nt)
        gs.FIELDX = gs.FIELDX * (1 + b / 10)
        gs.FIELDY = gs.FIELDY * (1 + b / 10)
        f = dict(fib)
        f.update(db0=draws[b][0], theta=draws[b][1], epsilon=draws[b][2])
        # user-given plates go through the PMF branch: dgdrms = dgd/nplates there, sqrt(3*pi/8)*dgd/sqrt(nplates) for
        # random plates (fiber.m:269 vs :277): rescale the DGD so that both sides see the same dgdrms
        f['dgd'] = np.sqrt(3 * np.pi / 8) * fib['dgd'] * np.sqrt(setup.nplates)
        orc.fiber(gs, f, 'gps-')
        err = rel_l2(gx[b].T, gy[b].T, gs.FIELDX, gs.FIELDY)
        assert err < TOL, (b, err)
        assert int(res.ncycle[b]) == gs.log['ncycle']
        ncycles.append(gs.log['ncycle'])
    assert len(set(ncycles)) > 1   # the batch really is ragged
    tx.close() if hasattr(tx, 'close') else None


def test_inverse_pmd_on_device(disp_mode):
    """inverse_pmd.m: fiber('gp--') followed by inverse_pmd(brf) gives the transmitted field back (self-check
    4), and equals the oracle's inverse_pmd applied to the same propagated field."""
    gs = make_tx(1 << 10, 16)
    tx_x, tx_y = np.array(pmx.GSTATE.FIELDX), np.array(pmx.GSTATE.FIELDY)
    fib = base_fiber(length=5e4, dgd=0.6, nplates=15, alphadB=0.0)
    brf_o = orc.fiber(gs, fib, 'gp--', rng=np.random.Generator(np.random.PCG64(11)))
    brf_g = pmx.fiber(fib, 'gp--', rng=np.random.Generator(np.random.PCG64(11)))
    orc.inverse_pmd(gs, [brf_o])
    pmx.inverse_pmd(brf_g)
    G = pmx.GSTATE
    assert rel_l2(G.FIELDX, G.FIELDY, gs.FIELDX, gs.FIELDY) < TOL
    assert rel_l2(G.FIELDX, G.FIELDY, tx_x, tx_y) < 1e-9
    # two fibers in a row, inverted together
    gs = make_tx(1 << 10, 16)
    b1 = pmx.fiber(fib, 'gp--', rng=np.random.Generator(np.random.PCG64(12)))
    b2 = pmx.fiber(fib, 'gp--', rng=np.random.Generator(np.random.PCG64(13)))
    pmx.inverse_pmd([b1, b2])
    assert rel_l2(pmx.GSTATE.FIELDX, pmx.GSTATE.FIELDY, tx_x, tx_y) < 1e-9


def test_resident_chain_equals_per_call_chain(disp_mode):
    """fiber -> ampliflat -> fiber with the field left in HBM between the calls (gstate.RESIDENT, the default) gives
    the bits of the same chain with a download and an upload at every call; arrays the caller still holds keep their
    values (the interpreter's value semantics: x0 = GSTATE.FIELDX; fiber(...) leaves x0 alone)"""
    from polmux_b200 import gstate
    fib = base_fiber(length=2e4, dgd=0.3, nplates=12, manakov='yes')
    outs = []
    for resident in (True, False):
        old = gstate.RESIDENT
        gstate.RESIDENT = resident
        try:
            gs = make_tx(1 << 10, 16)
            G = pmx.GSTATE
            hx, hy = G.FIELDX, G.FIELDY
            x0, y0 = hx.copy(), hy.copy()
            noise = np.random.Generator(np.random.PCG64(5)).standard_normal((1 << 14, 4)).view(np.complex128).copy()
            pmx.fiber(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(21)))
            assert G.is_resident() == resident
            pmx.ampliflat(4.0, 'gain', {'f': 5.0, 'noise': noise})
            assert G.is_resident() == resident
            pmx.fiber(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(22)))
            assert G.field_shape() == (1 << 14, 1) and G.has_y()
            assert G.is_resident() == resident
            rx, ry = G.FIELDX, G.FIELDY
            assert not G.is_resident() and rx is not hx and np.array_equal(hx, x0) and np.array_equal(hy, y0)
            hx, hy = rx, ry
            outs.append((np.array(hx), np.array(hy)))
            # the oracle on the same chain
            if resident:
                orc.fiber(gs, fib, 'gps-', rng=np.random.Generator(np.random.PCG64(21)))
                orc.ampliflat(gs, 4.0, 5.0, noise=noise)
                orc.fiber(gs, fib, 'gps-', rng=np.random.Generator(np.random.PCG64(22)))
                assert rel_l2(hx, hy, gs.FIELDX, gs.FIELDY) < TOL
        finally:
            gstate.RESIDENT = old
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_link_call_equals_the_script_loop(disp_mode):
    """pmx_link_run (one call: H2D, nspan x [fiber ; ampliflat] on the device, D2H) gives the bits of the loop of
    fiber() and ampliflat() calls the reference's scripts write (ex06_ber.m:110-115), with injected and with
    device-generated ASE; the oracle agrees on the injected-noise link"""
    fib = base_fiber(length=2e4, dgd=0.3, nplates=12, manakov='yes')
    nspan, n = 3, 1 << 14
    g = np.random.Generator(np.random.PCG64(9))
    noise = [g.standard_normal((n, 4)).view(np.complex128).copy() for _ in range(nspan)]
    for opts_of in (lambda k: {'f': 5.0, 'noise': noise[k]}, lambda k: {'f': 5.0}):
        gs = make_tx(1 << 10, 16)
        G = pmx.GSTATE
        r = np.random.Generator(np.random.PCG64(77))
        brfs, ncyc = [], []
        for k in range(nspan):
            brfs.append(pmx.fiber(fib, 'gps-', rng=r))
            ncyc.append(pmx.FIBER_LAST['ncycle'])
            pmx.ampliflat(4.0, 'gain', opts_of(k), seed=300 + k)
        loop = (np.array(G.FIELDX), np.array(G.FIELDY), np.array(G.DELAY), np.array(G.DISP))
        make_tx(1 << 10, 16)
        o = opts_of(0)
        if 'noise' in o:
            o['noise'] = noise
        out = pmx.link(fib, 'gps-', nspan, 4.0, o, rng=np.random.Generator(np.random.PCG64(77)), seed=300)
        assert np.array_equal(G.FIELDX, loop[0]) and np.array_equal(G.FIELDY, loop[1])
        assert np.array_equal(G.DELAY, loop[2]) and np.array_equal(G.DISP, loop[3])
        assert pmx.FIBER_LAST['ncycle_per_span'] == ncyc and len(out) == nspan
        for a, b in zip(out, brfs):
            assert np.array_equal(a['theta'], b['theta']) and np.array_equal(a['db0'], b['db0']) and a['lcorr'] == b['lcorr']
        if 'noise' in o:
            ro = np.random.Generator(np.random.PCG64(77))
            for k in range(nspan):
                orc.fiber(gs, fib, 'gps-', rng=ro)
                orc.ampliflat(gs, 4.0, 5.0, noise=noise[k])
            assert rel_l2(G.FIELDX, G.FIELDY, gs.FIELDX, gs.FIELDY) < TOL


@pytest.mark.parametrize('nch,lg,delay', [(3, 14, True), (9, 16, False), (1, 13, True)])
def test_create_field_unique_on_device(nch, lg, delay):
    """create_field('unique') with the multiplex on the device (field.DEVICE_MUX: spectrum shift = modulation in time,
    one pointwise pass) against the reference's fft / fastshift / ifft arithmetic on the host (create_field.m:180-199),
    with options.power = 'average' and integer-sample delays; the field it leaves in HBM feeds fiber() directly"""
    from polmux_b200 import field as fmod, synth
    nt = 64
    nsymb = (1 << lg) // nt
    ex, ey, _, _ = synth.pdm_qpsk(nsymb, nt, nch)
    lams = synth.wdm_lambdas(nch, 1550.0, 0.4)
    opts = {'power': 'average'}
    if delay:
        opts['delay'] = np.random.Generator(np.random.PCG64(3)).random((2, nch))
    res = {}
    for dev in (False, True):
        pmx.reset_all(nsymb, nt, nch)
        G = pmx.GSTATE
        G.SYMBOLRATE, G.LAMBDA, G.POWER = 28.0, lams.copy(), np.full(nch, 1.0)
        old = fmod.DEVICE_MUX
        fmod.DEVICE_MUX = dev
        try:
            pmx.create_field('unique', ex, ey, dict(opts))
        finally:
            fmod.DEVICE_MUX = old
        assert G.is_resident() == dev
        tx = (np.array(G.FIELDX_TX), np.array(G.FIELDY_TX))
        pw, dl = np.array(G.POWER), np.array(G.DELAY)
        fib = base_fiber(length=2e4, dgd=0.2, nplates=10, manakov='yes')
        pmx.fiber(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(8)))
        res[dev] = (tx, pw, dl, np.array(G.FIELDX), np.array(G.FIELDY), pmx.FIBER_LAST['ncycle'])
    h, d = res[False], res[True]
    assert rel_l2(d[0][0], d[0][1], h[0][0], h[0][1]) < 1e-13
    assert np.array_equal(h[1], d[1]) and np.array_equal(h[2], d[2])
    assert rel_l2(d[3], d[4], h[3], h[4]) < 1e-11 and h[5] == d[5]


@pytest.mark.parametrize('pol', ['asex', 'asey'])
def test_ampliflat_onepol(pol):
    """options.onepol (ampliflat.m:107-118): ASE on one polarization only, the other one just amplified"""
    gs = make_tx(1 << 9, 16)
    G = pmx.GSTATE
    noise = np.random.Generator(np.random.PCG64(4)).standard_normal((1 << 13, 4)).view(np.complex128).copy()
    orc.ampliflat(gs, 7.0, 5.0, noise=noise, onepol=pol)
    pmx.ampliflat(7.0, 'gain', {'f': 5.0, 'noise': noise, 'onepol': pol})
    assert rel_l2(G.FIELDX, G.FIELDY, gs.FIELDX, gs.FIELDY) < 1e-14
    quiet = G.FIELDY if pol == 'asex' else G.FIELDX
    tx = G.FIELDY_TX if pol == 'asex' else G.FIELDX_TX
    assert np.allclose(quiet, tx * np.sqrt(10 ** 0.7), rtol=1e-14, atol=0)
    with pytest.raises(ValueError):
        pmx.ampliflat(7.0, 'gain', {'f': 5.0, 'onepol': 'both'})


def test_ampliflat_default_calls_draw_fresh_noise():
    """ampliflat.m:132-135 draws new randn samples at every call: two calls without an explicit seed must add
    different ASE, and reseeding the global stream (randn('state',k) -> gstate.seed(k)) must repeat a run"""
    from polmux_b200 import gstate
    outs = []
    for rep in range(2):
        gstate.seed(77)
        got = []
        for call in range(2):
            make_tx(1 << 8, 16)
            pmx.ampliflat(16.0, 'gain', {'f': 5.0})
            got.append((np.array(pmx.GSTATE.FIELDX), np.array(pmx.GSTATE.FIELDY)))
        outs.append(got)
    assert not np.array_equal(outs[0][0][0], outs[0][1][0])          # successive calls: independent noise
    assert np.array_equal(outs[0][0][0], outs[1][0][0]) and np.array_equal(outs[0][1][1], outs[1][1][1])
    d = outs[0][0][0] - outs[0][1][0]
    assert abs(np.vdot(d, d).real / d.size) > 0


def test_inverse_pmd_gvd_no_and_link_onepol():
    """inverse_pmd(brf, options.gvd = 'no') (inverse_pmd.m:79,135) undoes the PMD only: what is left is the pure GVD of
    the fiber, ifft(fft(u) .* exp(-i*betat*L)); link(..., options.onepol) equals the loop of fiber() / ampliflat() calls"""
    make_tx(1 << 9, 16)
    G = pmx.GSTATE
    tx_x, tx_y = np.array(G.FIELDX), np.array(G.FIELDY)
    fib = base_fiber(length=5e4, dgd=0.6, nplates=15, alphadB=0.0)
    brf = pmx.fiber(fib, 'gp--', rng=np.random.Generator(np.random.PCG64(11)))
    pmx.inverse_pmd(brf, {'gvd': 'no'})
    ph = np.exp(-1j * brf['betat'][:, 0] * fib['length'])
    rx = np.fft.ifft(np.fft.fft(tx_x[:, 0]) * ph)
    ry = np.fft.ifft(np.fft.fft(tx_y[:, 0]) * ph)
    assert rel_l2(G.FIELDX[:, 0], G.FIELDY[:, 0], rx, ry) < 1e-9
    with pytest.raises(ValueError, match='unknown options'):
        pmx.inverse_pmd(brf, {'theta': 0.1})
    # options.mat = a rotation: inverting with it and then rotating back by hand equals the plain inversion
    c, s_ = np.cos(0.4), np.sin(0.4)
    before = (np.array(G.FIELDX), np.array(G.FIELDY))
    G.FIELDX, G.FIELDY = tx_x.copy(), tx_y.copy()
    brf = pmx.fiber(fib, 'gp--', rng=np.random.Generator(np.random.PCG64(11)))
    pmx.inverse_pmd(brf, {'gvd': 'no', 'mat': np.array([[c, s_], [-s_, c]])})
    bx, by = c * G.FIELDX + s_ * G.FIELDY, -s_ * G.FIELDX + c * G.FIELDY
    assert rel_l2(bx, by, before[0], before[1]) < 1e-12
    # link with ASE on one polarization
    fib = base_fiber(length=2e4, dgd=0.2, nplates=8, manakov='yes')
    outs = []
    for use_link in (True, False):
        make_tx(1 << 9, 16)
        rng = np.random.Generator(np.random.PCG64(3))
        if use_link:
            pmx.link(fib, 'gps-', 2, 6.0, {'f': 5.0, 'onepol': 'asey'}, rng=rng, seed=40)
        else:
            for k in range(2):
                pmx.fiber(fib, 'gps-', rng=rng)
                pmx.ampliflat(6.0, 'gain', {'f': 5.0, 'onepol': 'asey'}, seed=40 + k)
        outs.append((np.array(pmx.GSTATE.FIELDX), np.array(pmx.GSTATE.FIELDY)))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
