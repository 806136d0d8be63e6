"""-m gpu: the CUDA path at BASELINE.json's full sizes (C2/C3/C5: N = 2^20, C4: N = 2^22 with nine WDM channels).

The numpy oracle needs minutes for a whole span at these sizes, so full-size parity goes through
  * a bounded prefix of the C2 span against the oracle (first 16 km, 20 plates: about 5 s of CPU),
  * closed forms that need no step loop (linear GVD = one per-bin phase, numpy fft),
  * size-independent properties of the path: PMD round trip through inverse_pmd (propagate -> invert -> Tx field),
    energy ratio exp(-alpha L) of a chain of unitary sub-steps, and a batch of realizations giving the bits of the
    same realizations run one by one.
Tolerances: FP64 rel-L2 <= 1e-10 against oracle / closed form, round trip <= 1e-9, energy 1e-12, batch bit-exact."""
import math

import numpy as np
import pytest

import oracle.fiber_oracle as orc
import polmux_b200 as pmx
from polmux_b200 import _lib, mc, synth
from polmux_b200.fiber import fiber_setup, setup_to_desc
from common import base_fiber, make_tx, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-10


def product_tx(nsymb, nt, nch=1, rate=28.0, pavg_mw=2.0, spac=0.4):
    """the product's GSTATE only (no oracle twin): PDM-QPSK, nch channels multiplexed into one field"""
    ex, ey, _, _ = synth.pdm_qpsk(nsymb, nt, nch)
    pmx.reset_all(nsymb, nt, nch)
    G = pmx.GSTATE
    G.SYMBOLRATE, G.LAMBDA, G.POWER = rate, synth.wdm_lambdas(nch, 1550.0, spac), np.full(nch, float(pavg_mw))
    pmx.create_field('unique', ex, ey, {'power': 'average'})
    return G


def energy(G):
    return float(np.sum(np.abs(G.FIELDX) ** 2 + np.abs(G.FIELDY) ** 2))


def test_c2_span_prefix_against_oracle():
    """C2 (N = 2^20, 'gps-' Manakov, plates of 800 m, DGD 0.1): the first 16 km of the span against the oracle"""
    fib = base_fiber(length=16e3, dgd=0.1, nplates=20, manakov='yes')
    gs = make_tx(1 << 16, 16)
    orc.fiber(gs, fib, 'gps-', rng=np.random.Generator(np.random.PCG64(1000)))
    pmx.fiber(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(1000)), trace=True)
    G = pmx.GSTATE
    assert rel_l2(G.FIELDX, G.FIELDY, gs.FIELDX, gs.FIELDY) < TOL
    L = pmx.FIBER_LAST
    assert L['ncycle'] == gs.log['ncycle'] and L['ncycle'] > 5
    assert list(L['trace_ntrunk']) == [s['ntrunk'] for s in gs.log['schedule']]


@pytest.mark.parametrize('lg,nch,nt', [(20, 1, 16), (22, 9, 64)])
def test_linear_gvd_closed_form_full_size(lg, nch, nt):
    """'g---' at C2 and C4 size: u_out = ifft(fft(u) .* exp(-i*betat*L)) * exp(-alpha/2*L), betat as fiber.m:350-356"""
    G = product_tx(1 << lg >> int(math.log2(nt)), nt, nch, pavg_mw=1.0 if nch > 1 else 2.0)
    x0, y0 = np.array(G.FIELDX[:, 0]), np.array(G.FIELDY[:, 0])
    fib = base_fiber(length=8e4)
    s = fiber_setup(fib, 'g---')
    pmx.fiber(fib, 'g---')
    assert pmx.FIBER_LAST['ncycle'] == 1
    ph = np.exp(-1j * s.betat[:, 0] * s.length) * math.exp(-0.5 * s.alphalin * s.length)
    rx, ry = np.fft.ifft(np.fft.fft(x0) * ph), np.fft.ifft(np.fft.fft(y0) * ph)
    assert rel_l2(G.FIELDX[:, 0], G.FIELDY[:, 0], rx, ry) < TOL


def test_c3_pmd_round_trip_full_size():
    """C3 (N = 2^20, 200 waveplates per span, DGD 0.5 symbols, the reference's 'gp--' flag): two spans, then
    inverse_pmd of both brf structs gives the transmitted field back"""
    G = product_tx(1 << 16, 16)
    x0, y0 = np.array(G.FIELDX), np.array(G.FIELDY)
    fib = base_fiber(length=8e4, dgd=0.5, nplates=200, alphadB=0.0)
    b1 = pmx.fiber(fib, 'gp--', rng=np.random.Generator(np.random.PCG64(31)))
    assert pmx.FIBER_LAST['ntot'] == 200 and pmx.FIBER_LAST['ncycle'] == 1
    b2 = pmx.fiber(fib, 'gp--', rng=np.random.Generator(np.random.PCG64(32)))
    assert G.is_resident()
    mid = rel_l2(G.FIELDX, G.FIELDY, x0, y0)
    assert mid > 0.1                                  # the spans did something
    G.FIELDX, G.FIELDY = G.FIELDX, G.FIELDY           # (host round trip in the middle of the chain)
    pmx.inverse_pmd([b1, b2])
    assert rel_l2(G.FIELDX, G.FIELDY, x0, y0) < 1e-9


def test_c4_wdm_span_energy_full_size():
    """C4 (nine 28-GBaud channels in one field of 2^22 samples, Manakov 'gps-'): every sub-step but the attenuation is
    unitary, so sum|u|^2 out/in = exp(-alpha L) whatever the number of steps"""
    G = product_tx(1 << 16, 64, 9, pavg_mw=1.0)
    e_in = energy(G)
    fib = base_fiber(length=2e4, dgd=0.1, nplates=25, manakov='yes')
    pmx.fiber(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(41)))
    assert pmx.FIBER_LAST['ncycle'] > 20 and pmx.FIBER_LAST['ntot'] == 25
    alphalin = math.log(10) * 1e-4 * fib['alphadB']
    assert abs(energy(G) / e_in / math.exp(-alphalin * fib['length']) - 1) < 1e-12


def test_c5_batch_gives_the_bits_of_single_runs_full_size():
    """C5: realizations of C2 differing in the plate draw, propagated as one resident batch (two realization groups)
    and one by one: the same bits, the same step counts"""
    nsymb, nt, batch = 1 << 16, 16, 3
    n = nsymb * nt
    G = product_tx(nsymb, nt)
    fib = base_fiber(length=8e3, dgd=0.1, nplates=10, manakov='yes')
    setup = fiber_setup(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(0)))
    draws = [mc.draw_plates(mc.plate_seed(b, 0), setup.nplates) for b in range(batch)]
    pl = [np.stack([d[i] for d in draws]) for i in range(3)]
    ctx = _lib.default_context()
    tx = _lib.DeviceField(ctx, n, 1, 1)
    tx.upload(G.FIELDX, G.FIELDY)
    work = _lib.DeviceField(ctx, n, 1, batch)
    work.broadcast_from(tx)
    desc, keep = setup_to_desc(setup, batch=batch, plate_sets=batch, db0=pl[0], theta=pl[1], epsilon=pl[2])
    plan = _lib.Plan(ctx, desc, keep)
    res = plan.execute(work)
    gx, gy = work.download()
    plan.close()
    one = _lib.DeviceField(ctx, n, 1, 1)
    for b in range(batch):
        one.broadcast_from(tx)
        d1, k1 = setup_to_desc(setup, batch=1, plate_sets=1, db0=pl[0][b][None], theta=pl[1][b][None],
                               epsilon=pl[2][b][None])
        p1 = _lib.Plan(ctx, d1, k1)
        r1 = p1.execute(one)
        ox, oy = one.download()
        p1.close()
        assert int(r1.ncycle[0]) == int(res.ncycle[b]) > 3
        assert np.array_equal(ox[0], gx[b]) and np.array_equal(oy[0], gy[b])
    assert not np.array_equal(gx[0], gx[1])
    for f in (tx, work, one):
        f.close()


def test_fp32_c2_span_prefix_against_fp64():
    """FP32 mode at C2 size (reported separately, tolerance 1e-5): the first 16 km of the span in FP32 against the FP64
    run of the same call (which test_c2_span_prefix_against_oracle pins on the oracle), same step count; energy ratio
    exp(-alpha L) to FP32 accuracy"""
    fib = base_fiber(length=16e3, dgd=0.1, nplates=20, manakov='yes')
    out = {}
    for prec in ('f64', 'f32'):
        G = product_tx(1 << 16, 16)
        e_in = energy(G)
        pmx.fiber(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(1000)), precision=prec)
        out[prec] = (np.array(G.FIELDX), np.array(G.FIELDY), pmx.FIBER_LAST['ncycle'], energy(G) / e_in)
    assert rel_l2(out['f32'][0], out['f32'][1], out['f64'][0], out['f64'][1]) < 1e-5
    assert abs(out['f32'][2] - out['f64'][2]) <= 1          # the step control sees FP32-rounded maxima
    alphalin = math.log(10) * 1e-4 * fib['alphadB']
    assert abs(out['f32'][3] / math.exp(-alphalin * fib['length']) - 1) < 1e-5
