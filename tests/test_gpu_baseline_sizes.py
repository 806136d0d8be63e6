"""-m gpu: the CUDA path against the numpy oracle at EVERY BASELINE.json size and kernel instance.

The in-CTA transform length L of a pass is N1 (passes A, C) or N2 (pass B) of the four-step split N = N1*N2:
    2^17 -> 256 x 512     2^18 -> 512 x 512     2^19 -> 512 x 1024    2^20 -> 1024 x 1024  (C2, C3, C5)
    2^21 -> 1024 x 2048   2^22 -> 2048 x 2048 (C4)   2^23 -> 2048 x 4096   2^24 -> 4096 x 4096
so the cases below put every L in {256, 512, 1024, 2048, 4096} through the nonlinear + PMD path (fiber.m:459-555,
matrix_step :877-935) on the same seeded Tx field as the oracle (create_field.m:180-199 for the nine-channel
'unique' multiplex of C4).  The oracle needs 0.25 s (2^20) to 4 s (2^24) per trunk and about twice that per
step on one host core, so the long fields run a bounded prefix (>= 4 steps, partial + whole trunks, a plate
boundary; 2^23 shares its two kernel instances with 2^22 and 2^24); C2 runs one full 80 km span, C3 the reference's own 'gp--' call (one step, 200 trunks) in full.

Bar (north_star): FP64 rel-L2 <= 1e-10 on the output field; ncycle and the per-step trunk schedule equal."""
import numpy as np
import pytest

import oracle.fiber_oracle as orc
import polmux_b200 as pmx
from common import base_fiber, make_tx, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _both(gs, fib, flag, seed):
    orc.fiber(gs, fib, flag, rng=np.random.Generator(np.random.PCG64(seed)))
    pmx.fiber(fib, flag, rng=np.random.Generator(np.random.PCG64(seed)), trace=True)
    G, L = pmx.GSTATE, pmx.FIBER_LAST
    err = rel_l2(G.FIELDX, G.FIELDY, gs.FIELDX, gs.FIELDY)
    assert L['ncycle'] == gs.log['ncycle']
    assert list(L['trace_ntrunk'])[:L['ncycle']] == [s['ntrunk'] for s in gs.log['schedule']]
    np.testing.assert_allclose(list(L['trace_dz'])[:L['ncycle']], [s['dz'] for s in gs.log['schedule']], rtol=1e-9)
    return err, L


def test_c4_nine_channel_wdm_prefix_against_oracle():
    """C4: nine 28-GBaud channels multiplexed by create_field('unique') into one field of 2^22 samples (L = 2048, the
    4*8*8*8 transform), Manakov 'gps-' with PMD: the first 800 m of the span (plates of 200 m) -- >= 10 steps, every
    plate boundary crossed"""
    gs = make_tx(1 << 16, 64, nch=9, pavg_mw=1.0)
    fib = base_fiber(length=8e2, dgd=0.1, nplates=4, manakov='yes')
    err, L = _both(gs, fib, 'gps-', 41)
    assert L['ncycle'] >= 10 and L['ntot'] == 4 and sum(L['trace_ntrunk'][:L['ncycle']]) >= L['ncycle'] + 3
    assert err < TOL, err


@pytest.mark.parametrize('lg,nt,length,nplates,manakov', [
    (17, 16, 2.0e4, 10, 'no'),     # 256 x 512, CNLSE
    (18, 16, 2.0e4, 10, 'yes'),    # 512 x 512
    (19, 16, 1.6e4, 10, 'no'),     # 512 x 1024, CNLSE
    (21, 32, 8.0e3, 5, 'yes'),     # 1024 x 2048
    (24, 64, 3.0e3, 3, 'no'),      # 4096 x 4096, CNLSE
])
def test_gps_prefix_against_oracle_every_transform_length(lg, nt, length, nplates, manakov):
    """'gps-' (GVD + slope term of b30, PMD plates, SPM) at the sizes that select L = 256 ... 4096"""
    gs = make_tx((1 << lg) // nt, nt)
    fib = base_fiber(length=length, dgd=0.3, nplates=nplates, manakov=manakov)
    err, L = _both(gs, fib, 'gps-', 1000 + lg)
    assert L['ncycle'] >= 3 and L['ntot'] == nplates
    assert err < TOL, err


def test_c2_full_span_against_oracle():
    """C2: one full span of the ex20 link -- N = 2^20, 80 km, 'gps-' Manakov, 100 random plates, DGD 0.1 symbol"""
    gs = make_tx(1 << 16, 16)
    fib = base_fiber(length=8e4, dgd=0.1, nplates=100, manakov='yes')
    err, L = _both(gs, fib, 'gps-', 1000)
    assert L['ncycle'] > 30 and L['ntot'] == 100
    assert err < TOL, err


def test_c3_gp_200_plates_full_call_against_oracle():
    """C3 with the reference's own flag (ex24_pmd.m:93-100): 'gp--' = one linear step of 200 trunks over 80 km, DGD
    0.5 symbol, N = 2^20 -- the plate chunks beyond the 16 of the step package"""
    gs = make_tx(1 << 16, 16)
    fib = base_fiber(length=8e4, dgd=0.5, nplates=200)
    err, L = _both(gs, fib, 'gp--', 31)
    assert L['ncycle'] == 1 and L['ntot'] == 200
    assert err < TOL, err


def test_c3_gps_200_plate_pitch_prefix_against_oracle():
    """C3 'gps-': plates of 400 m (200 per 80 km span), DGD 0.5 symbol, Manakov: the first 8 km (20 plates)"""
    gs = make_tx(1 << 16, 16)
    fib = base_fiber(length=8e3, dgd=0.5 * np.sqrt(20.0 / 200.0), nplates=20, manakov='yes')
    err, L = _both(gs, fib, 'gps-', 32)
    assert L['ncycle'] > 5 and L['ntot'] == 20
    assert err < TOL, err
