"""Golden vectors produced by executing the reference's own fiber.m / create_field.m / ampliflat.m
source text (oracle/make_golden.py + oracle/mini_m) pin the numpy oracle on the CPU and the CUDA
path on the GPU."""
import glob
import json
import os

import numpy as np
import pytest

import oracle.fiber_oracle as orc
import polmux_b200 as pmx
from polmux_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, '*.npz')))


def load(name):
    z = np.load(os.path.join(GOLD, name + '.npz'))
    return z, json.loads(str(z['meta']))


def test_fixtures_present():
    assert len(CASES) >= 10


@pytest.mark.parametrize('name', CASES)
def test_inputs_reproduce(name):
    """the committed inputs are exactly what the seeded generator yields today"""
    z, m = load(name)
    ex, ey, _, _ = synth.pdm_qpsk(m['nsymb'], m['nt'], m['nch'])
    np.testing.assert_array_equal(ex, z['in_ex'])
    np.testing.assert_array_equal(ey, z['in_ey'])


def _oracle_run(z, m):
    gs = orc.reset_all(m['nsymb'], m['nt'], m['nch'])
    gs.SYMBOLRATE, gs.POWER = m['rate'], np.full(m['nch'], float(m['pavg']))
    gs.LAMBDA = synth.wdm_lambdas(m['nch'])
    orc.create_field(gs, m['ftype'], z['in_ex'], z['in_ey'] if m['two_pol'] else None, power_average=True)
    tx = (gs.FIELDX.copy(), None if gs.FIELDY is None else gs.FIELDY.copy())
    brf = orc.fiber(gs, m['fiber'], m['flag'], rng=np.random.Generator(np.random.PCG64(m['seed'])))
    return gs, tx, brf


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_source(name):
    """numpy restatement == interpreted reference, to rounding (same libm / pocketfft underneath)."""
    z, m = load(name)
    gs, tx, brf = _oracle_run(z, m)
    # create_field.m
    np.testing.assert_allclose(tx[0], z['tx_FIELDX'], rtol=0, atol=1e-14)
    if m['two_pol']:
        np.testing.assert_allclose(tx[1], z['tx_FIELDY'], rtol=0, atol=1e-14)
    # fiber.m
    if z['out_FIELDY'].size:
        err = orc.rel_l2(gs.FIELDX, gs.FIELDY, z['out_FIELDX'], z['out_FIELDY'])
    else:
        err = float(np.linalg.norm(gs.FIELDX - z['out_FIELDX']) / np.linalg.norm(z['out_FIELDX']))
    assert err < 1e-13, err
    np.testing.assert_allclose(gs.DELAY, z['out_DELAY'], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(gs.DISP, z['out_DISP'], rtol=1e-13, atol=1e-13)
    if 'brf_theta' in z.files:
        np.testing.assert_allclose(brf['theta'], z['brf_theta'], rtol=0, atol=0)
        np.testing.assert_allclose(brf['epsilon'], z['brf_epsilon'], rtol=0, atol=0)
        np.testing.assert_allclose(brf['db0'], z['brf_db0'], rtol=0, atol=0)
        np.testing.assert_allclose(brf['betat'], z['brf_betat'], rtol=1e-15, atol=0)
        np.testing.assert_allclose(brf['db1'], z['brf_db1'], rtol=1e-15, atol=0)
        assert abs(brf['lcorr'] - float(z['brf_lcorr'][0])) == 0
    # ampliflat.m with options.noise
    if 'amp_FIELDX' in z.files:
        orc.ampliflat(gs, m['amp']['gain'], m['amp']['f'], noise=z['amp_noise'], atype=m['amp'].get('atype', 'gain'))
        if z['amp_FIELDY'].size:
            assert orc.rel_l2(gs.FIELDX, gs.FIELDY, z['amp_FIELDX'], z['amp_FIELDY']) < 1e-14
        else:
            assert np.linalg.norm(gs.FIELDX - z['amp_FIELDX']) / np.linalg.norm(z['amp_FIELDX']) < 1e-14


VECTOR_CASES = [c for c in CASES if load(c)[1]['two_pol']]


@pytest.mark.gpu
@pytest.mark.parametrize('name', VECTOR_CASES)
def test_cuda_matches_reference_source(name):
    """CUDA path (fiber() -> pmx_fiber_run) against the interpreted reference: <= 1e-10 rel L2 (FP64)."""
    z, m = load(name)
    pmx.reset_all(m['nsymb'], m['nt'], m['nch'])
    G = pmx.GSTATE
    G.SYMBOLRATE, G.POWER, G.LAMBDA = m['rate'], np.full(m['nch'], float(m['pavg'])), synth.wdm_lambdas(m['nch'])
    pmx.create_field(m['ftype'], z['in_ex'], z['in_ey'], {'power': 'average'})
    pmx.fiber(m['fiber'], m['flag'], rng=np.random.Generator(np.random.PCG64(m['seed'])))
    err = orc.rel_l2(G.FIELDX, G.FIELDY, z['out_FIELDX'], z['out_FIELDY'])
    assert err < 1e-10, err
    np.testing.assert_allclose(G.DELAY, z['out_DELAY'], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(G.DISP, z['out_DISP'], rtol=1e-13, atol=1e-13)
    if 'amp_FIELDX' in z.files:
        pmx.ampliflat(m['amp']['gain'], m['amp'].get('atype', 'gain'), {'f': m['amp']['f'], 'noise': z['amp_noise']})
        assert orc.rel_l2(G.FIELDX, G.FIELDY, z['amp_FIELDX'], z['amp_FIELDY']) < 1e-10


SCALAR_CASES = [c for c in CASES if not load(c)[1]['two_pol']]


@pytest.mark.gpu
@pytest.mark.parametrize('name', SCALAR_CASES)
def test_cuda_scalar_path_matches_reference_source(name):
    """The scalar path (scalar_ssfm: FIELDY empty, no 'p' flag; nl_step with SPM and cross-column XPM,
    fiber.m:557-636,786-803; with x.ltol the local-error adaptive step scalar_a_ssfm / adaptssfm, :639-679,938-1010)
    on the device against the interpreted reference: <= 1e-10 rel L2 (FP64)."""
    z, m = load(name)
    pmx.reset_all(m['nsymb'], m['nt'], m['nch'])
    G = pmx.GSTATE
    G.SYMBOLRATE, G.POWER, G.LAMBDA = m['rate'], np.full(m['nch'], float(m['pavg'])), synth.wdm_lambdas(m['nch'])
    pmx.create_field(m['ftype'], z['in_ex'], None, {'power': 'average'})
    pmx.fiber(m['fiber'], m['flag'], rng=np.random.Generator(np.random.PCG64(m['seed'])))
    err = float(np.linalg.norm(G.FIELDX - z['out_FIELDX']) / np.linalg.norm(z['out_FIELDX']))
    assert err < 1e-10, err
    assert G.FIELDY is None or np.size(G.FIELDY) == 0
    np.testing.assert_allclose(G.DELAY, z['out_DELAY'], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(G.DISP, z['out_DISP'], rtol=1e-13, atol=1e-13)
    if 'amp_FIELDX' in z.files:       # ampliflat on a one-polarization field: the ASE creates FIELDY (ampliflat.m:132-146)
        pmx.ampliflat(m['amp']['gain'], m['amp'].get('atype', 'gain'), {'f': m['amp']['f'], 'noise': z['amp_noise']})
        assert orc.rel_l2(G.FIELDX, G.FIELDY, z['amp_FIELDX'], z['amp_FIELDY']) < 1e-10


# ---- BASELINE config C1 at its full size (2^16 samples), from the interpreted reference (oracle/make_golden.py c1) ----
BIG = os.path.join(GOLD, 'big', 'c1_cnlse_10plates_100km_2e16.npz')


def _load_c1():
    import hashlib
    z = np.load(BIG)
    m = json.loads(str(z['meta']))
    ex, ey, _, _ = synth.pdm_qpsk(m['nsymb'], m['nt'], 1)
    assert hashlib.sha256(np.ascontiguousarray(ex).tobytes() + np.ascontiguousarray(ey).tobytes()).hexdigest() == m['sha256_in']
    return z, m, ex, ey


def test_c1_full_size_oracle_matches_reference_source():
    """C1 (Run_my_PDM_QPSK: 2^12 symbols x 16 samples, 100 km, 'gps-' CNLSE, 10 plates): numpy restatement ==
    interpreted create_field.m + fiber.m at the configuration's own size"""
    import hashlib
    z, m, ex, ey = _load_c1()
    gs = orc.reset_all(m['nsymb'], m['nt'], 1)
    gs.SYMBOLRATE, gs.POWER, gs.LAMBDA = m['rate'], np.array([float(m['pavg'])]), synth.wdm_lambdas(1)
    orc.create_field(gs, 'unique', ex, ey, power_average=True)
    assert abs(np.sum(np.abs(gs.FIELDX) ** 2 + np.abs(gs.FIELDY) ** 2) / float(z['tx_power_sum'][0]) - 1) < 1e-14
    brf = orc.fiber(gs, m['fiber'], m['flag'], rng=np.random.Generator(np.random.PCG64(m['seed'])))
    assert orc.rel_l2(gs.FIELDX, gs.FIELDY, z['out_FIELDX'], z['out_FIELDY']) < 1e-13
    for k in ('db0', 'theta', 'epsilon'):
        np.testing.assert_array_equal(brf[k], z['brf_' + k])
    np.testing.assert_allclose(gs.DELAY, z['out_DELAY'], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(gs.DISP, z['out_DISP'], rtol=1e-13, atol=1e-13)


@pytest.mark.gpu
def test_c1_full_size_cuda_matches_reference_source():
    """C1 at full size on the CUDA path against the interpreted reference: <= 1e-10 rel L2 (FP64)"""
    z, m, ex, ey = _load_c1()
    pmx.reset_all(m['nsymb'], m['nt'], 1)
    G = pmx.GSTATE
    G.SYMBOLRATE, G.POWER, G.LAMBDA = m['rate'], np.array([float(m['pavg'])]), synth.wdm_lambdas(1)
    pmx.create_field('unique', ex, ey, {'power': 'average'})
    brf = pmx.fiber(m['fiber'], m['flag'], rng=np.random.Generator(np.random.PCG64(m['seed'])))
    err = orc.rel_l2(G.FIELDX, G.FIELDY, z['out_FIELDX'], z['out_FIELDY'])
    assert err < 1e-10, err
    for k in ('db0', 'theta', 'epsilon'):
        np.testing.assert_array_equal(brf[k], z['brf_' + k])
    np.testing.assert_allclose(G.DELAY, z['out_DELAY'], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(G.DISP, z['out_DISP'], rtol=1e-13, atol=1e-13)
