"""inverse_pmd.m with its options (options.mat, options.gvd, options.apply; [Uinv,U] outputs), pinned on the interpreted
reference file (oracle/make_golden.py invpmd -> tests/golden/invpmd/): the numpy restatement on the CPU, the device path
(pmx.inverse_pmd: reversed / negated single-step plans + pmx_field_jones + pmx_pmd_matrix) on the GPU."""
import json
import os

import numpy as np
import pytest

import oracle.fiber_oracle as orc
from polmux_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'invpmd', 'invpmd_two_fibers.npz')
VARIANTS = {'plain': None, 'mat_nogvd': {'mat': 'MAT', 'gvd': 'no'}, 'noapply': {'apply': 'no'},
            'apply_n': {'apply': 'n', 'mat': 'MAT'}}


def _load():
    z = np.load(GOLD)
    m = json.loads(str(z['meta']))
    brfs = []
    for k in range(2):
        brfs.append({key: z['brf%d_%s' % (k, key)] for key in ('db0', 'theta', 'epsilon', 'betat', 'db1')})
        brfs[-1]['lcorr'] = float(z['brf%d_lcorr' % k][0])
    return z, m, brfs


def _options(tag, z):
    opt = VARIANTS[tag]
    return None if opt is None else {k: (z['mat'] if v == 'MAT' else v) for k, v in opt.items()}


@pytest.mark.parametrize('tag', list(VARIANTS))
def test_oracle_inverse_pmd_options_match_reference_source(tag):
    z, m, brfs = _load()
    gs = orc.reset_all(m['nsymb'], m['nt'], 1)
    gs.FIELDX, gs.FIELDY = z['prop_FIELDX'].copy(), z['prop_FIELDY'].copy()
    gs.DISP = np.ones((2, 1))
    uinv, u = orc.inverse_pmd(gs, brfs, _options(tag, z))
    np.testing.assert_allclose(uinv, z[tag + '_Uinv'], rtol=0, atol=1e-13)
    np.testing.assert_allclose(u, z[tag + '_U'], rtol=0, atol=1e-13)
    assert orc.rel_l2(gs.FIELDX, gs.FIELDY, z[tag + '_FIELDX'], z[tag + '_FIELDY']) < 1e-13
    if tag == 'noapply':       # the field is left alone, bit for bit
        assert np.array_equal(gs.FIELDX, z['prop_FIELDX']) and np.array_equal(z[tag + '_FIELDX'], z['prop_FIELDX'])


@pytest.mark.gpu
@pytest.mark.parametrize('tag', list(VARIANTS))
def test_cuda_inverse_pmd_options_match_reference_source(tag):
    import polmux_b200 as pmx
    z, m, brfs = _load()
    pmx.reset_all(m['nsymb'], m['nt'], 1)
    G = pmx.GSTATE
    G.SYMBOLRATE, G.POWER, G.LAMBDA = m['rate'], np.array([float(m['pavg'])]), synth.wdm_lambdas(1)
    G.FIELDX, G.FIELDY = z['prop_FIELDX'].copy(), z['prop_FIELDY'].copy()
    G.DISP = np.ones((2, 1))
    uinv, u = pmx.inverse_pmd(brfs, _options(tag, z), nargout=2)
    assert uinv.shape == (2, 2, m['nsymb'] * m['nt']) and u.shape == uinv.shape
    np.testing.assert_allclose(uinv, z[tag + '_Uinv'], rtol=0, atol=1e-11)
    np.testing.assert_allclose(u, z[tag + '_U'], rtol=0, atol=1e-11)
    assert orc.rel_l2(G.FIELDX, G.FIELDY, z[tag + '_FIELDX'], z[tag + '_FIELDY']) < 1e-10
    applied = tag != 'noapply'
    np.testing.assert_array_equal(G.DISP, np.zeros((2, 1)) if applied else np.ones((2, 1)))     # inverse_pmd.m:141
    assert np.all(z[tag + '_DISP'] == 0) == applied
    if tag == 'noapply':
        assert np.array_equal(G.FIELDX, z['prop_FIELDX']) and np.array_equal(G.FIELDY, z['prop_FIELDY'])
