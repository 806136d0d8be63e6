"""GSTATE.PRINT: the summary block fiber() appends to GSTATE.DIR/simul_out (fiber.m:392-456) against the text the
reference's own source prints (tests/golden/simul_out_*.json, captured from the interpreted fiber.m), and the log
header of reset_all (reset_all.m:176-225)."""
import glob
import json
import os
import re

import numpy as np
import pytest

import oracle.fiber_oracle as orc
import polmux_b200 as pmx
from polmux_b200 import simul_out, synth
from polmux_b200.fiber import apply_side_effects, fiber_setup

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = sorted(glob.glob(os.path.join(GOLD, 'simul_out_*.json')))


def _product_state(m, *opts):
    ex, ey, _, _ = synth.pdm_qpsk(m['nsymb'], m['nt'], m['nch'])
    pmx.reset_all(m['nsymb'], m['nt'], m['nch'], *opts)
    G = pmx.GSTATE
    G.SYMBOLRATE, G.POWER, G.LAMBDA = m['rate'], np.full(m['nch'], float(m['pavg'])), synth.wdm_lambdas(m['nch'])
    pmx.create_field(m['ftype'], ex, ey if m['two_pol'] else None, {'power': 'average'})
    return ex, ey


def test_golden_text_present():
    assert len(CASES) >= 3


@pytest.mark.parametrize('path', CASES, ids=[os.path.basename(p)[10:-5] for p in CASES])
def test_fiber_block_equals_reference_text(path):
    """host side only: first step and step count from the oracle run of the same case"""
    m = json.load(open(path))
    ex, ey = _product_state(m)
    gs = orc.reset_all(m['nsymb'], m['nt'], m['nch'])
    gs.SYMBOLRATE, gs.POWER, gs.LAMBDA = m['rate'], np.full(m['nch'], float(m['pavg'])), synth.wdm_lambdas(m['nch'])
    orc.create_field(gs, m['ftype'], ex, ey if m['two_pol'] else None, power_average=True)
    orc.fiber(gs, m['fiber'], m['flag'], rng=np.random.Generator(np.random.PCG64(m['seed'])))
    s = fiber_setup(m['fiber'], m['flag'], rng=np.random.Generator(np.random.PCG64(m['seed'])))
    apply_side_effects(s)
    text = simul_out.fiber_block(m['fiber'], m['flag'], s, gs.log['firstdz'], gs.log['ncycle'])
    assert text == m['text']


def test_reset_all_print_options(tmp_path):
    """reset_all.m:121-150: outdir -> PRINT, 'noprint' in either position -> no log"""
    d = str(tmp_path / 'sim')
    pmx.reset_all(64, 8, 2, d)
    assert pmx.GSTATE.PRINT and pmx.GSTATE.DIR == d
    log = open(os.path.join(d, 'simul_out')).read()
    assert 'START OF SIMULATION' in log and 'Nsymb =     64\t (number of symbols)\n' in log
    assert 'Nch  =      2\t (number of channels)\n' in log and ('Output directory = %s\n' % d) in log
    assert re.search(r'\+\+\+\+ Date: \d\d-[A-Z][a-z][a-z]-\d{4} \d\d:\d\d:\d\d\t\n', log)
    assert os.path.isdir(os.path.join(d, 'sim.MOD')) and os.path.isdir(os.path.join(d, 'sim.ANG'))
    for opts in (('noprint', d), (d, 'noprint')):
        pmx.reset_all(64, 8, 2, *opts)
        assert not pmx.GSTATE.PRINT and pmx.GSTATE.DIR == d
    assert open(os.path.join(d, 'simul_out')).read() == log
    for bad in (('noprint',), (d, 'print'), ('noprint', 'noprint'), (3,)):
        with pytest.raises(ValueError):
            pmx.reset_all(64, 8, 2, *bad)
    pmx.reset_all(64, 8, 2)
    assert not pmx.GSTATE.PRINT


@pytest.mark.gpu
@pytest.mark.parametrize('path', CASES, ids=[os.path.basename(p)[10:-5] for p in CASES])
def test_fiber_appends_the_reference_block(path, tmp_path):
    """fiber() on the device with GSTATE.PRINT: the file ends with the reference's block (first step, number of steps,
    delays and cumulated dispersion included)"""
    m = json.load(open(path))
    d = str(tmp_path / 'sim')
    _product_state(m, d)
    pmx.fiber(m['fiber'], m['flag'], rng=np.random.Generator(np.random.PCG64(m['seed'])))
    log = open(os.path.join(d, 'simul_out')).read()
    assert log.endswith(m['text']) and log.count('===              fiber               ===') == 1
