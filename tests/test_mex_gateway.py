"""The MEX gateway (mex/ssfm_mex.c, compiled against the mex.h stand-in) driven natively with the
argument list a patched fiber.m would pass to it."""
import os
import subprocess

import numpy as np
import pytest

import oracle.fiber_oracle as orc
import polmux_b200 as pmx
from polmux_b200.fiber import fiber_setup
from common import base_fiber, make_tx, rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, 'mex', 'build', 'test_ssfm_mex')


def test_gateway_is_built_and_links():
    assert os.path.exists(EXE), 'run __graft_entry__.build()'
    out = subprocess.run(['ldd', EXE], capture_output=True, text=True).stdout
    assert 'libpolmux_ssfm.so' in out and 'not found' not in out.split('libpolmux_ssfm.so')[1].split('\n')[0]


def _write_inputs(path, s, G, scal):
    nfft, nfc = G.FIELDX.shape
    fy = G.FIELDY if G.FIELDY is not None else np.zeros_like(G.FIELDX)
    hdr = np.array([nfft, nfc, s.nplates, int(s.manakov), *s.fls, 1, int(scal is not None),
                    0 if scal is None else len(scal)], dtype=np.int64)
    with open(path, 'wb') as f:
        hdr.tofile(f)
        np.array([s.dzmaxt, s.dphimaxt, s.alphalin, s.length]).tofile(f)
        np.asarray(s.gam, dtype=np.float64).tofile(f)
        for a in (G.FIELDX, fy):
            np.ascontiguousarray(a.real.T).tofile(f)      # column-major planes
            np.ascontiguousarray(a.imag.T).tofile(f)
        np.ascontiguousarray(s.betat.T).tofile(f)
        np.ascontiguousarray(s.db1.T).tofile(f)
        for k in ('db0', 'theta', 'epsilon'):
            np.asarray(s.brf[k], dtype=np.float64).tofile(f)
        if scal is not None:
            np.asarray(scal, dtype=np.float64).tofile(f)


def _read_outputs(path, nfft, nfc):
    raw = np.fromfile(path, dtype=np.float64)
    status, firstdz, ncycle = raw[:3]
    if status != 0:
        return int(status), None, None, 0, 0
    p = raw[3:].reshape(4, nfc, nfft)
    return 0, (p[0] + 1j * p[1]).T, (p[2] + 1j * p[3]).T, firstdz, int(ncycle)


@pytest.mark.gpu
@pytest.mark.parametrize('mode', ['vector', 'scalar'])
def test_gateway_parity(tmp_path, mode):
    fib = base_fiber(length=1e5, dgd=1.0, nplates=10, manakov='no')
    gs = make_tx(1 << 9, 16)
    orc.fiber(gs, fib, 'gps-', rng=np.random.Generator(np.random.PCG64(3)))
    G = pmx.GSTATE
    s = fiber_setup(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(3)))
    sc = s.scalars
    scal = None if mode == 'vector' else [sc['symbolrate'], sc['nsymb'], sc['nt'], sc['b30'], sc['dgdrms'],
                                          *sc['beta1'], *sc['beta2']]
    _write_inputs(tmp_path / 'in.bin', s, G, scal)
    r = subprocess.run([EXE, str(tmp_path / 'in.bin'), str(tmp_path / 'out.bin')], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    st, ux, uy, firstdz, ncycle = _read_outputs(tmp_path / 'out.bin', *G.FIELDX.shape)
    assert st == 0, r.stderr
    assert rel_l2(ux, uy, gs.FIELDX, gs.FIELDY) < 1e-10
    assert ncycle == gs.log['ncycle'] and abs(firstdz - gs.log['firstdz']) < 1e-9 * gs.log['firstdz']


@pytest.mark.gpu
def test_gateway_reports_errors_through_mexerrmsgtxt(tmp_path):
    """bad arguments and the reference's own plate-index failure come back as mexErrMsgTxt text"""
    make_tx(1 << 8, 16)
    G = pmx.GSTATE
    s = fiber_setup(base_fiber(length=8e4, dgd=0.1, nplates=59), 'gp--', rng=np.random.Generator(np.random.PCG64(1)))
    _write_inputs(tmp_path / 'in.bin', s, G, None)
    r = subprocess.run([EXE, str(tmp_path / 'in.bin'), str(tmp_path / 'out.bin')], capture_output=True, text=True)
    st = _read_outputs(tmp_path / 'out.bin', *G.FIELDX.shape)[0]
    assert st == 1 and 'nplates' in r.stderr
