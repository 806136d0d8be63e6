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


def _write_inputs(path, s, G, scal, plates=None, amp=None):
    """plates: (db0, theta, epsilon) each [nspan, nplates] for a span loop; amp: [gain, sigma..., seed]"""
    nfft, nfc = G.FIELDX.shape
    fy = G.FIELDY if G.FIELDY is not None else np.zeros_like(G.FIELDX)
    nspan = 1 if plates is None else len(plates[0])
    hdr = np.array([nfft, nfc, s.nplates, int(s.manakov), *s.fls, 1, int(scal is not None),
                    0 if scal is None else len(scal), nspan, 0 if amp is None else len(amp)], dtype=np.int64)
    with open(path, 'wb') as f:
        hdr.tofile(f)
        np.array([s.dzmaxt, s.dphimaxt, s.alphalin, s.length]).tofile(f)
        np.asarray(s.gam, dtype=np.float64).tofile(f)
        for a in (G.FIELDX, fy):
            np.ascontiguousarray(a.real.T).tofile(f)      # column-major planes
            np.ascontiguousarray(a.imag.T).tofile(f)
        np.ascontiguousarray(s.betat.T).tofile(f)
        np.ascontiguousarray(s.db1.T).tofile(f)
        for i, k in enumerate(('db0', 'theta', 'epsilon')):
            np.ascontiguousarray(s.brf[k] if plates is None else plates[i], dtype=np.float64).tofile(f)
        if scal is not None:
            np.asarray(scal, dtype=np.float64).tofile(f)
        if amp is not None:
            np.asarray(amp, dtype=np.float64).tofile(f)


def _read_outputs(path, nfft, nfc):
    raw = np.fromfile(path, dtype=np.float64)
    status, firstdz, ncycle = raw[:3]
    if status != 0:
        return int(status), None, None, 0, 0
    p = raw[3:3 + 4 * nfc * nfft].reshape(4, nfc, nfft)
    _read_outputs.ncycles = [int(v) for v in raw[3 + 4 * nfc * nfft:]]
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


@pytest.mark.gpu
@pytest.mark.parametrize('mode', ['vector', 'scalar'])
def test_gateway_span_loop(tmp_path, mode):
    """plates as nplates x Nspan matrices + amp = [gain sigma seed]: the loop  fiber ; ampliflat  of the reference's
    scripts in one gateway call (pmx_link_run).  Noiseless amplifiers against the oracle loop; with ASE from the device
    generator against the bits of polmux_b200.link()."""
    fib = base_fiber(length=3e4, dgd=0.5, nplates=8, manakov='yes')
    nspan, gain_db = 3, 6.0
    gs = make_tx(1 << 9, 16)
    G = pmx.GSTATE
    tx = (np.array(G.FIELDX), np.array(G.FIELDY))
    r = np.random.Generator(np.random.PCG64(5))
    setups = [fiber_setup(fib, 'gps-', rng=r) for _ in range(nspan)]
    plates = [np.stack([st.brf[k] for st in setups]) for k in ('db0', 'theta', 'epsilon')]
    s = setups[0]
    sc = s.scalars
    scal = None if mode == 'vector' else [sc['symbolrate'], sc['nsymb'], sc['nt'], sc['b30'], sc['dgdrms'],
                                          *sc['beta1'], *sc['beta2']]
    ro = np.random.Generator(np.random.PCG64(5))
    ncyc = []
    for k in range(nspan):
        orc.fiber(gs, fib, 'gps-', rng=ro)
        ncyc.append(gs.log['ncycle'])
        orc.ampliflat(gs, gain_db)
    for sigma, seed in ((0.0, 0), (0.02, 41)):
        _write_inputs(tmp_path / 'in.bin', s, G if sigma == 0.0 else _Tx(tx), scal, plates,
                      [10 ** (gain_db / 10), sigma, seed])
        rr = subprocess.run([EXE, str(tmp_path / 'in.bin'), str(tmp_path / 'out.bin')], capture_output=True, text=True)
        assert rr.returncode == 0, rr.stderr
        st, ux, uy, firstdz, ncycle = _read_outputs(tmp_path / 'out.bin', *tx[0].shape)
        assert st == 0, rr.stderr
        if sigma == 0.0:
            assert rel_l2(ux, uy, gs.FIELDX, gs.FIELDY) < 1e-10
            assert _read_outputs.ncycles == ncyc
        else:
            # the same link through the Python mirror with the device noise generator: same library call, same bits
            from polmux_b200 import _lib
            from polmux_b200.fiber import setup_to_desc
            import ctypes
            ctx = _lib.default_context()
            desc, keep = setup_to_desc(s, disp_mode=mode)
            ld, lkeep = _lib.make_link(nspan, 10 ** (gain_db / 10), [sigma], plates=[p[:, None, :] for p in plates],
                                       plate_sets=1, seeds=[seed + k for k in range(nspan)])
            fx = np.ascontiguousarray(tx[0].T)[None].copy()
            fy = np.ascontiguousarray(tx[1].T)[None].copy()
            io = _lib.complex_field(fx, fy)
            res = _lib.Result(nspan)
            ctx.check(ctx.lib.pmx_link_run(ctx.h, ctypes.byref(desc), ctypes.byref(ld), ctypes.byref(io),
                                           ctypes.byref(res.c)))
            assert np.array_equal(ux, fx[0].T) and np.array_equal(uy, fy[0].T)
            assert rel_l2(ux, uy, gs.FIELDX, gs.FIELDY) > 1e-6      # the ASE is there


class _Tx:
    def __init__(self, tx):
        self.FIELDX, self.FIELDY = tx
