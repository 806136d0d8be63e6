"""CPU: the C-ABI library loads and exports every symbol include/polmux_ssfm.h declares, fails
loudly without a GPU, and the host-side mirror of the reference interface behaves like it."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

import oracle.fiber_oracle as orc
import polmux_b200 as pmx
from polmux_b200 import _lib, synth
from polmux_b200.fiber import fiber_setup, flag_to_fls
from common import base_fiber, make_tx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, 'include', 'polmux_ssfm.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(pmx_[a-z0-9_]+)\s*\(', txt)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), 'libpolmux_ssfm.so does not export %s' % n
    assert sorted(_lib.EXPORTS) == names
    assert lib.pmx_version() == 100


def test_struct_layout_matches_header(tmp_path):
    """compile a probe against the header with gcc: sizes and field offsets equal the ctypes mirror"""
    import subprocess
    structs = {'pmx_fiber_desc': _lib.FiberDesc, 'pmx_field': _lib.Field, 'pmx_fiber_result': _lib.FiberResult,
               'pmx_link_desc': _lib.LinkDesc, 'pmx_dsp_desc': _lib.DspDesc, 'pmx_mc_desc': _lib.McDesc,
               'pmx_mc_receiver': _lib.McReceiver, 'pmx_brf': _lib.BrfDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "polmux_ssfm.h"', 'int main(void){']
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines.append('return 0;}')
    src = tmp_path / 'probe.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'probe'
    subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out['%s.%s' % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


@pytest.mark.skipif(_lib.load().pmx_device_count() > 0, reason='a GPU is present')
def test_no_cpu_fallback():
    with pytest.raises(_lib.PolmuxError) as e:
        _lib.Context(0)
    assert e.value.code == _lib.PMX_ERR_CUDA
    make_tx(256, 16)
    with pytest.raises(_lib.PolmuxError):
        pmx.fiber(base_fiber(), 'g---')


def test_host_setup_equals_oracle_setup():
    """fiber.m:126-369 restated twice (product front-end, oracle): identical scalars and vectors."""
    fib = base_fiber(length=8e4, dgd=0.4, nplates=20, manakov='yes', slope=0.057)
    gs = make_tx(128, 16, nch=3, ftype='sepfields')
    s = fiber_setup(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(5)))
    brf = orc.fiber(gs, fib, 'gps-', rng=np.random.Generator(np.random.PCG64(5)))
    np.testing.assert_array_equal(s.betat, brf['betat'])
    np.testing.assert_array_equal(s.db1, brf['db1'])
    np.testing.assert_array_equal(s.brf['theta'], brf['theta'])
    np.testing.assert_array_equal(s.brf['epsilon'], brf['epsilon'])
    np.testing.assert_array_equal(s.brf['db0'], brf['db0'])
    np.testing.assert_array_equal(s.gam, gs.log['gam'])
    assert s.alphalin == gs.log['alphalin']
    assert s.manakov and s.nplates == 20 and s.fls == (1, 1, 1, 0)


def test_flag_semantics():
    x = {'length': 1e5, 'dzmax': 2e4, 'dphimax': 5e-3}
    assert flag_to_fls('gp--', 1, x) == ((1, 1, 0, 0), math.inf, 1e5)
    assert flag_to_fls('--sx', 1, x) == ((0, 0, 1, 0), math.inf, 1e5)
    assert flag_to_fls('--sx', 2, x) == ((0, 0, 1, 1), 5e-3, 2e4)
    assert flag_to_fls('g-sx', 1, x) == ((1, 0, 1, 0), 5e-3, 2e4)
    for bad in ('---x', 'g--x', '-p-x', 'gp-x'):
        with pytest.raises(ValueError):
            flag_to_fls(bad, 1, x)
    with pytest.raises(ValueError):
        flag_to_fls('-s--', 1, x)
    for f in orc._FLAGS:
        for nfc in (1, 2):
            try:
                a = orc.parse_flag(f, nfc, x)
            except ValueError:
                with pytest.raises(ValueError):
                    flag_to_fls(f, nfc, x)
                continue
            b = flag_to_fls(f, nfc, x)
            assert (tuple(a[0]), a[1], a[2]) == b


def test_setup_errors_like_the_reference():
    make_tx(64, 16)
    with pytest.raises(ValueError, match='Missing DGD'):
        fiber_setup(base_fiber(), 'gp--')
    with pytest.raises(ValueError, match='Missing one of'):
        fiber_setup(base_fiber(dgd=0.1, theta=[0.1]), 'gp--')
    s = fiber_setup(base_fiber(dgd=0.1, theta=[0.1, 0.2], epsilon=[0.0, 0.1], db0=[1.0, 2.0]), 'gp--')
    assert s.nplates == 2                                   # PMF: nplates = length(theta), fiber.m:265
    s = fiber_setup(base_fiber(dzmax=1e9), 'g-s-')
    assert s.dzmaxt == 8e4                                  # dzmax clipped to the length, fiber.m:139-141
    s = fiber_setup(base_fiber(manakov='yes'), 'g-s-')
    assert not s.manakov and s.nplates == 1                 # forced without 'p', fiber.m:296-297


def test_create_field_matches_oracle():
    for ftype, nch in (('sepfields', 2), ('unique', 3)):
        gs = make_tx(64, 32, nch=nch, ftype=ftype)
        G = pmx.GSTATE
        np.testing.assert_allclose(G.FIELDX, gs.FIELDX, rtol=0, atol=1e-15)
        np.testing.assert_allclose(G.FIELDY, gs.FIELDY, rtol=0, atol=1e-15)
        np.testing.assert_allclose(G.POWER, gs.POWER, rtol=1e-15)
        assert G.FIELDX.shape == ((64 * 32, nch) if ftype == 'sepfields' else (64 * 32, 1))


def test_reset_all_frequency_grid():
    G = pmx.reset_all(8, 4, 1)
    assert G.FN[0] == 0.0 and G.FN[1] == 0.125 and G.FN[16] == -2.0 and G.FN[-1] == -0.125
    np.testing.assert_array_equal(G.FN, orc.reset_all(8, 4, 1).FN)


def test_synth_is_deterministic_and_unit_power():
    ex, ey, sx, sy = synth.pdm_qpsk(64, 16, 2)
    ex2, _, sx2, _ = synth.pdm_qpsk(64, 16, 2)
    np.testing.assert_array_equal(ex, ex2)
    np.testing.assert_array_equal(sx, sx2)
    assert ex.shape == (1024, 2) and set(np.unique(sx)) <= {0, 1, 2, 3}
    # symbol centres carry the symbols
    c = ex[0::16, 0] * np.sqrt(2)
    np.testing.assert_allclose(np.sign(c.real), 2 * (sx[:, 0] & 1) - 1)


class _FakeDeviceField:
    """stands in for _lib.DeviceField in the residency bookkeeping of GSTATE (no GPU here)"""

    def __init__(self, x, y):
        self.nfft, self.nfc, self.batch, self.ctx, self.precision = x.shape[0], x.shape[1], 1, 'ctx', _lib.PMX_F64
        self.x, self.y, self.closed, self.downloads = x.copy(), y.copy(), False, 0

    def download(self):
        self.downloads += 1
        return np.ascontiguousarray(self.x.T)[None], np.ascontiguousarray(self.y.T)[None]

    def download_into(self, x, y):
        self.downloads += 1
        x[...] = self.x
        y[...] = self.y

    def close(self):
        self.closed = True


def test_gstate_resident_field_reads_like_the_host_arrays(monkeypatch):
    """GSTATE.FIELDX/FIELDY of a two-polarization link may live in HBM between in-line devices (gstate.RESIDENT):
    reading or assigning either downloads once and gives the device copy up.  Ordinary host arrays the caller may
    still hold are never overwritten (the interpreter's value semantics); PINNED arrays are download targets."""
    from polmux_b200 import gstate
    pmx.reset_all(8, 4, 1)
    G = pmx.GSTATE
    # ordinary arrays: x0 = GSTATE.FIELDX; fiber(); x1 = GSTATE.FIELDX leaves x0 alone
    monkeypatch.setattr(_lib, 'host_is_pinned', lambda a: False)
    x0 = np.zeros((32, 1), dtype=np.complex128)
    y0 = np.zeros((32, 1), dtype=np.complex128)
    G.FIELDX, G.FIELDY = x0, y0
    fk = _FakeDeviceField(x0 + 3.0, y0 + 5.0)
    G.put_device(fk, x0, y0)
    x1 = G.FIELDX
    assert x1 is not x0 and np.all(x0 == 0) and np.all(x1 == 3.0) and np.all(G.FIELDY == 5.0) and fk.closed
    # a failed device call gives the host arrays back untouched
    fk = _FakeDeviceField(x0, y0)
    G.FIELDX, G.FIELDY = x0, y0
    G.__dict__['_taken_resident'] = False
    G.restore_host(fk, x0, y0)
    assert fk.closed and G.FIELDX is x0 and G.FIELDY is y0
    # pinned arrays (from here on): results land in them
    monkeypatch.setattr(_lib, 'host_is_pinned', lambda a: True)
    pmx.reset_all(8, 4, 1)
    G = pmx.GSTATE
    hx = np.zeros((32, 1), dtype=np.complex128)
    hy = np.zeros((32, 1), dtype=np.complex128)
    G.FIELDX, G.FIELDY = hx, hy
    assert G.FIELDX is hx and G.has_y() and G.field_shape() == (32, 1) and not G.is_resident()
    fake = _FakeDeviceField(hx + 1.0, hy + 2.0j)
    G.put_device(fake, hx, hy)
    assert G.is_resident() and G.field_shape() == (32, 1) and G.has_y() and fake.downloads == 0
    assert G.take_device('ctx', _lib.PMX_F64)[0] is fake and not G.is_resident()    # the next device finds it there
    G.put_device(fake, hx, hy)
    assert G.FIELDY is hy and fake.downloads == 1 and fake.closed and not G.is_resident()  # in place, as the reference
    assert np.all(hx == 1.0) and np.all(hy == 2.0j) and G.FIELDX is hx and fake.downloads == 1
    # assigning one polarization keeps the other
    fake2 = _FakeDeviceField(hx + 4.0, hy + 4.0)
    G.put_device(fake2, None, None)
    G.FIELDX = np.full((32, 1), 7.0 + 0j)
    assert fake2.closed and np.all(G.FIELDY == 4.0 + 2.0j) and np.all(G.FIELDX == 7.0)
    # multi-column fields come back as [N, nfc]
    x3 = np.arange(96, dtype=np.complex128).reshape(32, 3)
    G.put_device(_FakeDeviceField(x3, -x3), None, None)
    assert G.field_shape() == (32, 3) and np.array_equal(G.FIELDX, x3) and np.array_equal(G.FIELDY, -x3)
    # per-call mode: nothing stays behind
    old = gstate.RESIDENT
    try:
        gstate.RESIDENT = False
        fake3 = _FakeDeviceField(hx, hy)
        G.put_device(fake3, hx, hy)
        assert not G.is_resident() and fake3.closed
    finally:
        gstate.RESIDENT = old
    # reset_all drops a resident field
    fake4 = _FakeDeviceField(hx, hy)
    G.put_device(fake4, hx, hy)
    pmx.reset_all(8, 4, 1)
    assert fake4.closed and pmx.GSTATE.FIELDX is None and not pmx.GSTATE.has_y()


def test_link_descriptor_checks():
    """make_link (pmx_link_desc): one plate draw and one ASE seed per span"""
    pl = [np.zeros((3, 1, 4))] * 3
    l, keep = _lib.make_link(3, 2.0, [0.1], plates=pl, plate_sets=1, seeds=[5, 6, 7])
    assert l.nspan == 3 and l.plate_sets == 1 and l.gain == 2.0 and keep['seeds'].dtype == np.uint64
    assert keep['db0'].size == 12 and not l.noise
    with pytest.raises(ValueError):
        _lib.make_link(3, 2.0, [0.1], seeds=[1, 2])
    with pytest.raises(ValueError):
        _lib.make_link(3, 2.0, [0.1], plates=[np.zeros(10)] * 3)
    l0, _ = _lib.make_link(2)
    assert not l0.db0 and not l0.sigma and not l0.seeds and l0.gain == 0.0


def test_ampliflat_argument_errors_like_the_reference():
    """ampliflat.m:65-75: unknown atype; 'fixpower' on a field that is not one column per channel (raised on the host,
    before any device call)"""
    import polmux_b200 as pmx
    pmx.reset_all(16, 8, 3)
    G = pmx.GSTATE
    G.SYMBOLRATE, G.LAMBDA, G.POWER = 10.0, np.array([1549.6, 1550.0, 1550.4]), np.ones(3)
    G.FIELDX = np.ones((128, 1), dtype=np.complex128)
    G.FIELDY = np.ones((128, 1), dtype=np.complex128)
    with pytest.raises(ValueError, match='wrong string atype'):
        pmx.ampliflat(3.0, 'boost')
    with pytest.raises(ValueError, match='only for channels separated'):
        pmx.ampliflat(1.0, 'fixpower')
