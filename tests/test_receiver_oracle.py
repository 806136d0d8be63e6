"""The coherent receiver's front-end (receiver_cohmix.m, myfilter.m, evaldelay.m), pinned on the interpreted reference
files (oracle/make_golden.py rx -> tests/golden/rx/): the numpy restatement oracle/receiver_oracle.py, and the host
logic of polmux_b200/receiver.py (filter responses, delays, channel position, Hermitian split of the low-pass filter)."""
import glob
import json
import os

import numpy as np
import pytest

import oracle.fiber_oracle as orc
import oracle.receiver_oracle as rxo
from polmux_b200 import receiver as rx
from polmux_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'rx')
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, 'rx_*.npz')))
FILTERS = ('movavg', 'gauss', 'gauss_off', 'butt2', 'butt4', 'butt6', 'ideal', 'bessel5', 'rc1', 'rc2', 'supergauss')


def load(name):
    z = np.load(os.path.join(GOLD, name + '.npz'))
    m = json.loads(str(z['meta']))
    x = dict(m['x'])
    if x.get('lophasenoise') == 'PN':
        x['lophasenoise'] = z['lophasenoise']
    return z, m, x


def oracle_state(z, m):
    gs = orc.reset_all(m['nsymb'], m['nt'], m['nch'])
    gs.SYMBOLRATE, gs.POWER, gs.LAMBDA = m['rate'], z['POWER'].ravel(), synth.wdm_lambdas(m['nch'])   # (create_field set POWER)
    gs.FIELDX, gs.FIELDY = z['FIELDX'], (z['FIELDY'] if z['FIELDY'].size else None)
    gs.FIELDX_TX, gs.FIELDY_TX = z['FIELDX_TX'], (z['FIELDY_TX'] if z['FIELDY_TX'].size else None)
    return gs


def test_case_list():
    assert len(CASES) == 5


@pytest.mark.parametrize('ft', FILTERS)
def test_myfilter_and_evaldelay_match_reference_source(ft):
    z = np.load(os.path.join(GOLD, 'myfilter_all.npz'))
    bw, od = z['P_' + ft]
    for impl in (rxo, rx):
        h = impl.myfilter(ft, z['f'], bw, od)
        np.testing.assert_allclose(h, z['H_' + ft], rtol=0, atol=1e-15)
        assert abs(impl.evaldelay(ft, bw) - float(z['D_' + ft][0])) <= 1e-16
    with pytest.raises(ValueError, match='does not exist'):
        rx.myfilter('elliptic', z['f'], bw)


@pytest.mark.parametrize('name', CASES)
def test_oracle_receiver_matches_reference_source(name):
    z, m, x = load(name)
    iric, xo = rxo.receiver_cohmix(oracle_state(z, m), m['ich'], x)
    assert iric.shape == z['Iric'].shape
    assert np.linalg.norm(iric - z['Iric']) <= 1e-13 * np.linalg.norm(z['Iric'])
    assert abs(xo['avgebx'] / float(z['avgebx'][0]) - 1) < 1e-13
    if z['avgeby'].size:
        assert abs(xo['avgeby'] / float(z['avgeby'][0]) - 1) < 1e-13
    assert abs(xo['post_delay'] - float(z['post_delay'][0])) <= 1e-13 * max(1.0, abs(float(z['post_delay'][0])))


@pytest.mark.parametrize('name', CASES)
def test_host_setup_of_the_device_receiver(name):
    """CohmixSetup (the parameter arithmetic in front of the kernels): the channel's bin shift and band, post_delay, and
    the optical response -- checked through the oracle: filtering with its Hf must give the oracle's filtered spectrum"""
    z, m, x = load(name)
    gs = oracle_state(z, m)

    class G:   # the few GSTATE members CohmixSetup reads
        FN, LAMBDA, SYMBOLRATE, NSYMB, NT, NCH, POWER = gs.FN, gs.LAMBDA, gs.SYMBOLRATE, gs.NSYMB, gs.NT, gs.NCH, gs.POWER
    S = rx.CohmixSetup(m['ich'], x, G, nfc=z['FIELDX'].shape[1])
    assert abs(S.x['post_delay'] - float(z['post_delay'][0])) <= 1e-13 * max(1.0, abs(float(z['post_delay'][0])))
    assert S.balanced == (x.get('pdtype') != 'normal') and S.b2b == ('b2b' in x)
    # the device path in numpy: modulate, filter, mix, filter with the Hermitian part (real currents on one transform)
    n = np.arange(S.nfft)
    src = (z['FIELDX_TX'], z['FIELDY_TX']) if S.b2b else (z['FIELDX'], z['FIELDY'])
    cols = []
    for f in src:
        if not f.size:
            continue
        s = f[:, S.nch - 1] * np.exp(2j * np.pi * ((S.ndfn * n) % S.nfft) / S.nfft)
        s = np.fft.ifft(np.fft.fft(s) * S.hf_opt)
        lo = S.ecw * np.exp(1j * (S.detune * (n + 1) + (S.lophase if S.lophase is not None else 0.0)))
        e = [1j * s + 1j * lo, s - lo, 1j * s - lo, -s + 1j * lo]
        i = [np.abs(v) ** 2 for v in e]
        zc = (i[0] - i[1]) + 1j * (i[2] - i[3]) if S.balanced else i[0] + 1j * i[2]
        zc = np.fft.ifft(np.fft.fft(zc) * rx.hermitian_part(S.hf_el))
        cols += [zc.real, zc.imag]
    got = np.stack(cols, axis=1)
    assert np.linalg.norm(got - z['Iric']) <= 1e-12 * np.linalg.norm(z['Iric'])


def test_hermitian_part_filters_real_sequences_like_real_of_ifft():
    g = np.random.Generator(np.random.PCG64(5))
    h = g.standard_normal(256) + 1j * g.standard_normal(256)          # no symmetry at all
    a, b = g.standard_normal(256), g.standard_normal(256)
    ref_a = np.real(np.fft.ifft(np.fft.fft(a) * h))
    ref_b = np.real(np.fft.ifft(np.fft.fft(b) * h))
    zc = np.fft.ifft(np.fft.fft(a + 1j * b) * rx.hermitian_part(h))
    np.testing.assert_allclose(zc.real, ref_a, atol=1e-13)
    np.testing.assert_allclose(zc.imag, ref_b, atol=1e-13)
