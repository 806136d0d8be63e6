"""-m gpu: the blind DSP core on the device (pmx_dsp_count) against the DSP oracle (oracle/dsp_oracle.py, itself pinned
on the reference's M source by tests/test_dsp_oracle.py): a PDM-QPSK field through a fiber with PMD and an amplifier
with ASE, chromatic dispersion compensated, then CMA + Viterbi & Viterbi + differential decision.  Integer error
counts and the number of CMA passes must be equal."""
import math

import numpy as np
import pytest

import oracle.dsp_oracle as dsp_orc
import polmux_b200 as pmx
from polmux_b200 import _lib, dsp, synth
from common import base_fiber

pytestmark = pytest.mark.gpu


def _received_field(nsymb, nt, nf_db, seed, dgd=0.3, length=4e4, batch=2):
    """-> (ctx, DeviceField with `batch` realizations after fiber + amplifier + CD compensation, symbols)"""
    import torch  # noqa: F401  (device buffer for the counts)
    ex, ey, sx, sy = synth.pdm_qpsk(nsymb, nt, 1)
    pmx.reset_all(nsymb, nt, 1)
    G = pmx.GSTATE
    G.SYMBOLRATE, G.LAMBDA, G.POWER = 28.0, np.array([1550.0]), np.array([1.0])
    pmx.create_field('unique', ex, ey, {'power': 'average'})
    ctx = _lib.default_context()
    n = nsymb * nt
    fld = _lib.DeviceField(ctx, n, 1, batch)
    xs, ys = [], []
    for b in range(batch):
        G.POWER = np.array([1.0])          # (create_field rewrites POWER: every realization starts from the same one)
        pmx.create_field('unique', ex, ey, {'power': 'average'})
        G.DELAY, G.DISP = np.zeros((2, 1)), np.zeros((2, 1))
        pmx.fiber(base_fiber(length=length, dgd=dgd, nplates=20, manakov='yes'), 'gps-',
                  rng=np.random.Generator(np.random.PCG64(seed + b)))
        pmx.ampliflat(0.2 * length * 1e-3, 'gain', {'f': nf_db}, seed=seed + 100 + b)
        pmx.fiber(base_fiber(length=length, disp=-17.0, alphadB=0.0), 'g---')          # dispersion compensation
        xs.append(np.ascontiguousarray(np.array(G.FIELDX).T))
        ys.append(np.ascontiguousarray(np.array(G.FIELDY).T))
    fld.upload(np.stack(xs), np.stack(ys))
    return ctx, fld, np.stack(xs)[:, 0, :], np.stack(ys)[:, 0, :], sx[:, 0], sy[:, 0]


@pytest.mark.parametrize('nf_db,expect_errors', [(5.0, False), (33.0, True)])
def test_dsp_count_matches_the_oracle(nf_db, expect_errors):
    import torch
    nsymb, nt, batch = 1 << 11, 16, 2
    ctx, fld, hx, hy, sx, sy = _received_field(nsymb, nt, nf_db, seed=50, batch=batch)
    params = dict(taps=7, mu=1 / 2000, freqavg=200, phasavg=3, poworder=2)
    ref = dsp.reference_pattern(sx, sy)
    counts = torch.zeros(batch, dtype=torch.int64, device='cuda')
    passes = dsp.dsp_count(ctx, fld, nsymb, nt, ref, counts.data_ptr(), **params)
    got = counts.cpu().numpy()
    tx_phase = np.stack([np.angle((2.0 * (s & 1) - 1) + 1j * (2.0 * ((s >> 1) & 1) - 1)) for s in (sx, sy)], axis=1)
    for b in range(batch):
        s = np.stack([hx[b, ::nt], hy[b, ::nt]], axis=1)
        s = s / math.sqrt(np.mean(np.abs(s) ** 2))
        y, npass = dsp_orc.cma_polar_demux(s, mu=params['mu'], taps=params['taps'])
        ph = dsp_orc.carrier_recovery(y, 2, params['freqavg'], params['phasavg'], params['poworder'])
        want = dsp_orc.count_errors_dqpsk(ph, tx_phase)
        assert int(passes[b]) == npass and npass >= 1
        assert int(got[b]) == want
        assert (want > 0) == expect_errors
    # the product's reference pattern is the oracle's decoding of the transmitted phases
    o = dsp_orc.samp2pat_coherent(tx_phase)
    np.testing.assert_array_equal(ref[:, 0:2], dsp_orc.pat_decoder_dqpsk_binary(o[:, 0:2])[1])
    np.testing.assert_array_equal(ref[:, 2:4], dsp_orc.pat_decoder_dqpsk_binary(o[:, 2:4])[1])
    fld.close()


def test_dsp_count_without_polarization_demultiplexer_and_argument_checks():
    import torch
    nsymb, nt = 1 << 10, 16
    ctx, fld, hx, hy, sx, sy = _received_field(nsymb, nt, 5.0, seed=60, dgd=0.0, batch=1)
    counts = torch.zeros(1, dtype=torch.int64, device='cuda')
    ref = dsp.reference_pattern(sx, sy)
    dsp.dsp_count(ctx, fld, nsymb, nt, ref, counts.data_ptr(), applypol=False, freqavg=100)
    s = np.stack([hx[0, ::nt], hy[0, ::nt]], axis=1)
    s = s / math.sqrt(np.mean(np.abs(s) ** 2))
    tx_phase = np.stack([np.angle((2.0 * (q & 1) - 1) + 1j * (2.0 * ((q >> 1) & 1) - 1)) for q in (sx, sy)], axis=1)
    assert int(counts[0]) == dsp_orc.count_errors_dqpsk(dsp_orc.carrier_recovery(s, 2, 100, 3, 2), tx_phase)
    with pytest.raises(_lib.PolmuxError):
        dsp.dsp_count(ctx, fld, nsymb, nt, ref, counts.data_ptr(), taps=4)
    fld.close()


@pytest.mark.parametrize('nf_db', [5.0, 38.0])
def test_receiver_front_end_plus_dsp_chain_matches_the_oracles(nf_db):
    """the whole receive chain of McRunner(receiver='cohmix') on the device -- optical filter, LO (detuned, with phase
    noise), photodiodes, low-pass filter, sampling at the delayed symbol centres, /peak, CMA, Viterbi & Viterbi, decision --
    against oracle/receiver_oracle.py followed by oracle/dsp_oracle.py on the same received fields: equal error counts"""
    import torch
    import oracle.receiver_oracle as rxo
    from polmux_b200 import receiver as rx
    nsymb, nt, batch = 1 << 11, 16, 2
    ctx, fld, hx, hy, sx, sy = _received_field(nsymb, nt, nf_db, seed=70, batch=batch)
    n = nsymb * nt
    G = pmx.GSTATE
    pn = np.cumsum(2e-3 * np.random.Generator(np.random.PCG64(4)).standard_normal(n))
    x = {'oftype': 'gauss', 'obw': 1.9, 'eftype': 'bessel5', 'ebw': 0.65, 'lopower': 0.0, 'lodetuning': 4.0e7,
         'lophasenoise': pn - np.arange(n) / (n - 1) * pn[-1]}
    S = rx.CohmixSetup(1, x, G, nfc=1)
    shift = int(round((rx.evaldelay('gauss', 0.95) + rx.evaldelay('bessel5', 0.65)) * nt))
    peak = 4.0 * math.sqrt(float(G.POWER[0]))
    assert shift == round(0.3863 / 0.65 * nt)
    fo = _lib.Filter(ctx, n, 1, S.hf_opt, batch=batch)
    fe = _lib.Filter(ctx, n, 1, rx.hermitian_part(S.hf_el), batch=batch)
    fo.execute(fld)
    _lib.cohmix_exec(ctx, fld, S.ecw, S.detune, S.lophase, S.balanced)
    fe.execute(fld)
    params = dict(taps=7, mu=1 / 2000, freqavg=200, phasavg=3, poworder=2)
    ref = dsp.reference_pattern(sx, sy)
    counts = torch.zeros(batch, dtype=torch.int64, device='cuda')
    passes = dsp.dsp_count(ctx, fld, nsymb, nt, ref, counts.data_ptr(), sample_shift=shift, peak=peak, **params)
    got = counts.cpu().numpy()
    tx_phase = np.stack([np.angle((2.0 * (s & 1) - 1) + 1j * (2.0 * ((s >> 1) & 1) - 1)) for s in (sx, sy)], axis=1)

    class GS:
        pass
    for b in range(batch):
        gs = GS()
        for k in ('FN', 'LAMBDA', 'SYMBOLRATE', 'NSYMB', 'NT', 'NCH', 'POWER'):
            setattr(gs, k, getattr(G, k))
        gs.FIELDX, gs.FIELDY = hx[b][:, None], hy[b][:, None]
        iric, _ = rxo.receiver_cohmix(gs, 1, x)
        idx = (np.arange(nsymb) * nt + shift) % n
        s = np.stack([iric[idx, 0] + 1j * iric[idx, 1], iric[idx, 2] + 1j * iric[idx, 3]], axis=1) / peak
        y, npass = dsp_orc.cma_polar_demux(s, mu=params['mu'], taps=params['taps'])
        ph = dsp_orc.carrier_recovery(y, 2, params['freqavg'], params['phasavg'], params['poworder'])
        want = dsp_orc.count_errors_dqpsk(ph, tx_phase)
        assert int(passes[b]) == npass
        assert int(got[b]) == want
        assert (want > 0) == (nf_db > 20)
    fo.close()
    fe.close()
    fld.close()


def test_dsp4cohdec_returns_the_oracles_phases_and_amplitudes():
    """polmux_b200.dsp.dsp4cohdec(ich, pat, x, p) -- the reference's call (ex20_coherent_polmux.m:151), x.delay = 'theory':
    Phases / Amplitudes equal those of oracle/receiver_oracle.py + oracle/dsp_oracle.py (<= 1e-9 rad away from the +-pi
    cut), and samp2pat / pat_decoder on them give the counts the device counter gives"""
    import oracle.receiver_oracle as rxo
    nsymb, nt = 1 << 11, 16
    n = nsymb * nt
    ex, ey, sx, sy = synth.pdm_qpsk(nsymb, nt, 1)
    pmx.reset_all(nsymb, nt, 1)
    G = pmx.GSTATE
    G.SYMBOLRATE, G.LAMBDA, G.POWER = 28.0, np.array([1550.0]), np.array([1.0])
    pmx.create_field('unique', ex, ey, {'power': 'average'})
    pmx.fiber(base_fiber(length=4e4, dgd=0.3, nplates=20, manakov='yes', disp=0.0), 'gps-',
              rng=np.random.Generator(np.random.PCG64(81)))
    pmx.ampliflat(8.0, 'gain', {'f': 30.0}, seed=5)
    x = {'rec': 'coherent', 'oftype': 'gauss', 'obw': 1.9, 'eftype': 'bessel5', 'ebw': 0.65, 'lopower': 0.0, 'delay': 'theory'}
    p = {'sps': nt, 'workatbaudrate': True, 'applyadc': False, 'applydcf': False, 'applynlr': False, 'applypol': True,
         'polmethod': 'cma', 'cmaparams': {'R': (1.0, 1.0), 'mu': 1 / 2000, 'taps': 7, 'txpolars': 2, 'phizero': 0.0},
         'modorder': 2, 'freqavg': 200, 'phasavg': 3, 'poworder': 2}
    pat = np.stack([sx[:, 0], sy[:, 0]], axis=1)
    hx, hy = np.array(G.FIELDX), np.array(G.FIELDY)
    delay0 = float(0.5 * (G.DELAY[0, 0] + G.DELAY[1, 0]))
    phases, amps = dsp.dsp4cohdec(1, pat, x, p, decimator='sample')
    assert phases.shape == (nsymb, 2) and amps.shape == (nsymb, 2)

    class GS:
        pass
    gs = GS()
    for k in ('FN', 'LAMBDA', 'SYMBOLRATE', 'NSYMB', 'NT', 'NCH', 'POWER'):
        setattr(gs, k, getattr(G, k))
    gs.FIELDX, gs.FIELDY = hx, hy
    iric, xo = rxo.receiver_cohmix(gs, 1, x)
    shift = int(round((delay0 + rxo.evaldelay('gauss', 0.95) + rxo.evaldelay('bessel5', 0.65) + xo['post_delay']) * nt))
    idx = (np.arange(nsymb) * nt + shift) % n
    s = np.stack([iric[idx, 0] + 1j * iric[idx, 1], iric[idx, 2] + 1j * iric[idx, 3]], axis=1) / (4 * math.sqrt(float(G.POWER[0])))
    y, _ = dsp_orc.cma_polar_demux(s, mu=1 / 2000, taps=7)
    want = dsp_orc.carrier_recovery(y, 2, 200, 3, 2)
    np.testing.assert_allclose(amps, np.abs(y), rtol=1e-9, atol=1e-12)
    away = np.abs(np.abs(want) - math.pi) > 1e-6
    np.testing.assert_allclose(phases[away], want[away], rtol=0, atol=1e-9)
    tx_phase = np.stack([np.angle((2.0 * (q & 1) - 1) + 1j * (2.0 * ((q >> 1) & 1) - 1)) for q in (sx[:, 0], sy[:, 0])], axis=1)
    assert dsp_orc.count_errors_dqpsk(phases, tx_phase) == dsp_orc.count_errors_dqpsk(want, tx_phase)
    with pytest.raises(NotImplementedError, match='theory'):
        dsp.dsp4cohdec(1, pat, dict(x, delay='estimate'), p)
    with pytest.raises(NotImplementedError, match='applydcf'):
        dsp.dsp4cohdec(1, pat, x, dict(p, applydcf=True, workatbaudrate=False))
    # p.applyadc (5 bits) and p.applynlr: the same chain with the oracle's ADC on the currents and NLRotation on the sampled,
    # not yet normalised signals
    p2 = dict(p, applyadc=True, adcbits=5, applynlr=True, nlralpha=0.02)
    ph2, am2 = dsp.dsp4cohdec(1, pat, x, p2, decimator='sample')
    iq = dsp_orc.adc_quantize(iric, 5)
    s2 = np.stack([iq[idx, 0] + 1j * iq[idx, 1], iq[idx, 2] + 1j * iq[idx, 3]], axis=1)
    s2 = dsp_orc.nl_rotation(s2, 0.02) / (4 * math.sqrt(float(G.POWER[0])))
    y2, _ = dsp_orc.cma_polar_demux(s2, mu=1 / 2000, taps=7)
    want2 = dsp_orc.carrier_recovery(y2, 2, 200, 3, 2)
    np.testing.assert_allclose(am2, np.abs(y2), rtol=1e-9, atol=1e-12)
    away = np.abs(np.abs(want2) - math.pi) > 1e-6
    np.testing.assert_allclose(ph2[away], want2[away], rtol=0, atol=1e-9)
    assert not np.allclose(am2, amps)
    # p.applydcf at one sample per symbol: the truncated-FIR dispersion compensation of DispCompFilter on the sampled signals
    p3 = dict(p, applydcf=True, dispersion=300.0, ndispsym=16, baudrate=28e9)
    p3['lambda'] = 1550.0
    ph3, am3 = dsp.dsp4cohdec(1, pat, x, p3, decimator='sample')
    s3 = np.stack([iric[idx, 0] + 1j * iric[idx, 1], iric[idx, 2] + 1j * iric[idx, 3]], axis=1)
    s3 = dsp_orc.apply_dcf(s3, 300.0, 1550.0, 28e9, 16) / (4 * math.sqrt(float(G.POWER[0])))
    y3, _ = dsp_orc.cma_polar_demux(s3, mu=1 / 2000, taps=7)
    want3 = dsp_orc.carrier_recovery(y3, 2, 200, 3, 2)
    np.testing.assert_allclose(am3, np.abs(y3), rtol=1e-8, atol=1e-11)
    away = np.abs(np.abs(want3) - math.pi) > 1e-6
    np.testing.assert_allclose(ph3[away], want3[away], rtol=0, atol=1e-8)
    assert not np.allclose(am3, amps)
    # the default: the decimator's anti-alias FIR (fir1(16, 1/NT), restated from the toolbox's published description)
    # in front of the sampling
    ph4, am4 = dsp.dsp4cohdec(1, pat, x, p)
    dec = dsp_orc.decimate_fir(np.roll(iric, -shift, axis=0), nt)
    s4 = np.stack([dec[:, 0] + 1j * dec[:, 1], dec[:, 2] + 1j * dec[:, 3]], axis=1) / (4 * math.sqrt(float(G.POWER[0])))
    y4, _ = dsp_orc.cma_polar_demux(s4, mu=1 / 2000, taps=7)
    want4 = dsp_orc.carrier_recovery(y4, 2, 200, 3, 2)
    np.testing.assert_allclose(am4, np.abs(y4), rtol=1e-9, atol=1e-12)
    away = np.abs(np.abs(want4) - math.pi) > 1e-6
    np.testing.assert_allclose(ph4[away], want4[away], rtol=0, atol=1e-9)
    assert not np.allclose(am4, amps)
    assert dsp_orc.count_errors_dqpsk(ph4, tx_phase) <= dsp_orc.count_errors_dqpsk(phases, tx_phase)   # less noise behind the FIR


@pytest.mark.parametrize('method', ['easi', 'combo'])
def test_easi_and_combo_demultiplexers_match_the_oracle(method):
    """p.polmethod = 'easi' / 'combo' (dsp4cohdec.m:234-241): EASI source separation (easipolardemux around
    easiadaptivefilter.m) alone or in front of the CMA, on the device against oracle/dsp_oracle.py: equal counts and pass
    counts of both stages"""
    import torch
    nsymb, nt, batch = 1 << 11, 16, 2
    ctx, fld, hx, hy, sx, sy = _received_field(nsymb, nt, 26.0, seed=90, dgd=0.5, batch=batch)
    params = dict(taps=7, mu=1 / 2000, freqavg=200, phasavg=3, poworder=2, applyeasi=True, easi_mu=1 / 1500, easi_phizero=0.05,
                  applypol=(method == 'combo'))
    ref = dsp.reference_pattern(sx, sy)
    counts = torch.zeros(batch, dtype=torch.int64, device='cuda')
    ep = np.zeros(batch, dtype=np.int32)
    passes = dsp.dsp_count(ctx, fld, nsymb, nt, ref, counts.data_ptr(), easi_passes_out=ep, **params)
    got = counts.cpu().numpy()
    tx_phase = np.stack([np.angle((2.0 * (s & 1) - 1) + 1j * (2.0 * ((s >> 1) & 1) - 1)) for s in (sx, sy)], axis=1)
    for b in range(batch):
        s = np.stack([hx[b, ::nt], hy[b, ::nt]], axis=1)
        s = s / math.sqrt(np.mean(np.abs(s) ** 2))
        y, ne = dsp_orc.easi_polar_demux(s, mu=1 / 1500, phizero=0.05)
        nc = 0
        if method == 'combo':
            y, nc = dsp_orc.cma_polar_demux(y, mu=params['mu'], taps=params['taps'])
        ph = dsp_orc.carrier_recovery(y, 2, params['freqavg'], params['phasavg'], params['poworder'])
        want = dsp_orc.count_errors_dqpsk(ph, tx_phase)
        assert int(ep[b]) == ne and ne >= 1
        assert int(passes[b]) == nc
        assert int(got[b]) == want
    fld.close()
