"""Blind DSP core of the reference's coherent receiver on the device (pmx_dsp_count): constant-modulus polarization
demultiplexer (cmapolardemux / cmaadaptivefilter, dsp4cohdec.m:353-427, cmaadaptivefilter.m:33-55), Viterbi & Viterbi
carrier frequency and phase (dsp4cohdec.m:241-283, 320-345), decision (samp2pat.m:60-67), differential decoding
(pat_decoder.m:68-82), X/Y swap test and error count (ex20_coherent_polmux.m:168-176).  The parameters are the fields
of the reference's dspParameters struct (ex20_coherent_polmux.m:62-91)."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib

DEFAULTS = dict(applypol=True, taps=7, mu=1 / 6000, R=(1.0, 1.0), phizero=0.0, max_passes=0, modorder=2, freqavg=500,
                phasavg=3, poworder=2, sample_shift=0, peak=0.0, applyeasi=False, easi_mu=1 / 6000, easi_phizero=0.0,
                easi_max_passes=0, nlr_alpha=0.0, dcf_h=None, decim_taps=None)


def reference_pattern(sym_x, sym_y):
    """The pattern the received one is counted against: the transmitted QPSK symbols (indices 0..3, bit 0 -> sign of the
    real part, bit 1 -> sign of the imaginary part, polmux_b200.synth) put through the receiver's own decision rule
    (samp2pat 'coherent') and differential decoder (pat_decoder 'dqpsk', binary).  -> uint8 [nsymb, 4]"""
    out = []
    for s in (np.asarray(sym_x).ravel(), np.asarray(sym_y).ravel()):
        ph = np.angle((2.0 * (s & 1) - 1) + 1j * (2.0 * ((s >> 1) & 1) - 1))
        first, second = (np.abs(ph) <= math.pi / 2).astype(np.int64), (ph > 0).astype(np.int64)
        quad = np.where(first == 0, np.where(second == 0, 0, 1), np.where(second == 1, 2, 3))     # pat2stars (binary)
        d = (np.roll(quad, 1) - quad) & 3                                                         # conj(s).*shift(s,1)
        m0, m1 = ((d == 2) | (d == 3)).astype(np.uint8), ((d == 1) | (d == 2)).astype(np.uint8)   # stars2pat
        out += [1 - m0, 1 - m1]                                                                   # patmat = ~patmat
    return np.ascontiguousarray(np.stack(out, axis=1), dtype=np.uint8)


def fir1_lowpass(order, wn):
    """The low-pass design decimate(x, r, order, 'fir') asks fir1 for, fir1(order, 1/r) -- Signal Processing Toolbox functions
    that are NOT in the reference tree (dsp4cohdec.m:176-184 calls decimate).  Restated from the toolbox's published
    description: ideal low-pass impulse response wn*sinc(wn*(n - order/2)), Hamming window 0.54 - 0.46*cos(2*pi*n/order),
    scaled to unit gain at DC; decimate filters in one direction and takes every r-th sample from the filter's group delay
    on, i.e. a zero-phase FIR read at the sampling instants.  PARITY UNPINNED: no toolbox source and no golden vector
    exist here; decimate's mirror extension at the two ends of the record is replaced by the circular one of the rest
    of the chain."""
    n = np.arange(order + 1)
    h = wn * np.sinc(wn * (n - order / 2.0)) * (0.54 - 0.46 * np.cos(2 * math.pi * n / order))
    return h / h.sum()


def disp_comp_filter(beta2l, bw, n, flen):
    """Hfilt = DispCompFilter(Beta2L, B, N, FilterLength), dsp4cohdec.m:289-297: the all-pass response of the dispersion to
    undo, its impulse response truncated to FilterLength + 1 taps, the filter's delay taken out again.  Host set-up (a
    filter design: two length-N transforms, once per receiver), like the evaluation of myfilter."""
    freq = np.fft.ifftshift(-bw / 2 + np.arange(n) * (bw / n))
    delay = 2 * math.pi * freq / bw * (flen / 2)
    argum = (2 * math.pi * freq) ** 2 * beta2l / 2 - delay
    b = np.fft.ifft(np.cos(argum) + 1j * np.sin(argum))[:int(flen) + 1]
    return np.fft.fft(b, n) * (np.cos(delay) + 1j * np.sin(delay))


def _desc(nsymb, nt, params, easi_passes=None):
    p = dict(DEFAULTS)
    p.update(params)
    d = _lib.DspDesc()
    d.nsymb, d.nt, d.apply_cma, d.taps, d.mu = int(nsymb), int(nt), int(bool(p['applypol'])), int(p['taps']), float(p['mu'])
    d.R[0], d.R[1], d.phizero, d.max_passes = float(p['R'][0]), float(p['R'][1]), float(p['phizero']), int(p['max_passes'])
    d.modorder, d.freqavg, d.phasavg, d.poworder = int(p['modorder']), int(p['freqavg']), int(p['phasavg']), int(p['poworder'])
    d.sample_shift, d.peak = int(p['sample_shift']), float(p['peak'])
    d.apply_easi, d.easi_mu, d.easi_phizero = int(bool(p['applyeasi'])), float(p['easi_mu']), float(p['easi_phizero'])
    d.easi_max_passes = int(p['easi_max_passes'])
    d.nlr_alpha = float(p['nlr_alpha'])
    if p['dcf_h'] is not None:
        h = np.ascontiguousarray(np.asarray(p['dcf_h'], dtype=np.complex128).ravel())
        if h.size != int(nsymb):
            raise ValueError('dcf_h: one response value per symbol')
        d._keep_dcf = h                                   # (the descriptor keeps the array alive)
        d.dcf_h = h.ctypes.data_as(_lib._dp)
    if p['decim_taps'] is not None:
        t = np.ascontiguousarray(np.asarray(p['decim_taps'], dtype=np.float64).ravel())
        d._keep_taps = t
        d.decim_ntaps, d.decim_taps = int(t.size), t.ctypes.data_as(_lib._dp)
    if easi_passes is not None:
        d.easi_passes = easi_passes.ctypes.data_as(C.POINTER(C.c_int32))
    return d


def dsp_phases(ctx: _lib.Context, field: _lib.DeviceField, nsymb: int, nt: int, want_amplitudes=True, easi_passes_out=None,
               **params):
    """-> (Phases, Amplitudes, passes): [batch, nsymb, 2] each, what dsp4cohdec returns (pmx_dsp_phases)."""
    d = _desc(nsymb, nt, params, easi_passes_out)
    ph = np.zeros((field.batch, 2, nsymb), dtype=np.float64)
    am = np.zeros_like(ph) if want_amplitudes else None
    passes = np.zeros(field.batch, dtype=np.int32)
    ctx.check(ctx.lib.pmx_dsp_phases(ctx.h, field.h, C.byref(d), ph.ctypes.data_as(_lib._dp),
                                     None if am is None else am.ctypes.data_as(_lib._dp),
                                     passes.ctypes.data_as(C.POINTER(C.c_int32))))
    tr = lambda a: None if a is None else np.ascontiguousarray(a.transpose(0, 2, 1))
    return tr(ph), tr(am), passes


def _dcf_response(p, G):
    """dsp4cohdec.m:198-207 at one sample per symbol (p.workatbaudrate)"""
    from .gstate import CONSTANTS
    beta2l = -p['dispersion'] * p['lambda'] ** 2 / 2 / math.pi / CONSTANTS.CLIGHT * 1e-21
    return disp_comp_filter(beta2l, float(p['baudrate']), int(G.NSYMB), int(p['ndispsym']))


def dsp4cohdec(ich, pat, x, p, ctx=None, decimator='fir'):
    """[Phases, Amplitudes] = dsp4cohdec(ich, pat, x, p) -- dsp4cohdec.m:1, for a two-polarization QPSK field, on the device:
    receiver_cohmix (polmux_b200/receiver.py), the shift by the receiver's delay (:167-169), one sample per symbol,
    normalisation by 4*sqrt(POWER(ich)) (:226-227), polarization demultiplexer (p.applypol, 'cma'), carrier recovery
    (p.freqavg / phasavg / poworder).  -> (Phases [nsymb, 2], Amplitudes [nsymb, 2]).
    Differences stated in DESIGN.md: x.delay must be 'theory' (the pattern-correlation search of mygeteyeinfo is not
    built, so `pat` only tells the number of polarizations and worsteyeop is not returned); the decimator (`decimate`, a
    Signal Processing Toolbox function outside the reference tree) is restated from its published description
    (decimator='fir': fir1_lowpass(16, 1/NT) read at the symbol centres, parity unpinned) or left out (decimator='sample');
    the 'singlepol' demultiplexer raises; p.applyadc, p.applydcf (with p.workatbaudrate), p.applynlr, 'cma', 'easi' and 'combo' are built."""
    from . import receiver as _rx
    from .gstate import GSTATE as G
    if x.get('rec', 'coherent') != 'coherent':
        raise ValueError("Flag X.rec must be 'coherent'")                       # dsp4cohdec.m:143
    if decimator not in ('fir', 'sample'):
        raise ValueError("decimator must be 'fir' or 'sample'")
    if p.get('applydcf') and not p.get('workatbaudrate'):
        raise NotImplementedError('dsp4cohdec: p.applydcf at two samples per symbol needs the decimator (not built); '
                                  'set p.workatbaudrate, or compensate with x.dpost')
    method = str(p.get('polmethod', 'cma')).lower()
    if p.get('applypol') and method not in ('cma', 'easi', 'combo'):
        if method == 'singlepol':
            raise NotImplementedError("dsp4cohdec: polarization demultiplexer 'singlepol' is not built")
        raise ValueError('Unknown Polar Rotation method.')                      # dsp4cohdec.m:242-243
    if x.get('delay') != 'theory':
        raise NotImplementedError("dsp4cohdec: x.delay = 'theory' only (the pattern-correlation timing search is not built)")
    if not G.has_y() or np.ndim(pat) < 2 or np.shape(pat)[1] == 1:
        raise NotImplementedError('dsp4cohdec: two-polarization fields and patterns only')
    if int(p.get('modorder', 2)) != 2:
        raise NotImplementedError('dsp4cohdec: QPSK (modorder 2) only')
    ctx = ctx or _lib.default_context()
    col, S, xo = _rx.front_end(ich, x, ctx)
    try:
        if p.get('applyadc'):                                                   # dsp4cohdec.m:157-162
            _lib.field_quantize(ctx, col, int(p['adcbits']))
        if 'b2b' in x:
            avgdelay = 0.0                                                      # mygeteyeinfo, dsp4cohdec.m:491-493
        else:
            dl = np.asarray(G.DELAY, dtype=np.float64)
            avgdelay = 0.5 * (dl[0, ich - 1] + dl[1, ich - 1])
        delay = avgdelay + _rx.evaldelay(xo['oftype'], xo['obw'] * 0.5) + _rx.evaldelay(xo['eftype'], xo['ebw']) + xo['post_delay']
        cma = dict(p.get('cmaparams') or {})
        easi = dict(p.get('easiparams') or {})
        on = bool(p.get('applypol'))
        params = dict(applypol=on and method in ('cma', 'combo'), applyeasi=on and method in ('easi', 'combo'),
                      easi_mu=float(easi.get('mu', 1 / 6000)), easi_phizero=float(easi.get('phizero', 0.0)),
                      taps=int(cma.get('taps', 7)), mu=float(cma.get('mu', 1 / 6000)),
                      R=tuple(cma.get('R', (1.0, 1.0))), phizero=float(cma.get('phizero', 0.0)),
                      modorder=2, freqavg=int(p.get('freqavg', 500)), phasavg=int(p.get('phasavg', 3)),
                      poworder=int(p.get('poworder', 2)), sample_shift=int(round(delay * G.NT)),
                      nlr_alpha=float(p['nlralpha']) if p.get('applynlr') else 0.0,
                      dcf_h=_dcf_response(p, G) if p.get('applydcf') else None,
                      decim_taps=fir1_lowpass(16, 1.0 / G.NT) if decimator == 'fir' else None,
                      peak=4.0 * math.sqrt(float(np.asarray(G.POWER).ravel()[ich - 1])))
        ph, am, _ = dsp_phases(ctx, col, G.NSYMB, G.NT, **params)
    finally:
        col.close()
    return ph[0], am[0]


def dsp_count(ctx: _lib.Context, field: _lib.DeviceField, nsymb: int, nt: int, ref_patmat, counts_dev_ptr: int,
              easi_passes_out=None, **params):
    """Error counts of every realization of `field` (CD already compensated) into a device int64 buffer.  sample_shift /
    peak: the receiver's currents are sampled at k*nt + sample_shift and divided by peak (0: unit mean power).
    applyeasi / easi_mu / easi_phizero: EASI before (polmethod 'combo') or instead of (applypol=False: 'easi') the CMA;
    easi_passes_out: int32 [batch] array that receives the passes of that stage.
    -> passes the CMA polarization demultiplexer ran, per realization."""
    d = _desc(nsymb, nt, params, easi_passes_out)
    ref = np.ascontiguousarray(ref_patmat, dtype=np.uint8).reshape(nsymb, 4)
    passes = np.zeros(field.batch, dtype=np.int32)
    ctx.check(ctx.lib.pmx_dsp_count(ctx.h, field.h, C.byref(d), ref.ctypes.data_as(C.POINTER(C.c_uint8)),
                                    C.c_void_p(int(counts_dev_ptr)), passes.ctypes.data_as(C.POINTER(C.c_int32))))
    return passes
