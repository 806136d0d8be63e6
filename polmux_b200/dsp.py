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
                phasavg=3, poworder=2, sample_shift=0, peak=0.0)


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


def dsp_count(ctx: _lib.Context, field: _lib.DeviceField, nsymb: int, nt: int, ref_patmat, counts_dev_ptr: int, **params):
    """Error counts of every realization of `field` (CD already compensated) into a device int64 buffer.  sample_shift /
    peak: the receiver's currents are sampled at k*nt + sample_shift and divided by peak (0: unit mean power).
    -> passes the polarization demultiplexer ran, per realization."""
    p = dict(DEFAULTS)
    p.update(params)
    d = _lib.DspDesc()
    d.nsymb, d.nt, d.apply_cma, d.taps, d.mu = int(nsymb), int(nt), int(bool(p['applypol'])), int(p['taps']), float(p['mu'])
    d.R[0], d.R[1], d.phizero, d.max_passes = float(p['R'][0]), float(p['R'][1]), float(p['phizero']), int(p['max_passes'])
    d.modorder, d.freqavg, d.phasavg, d.poworder = int(p['modorder']), int(p['freqavg']), int(p['phasavg']), int(p['poworder'])
    d.sample_shift, d.peak = int(p['sample_shift']), float(p['peak'])
    ref = np.ascontiguousarray(ref_patmat, dtype=np.uint8).reshape(nsymb, 4)
    passes = np.zeros(field.batch, dtype=np.int32)
    ctx.check(ctx.lib.pmx_dsp_count(ctx.h, field.h, C.byref(d), ref.ctypes.data_as(C.POINTER(C.c_uint8)),
                                    C.c_void_p(int(counts_dev_ptr)), passes.ctypes.data_as(C.POINTER(C.c_int32))))
    return passes
