"""NumPy restatement of the blind DSP core of the reference's coherent receiver (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product never does.

Reference code followed (all under /root/reference):
  cmaadaptivefilter.m:33-55 (C twin cmaadaptivefilter.c:57-91)   constant-modulus 2x2 FIR update, sample by sample
  dsp4cohdec.m:353-427   cmapolardemux: initial taps, circular extension, passes until the taps move < 5e-5
  easiadaptivefilter.m:28-61 (C twin easiadaptivefilter.c)   EASI source separation, one 2x2 tap: Y = H*x,
                         H <- (I - mu*E(Y))*H with the nonlinear error matrix E (errorfun, :55-61)
  dsp4cohdec.m:157-162   ADC emulation: round((Irx + M)/2./M*2^bits)*2.*M/2^bits - M, M = max(max(abs(Irx)))
  dsp4cohdec.m:198-210,289-297   digital dispersion compensation: DispCompFilter (ideal all-pass response, impulse response
                         truncated to FilterLength+1 taps, the delay taken out again) and ifft(fft(Signals).*Hfilt)
  dsp4cohdec.m:308-315   NLRotation: Phases + alpha*(sum(|s|^2,2) - mean), on the sampled signals before the normalisation
  dsp4cohdec.m:428-482   easipolardemux: initial rotation, passes until the taps move < 5e-5 (at most 20*ceil(1/(L*mu)) - 1)
  dsp4cohdec.m:320-345   vitvit: M-th power, circular moving average of 2k+1 samples, (unwrapped) angle / M
  dsp4cohdec.m:241-283   carrier recovery: frequency from s.*conj(shift(s)) (navg = freqavg), phase (navg = phasavg)
  samp2pat.m:60-67       'coherent': first_bit = |phase| <= pi/2, second_bit = phase > 0
  pat_decoder.m:68-82    'dqpsk' with binary patterns: pat2stars (pat2stars.m:51-63), conj(s).*shift(s,1),
                         stars2pat (stars2pat.m:28-41), both bits inverted
  ex20_coherent_polmux.m:168-176   X/Y swap test, ber_estimate.m:118 error count

PARITY PINNING: tests/test_dsp_oracle.py executes the reference's own cmaadaptivefilter.m, easiadaptivefilter.m, samp2pat.m,
pat_decoder.m (+ pat2stars.m, stars2pat.m, fastshift.m) and the sub-functions vitvit / cmapolardemux / easipolardemux of
dsp4cohdec.m with the mini interpreter (oracle/mini_m) on seeded inputs and compares them with this file.

The front-end of receiver_cohmix.m is restated in oracle/receiver_oracle.py.  NOT restated: mygeteyeinfo's timing
search.  PARITY UNPINNED for one function: `decimate_fir` below stands in for the toolbox functions decimate / fir1
(dsp4cohdec.m:176-184; Signal Processing Toolbox, not in the reference tree, no golden vector): it restates their
published description and is checked against the product's own design only.
"""
from __future__ import annotations

import math

import numpy as np


def errorfuncma(x, m):
    """cmaadaptivefilter.m:54-55: E = X.*(M - abs(X).^2)"""
    return x * (m - np.abs(x) ** 2)


def cma_adaptive_filter(xx, h1, h2, mu, R):
    """cmaadaptivefilter.m:33-52.  xx: [L+ntap-1, 2] circularly extended input; h1, h2: [ntap, 2].
    -> (Y [L,2], h1, h2).  Sums in the interpreter's order: columns first, then across the two columns."""
    xx = np.asarray(xx, dtype=np.complex128)
    h1 = np.array(h1, dtype=np.complex128)
    h2 = np.array(h2, dtype=np.complex128)
    ntap = h1.shape[0]
    L = xx.shape[0] - ntap + 1
    y = np.zeros((L, 2), dtype=np.complex128)
    for k in range(L):
        w = xx[k:k + ntap, :]
        y1 = np.sum(np.sum(w * h1, axis=0))
        y2 = np.sum(np.sum(w * h2, axis=0))
        y[k, 0], y[k, 1] = y1, y2
        h1 = h1 + mu * errorfuncma(y1, R[0]) * np.conj(w)
        h2 = h2 + mu * errorfuncma(y2, R[1]) * np.conj(w)
    return y, h1, h2


def cma_polar_demux(x, mu=1 / 6000, taps=7, R=(1.0, 1.0), phizero=0.0, max_passes=None):
    """cmapolardemux (dsp4cohdec.m:353-427) for two transmitted polarizations.  -> (y [L,2], passes run)"""
    x = np.asarray(x, dtype=np.complex128)
    half = taps // 2
    M = np.array([[math.cos(phizero), math.sin(phizero)], [-math.sin(phizero), math.cos(phizero)]], dtype=np.complex128)
    h1 = np.zeros((taps, 2), dtype=np.complex128)
    h2 = np.zeros((taps, 2), dtype=np.complex128)
    h1[half, :] = M[0, :]          # hzero(halftaps+1,:,:) = M; h1 = squeeze(hzero(:,1,:))
    h2[half, :] = M[1, :]
    ext = np.concatenate([x[len(x) - half:], x, x[:half]]) if half else x
    L = len(x)
    repetitions = 50 * math.ceil(1.0 / (L * mu))
    if max_passes is not None:
        repetitions = min(repetitions, max_passes + 1)
    c, conv, y = 1, False, None
    while not conv and c < repetitions:
        h1o, h2o = h1.copy(), h2.copy()
        y, h1n, h2n = cma_adaptive_filter(ext, h1, h2, mu, R)
        if np.any(h1n) or np.any(h2n):
            h1, h2 = h1n, h2n
        if max(np.max(np.abs(h1o - h1)), np.max(np.abs(h2o - h2))) < 5e-5:
            conv = True
        c += 1
    return y, c - 1


def easi_errorfun(a, b, lam):
    """errorfun of easiadaptivefilter.m:55-61 -> E [2,2] (the products are a*b, not a*conj(b), as the reference writes them)"""
    na, nb = abs(a) ** 2, abs(b) ** 2
    d1 = 1 + lam * (na + nb)
    d2 = 1 + lam * (a * abs(a) + b * abs(b))
    e = np.empty((2, 2), dtype=np.complex128)
    e[0, 0] = (na - 1) / d1
    e[0, 1] = (a * b) / d1 + (a * b * (na - nb)) / d2
    e[1, 0] = (a * b) / d1 + (a * b * (nb - na)) / d2
    e[1, 1] = (nb - 1) / d1
    return e


def easi_adaptive_filter(xx, h1, h2, mu):
    """easiadaptivefilter.m:28-51 for one tap: xx [L,2]; h1, h2 [1,2] (rows of the separation matrix).
    -> (Y [L,2], h1, h2)"""
    xx = np.asarray(xx, dtype=np.complex128)
    h1 = np.array(h1, dtype=np.complex128).reshape(1, 2)
    h2 = np.array(h2, dtype=np.complex128).reshape(1, 2)
    L = xx.shape[0]
    y = np.zeros((L, 2), dtype=np.complex128)
    for k in range(L):
        w = xx[k:k + 1, :]
        y1 = np.sum(np.sum(w * h1))
        y2 = np.sum(np.sum(w * h2))
        y[k, 0], y[k, 1] = y1, y2
        e = easi_errorfun(y1, y2, mu)
        h11 = (1 - mu * e[0, 0]) * h1[:, 0] + (-mu * e[0, 1]) * h2[:, 0]
        h12 = (1 - mu * e[0, 0]) * h1[:, 1] + (-mu * e[0, 1]) * h2[:, 1]
        h21 = (-mu * e[1, 0]) * h1[:, 0] + (1 - mu * e[1, 1]) * h2[:, 0]
        h22 = (-mu * e[1, 0]) * h1[:, 1] + (1 - mu * e[1, 1]) * h2[:, 1]
        h1 = np.stack([h11, h12], axis=1)
        h2 = np.stack([h21, h22], axis=1)
    return y, h1, h2


def easi_polar_demux(x, mu=1 / 6000, phizero=0.0, max_passes=None):
    """easipolardemux (dsp4cohdec.m:428-482) for two transmitted polarizations.  -> (y [L,2], passes run)"""
    x = np.asarray(x, dtype=np.complex128)
    M = np.array([[math.cos(phizero), math.sin(phizero)], [-math.sin(phizero), math.cos(phizero)]], dtype=np.complex128)
    h1, h2 = M[0:1, :].copy(), M[1:2, :].copy()
    L = len(x)
    repetitions = 20 * math.ceil(1.0 / (L * mu))
    if max_passes is not None:
        repetitions = min(repetitions, max_passes + 1)
    c, conv, y = 1, False, None
    while not conv and c < repetitions:
        h1o, h2o = h1.copy(), h2.copy()
        y, h1n, h2n = easi_adaptive_filter(x, h1, h2, mu)
        if np.any(h1n) or np.any(h2n):
            h1, h2 = h1n, h2n
        if max(np.max(np.abs(h1o - h1)), np.max(np.abs(h2o - h2))) < 5e-5:
            conv = True
        c += 1
    return y, c - 1


def decimate_fir(x, r, order=16):
    """decimate(x, r, order, 'fir') after the toolbox's published description, PARITY UNPINNED: b = fir1(order, 1/r)
    (Hamming-windowed sinc, unit DC gain), filter(b, 1, x), every r-th sample from the group delay order/2 on; circular
    at the ends (the toolbox mirrors the record there).  x: [N] or [N, ncol] -> [N/r, ...]"""
    n = np.arange(order + 1)
    wn = 1.0 / r
    ideal = np.where(n == order / 2.0, wn, np.sin(math.pi * wn * (n - order / 2.0)) / (math.pi * (n - order / 2.0) + (n == order / 2.0)))
    b = ideal * (0.54 - 0.46 * np.cos(2 * math.pi * n / order))
    b = b / np.sum(b)
    x = np.asarray(x)
    y = np.zeros(x.shape, dtype=np.result_type(x.dtype, np.float64))
    for j in range(order + 1):
        y = y + b[j] * np.roll(x, j - order // 2, axis=0)     # filter(b,1,x)(m + order/2) = sum_j b(j) x(m + order/2 - j)
    return y[::r]


def disp_comp_filter(beta2l, bw, n, flen):
    """DispCompFilter, dsp4cohdec.m:289-297 -> Hfilt [n]"""
    freq = -bw / 2 + np.arange(n) * (bw / n)              # ( -B/2 : B/N : B/2*(N-2)/N )
    freq = np.fft.ifftshift(freq)
    delay = 2 * math.pi * freq / bw * (flen / 2)
    argum = (2 * math.pi * freq) ** 2 * beta2l / 2 - delay
    h = np.cos(argum) + 1j * np.sin(argum)
    b = np.fft.ifft(h)[:int(flen) + 1]
    return np.fft.fft(b, n) * (np.cos(delay) + 1j * np.sin(delay))


def apply_dcf(signals, dispersion, lam, baudrate, ndispsym, workatbaudrate=True):
    """dsp4cohdec.m:198-210"""
    beta2l = -dispersion * lam ** 2 / 2 / math.pi / 299792458.0 * 1e-21
    sps = 1 + (0 if workatbaudrate else 1)
    s = np.asarray(signals, dtype=np.complex128)
    h = disp_comp_filter(beta2l, sps * baudrate, s.shape[0], ndispsym * sps)
    return np.fft.ifft(np.fft.fft(s, axis=0) * h[:, None], axis=0)


def adc_quantize(irx, bits):
    """dsp4cohdec.m:159-161 (round = half away from zero, as the interpreter's)"""
    irx = np.asarray(irx, dtype=np.float64)
    m = np.max(np.abs(irx))
    v = (irx + m) / 2 / m * 2 ** bits
    r = np.sign(v) * np.floor(np.abs(v) + 0.5)
    return r * 2 * m / 2 ** bits - m


def nl_rotation(s, alpha):
    """NLRotation, dsp4cohdec.m:308-315"""
    s = np.asarray(s, dtype=np.complex128)
    phases, amps = np.angle(s), np.abs(s)
    asq = np.sum(amps * amps, axis=1)
    dp = asq - np.mean(asq)
    phases = phases + (alpha * dp)[:, None] * np.ones((1, s.shape[1]))
    return amps * (np.cos(phases) + 1j * np.sin(phases))


def circ_moving_average(s, k):
    """The smoothing of vitvit (dsp4cohdec.m:329-340): ifft(fft(s).*fft(ones(N,1)/N, L)), N = 2k+1, i.e. the circular
    causal average out(n) = mean(s(n-N+1 .. n)).  Evaluated directly (the transform pair is an implementation detail of
    the interpreter code; the two agree to rounding)."""
    s = np.asarray(s, dtype=np.complex128)
    L, N = s.shape[0], 2 * k + 1
    if N >= L:
        reps = math.ceil(N / L)
        long = np.concatenate([s] * reps)
        return circ_moving_average_exact(long, N)[:L]
    return circ_moving_average_exact(s, N)


def circ_moving_average_exact(s, N):
    L = s.shape[0]
    out = np.zeros_like(s)
    for j in range(N):
        out += np.roll(s, j, axis=0)
    return out / N


def vitvit(s, P, M, k, applyunwrap):
    """dsp4cohdec.m:320-345"""
    s = np.asarray(s, dtype=np.complex128)
    if P == M:
        s = s ** P
    else:
        s = np.abs(s) ** P * np.exp(1j * np.angle(s ** M))
    if k > 0:
        s = circ_moving_average(s, k)
    if applyunwrap:
        return np.unwrap(np.angle(s), axis=0) / M
    return np.angle(s) / M


def carrier_recovery(signals, modorder=2, freqavg=500, phasavg=3, poworder=2):
    """dsp4cohdec.m:241-283 -> Phases = angle(Signals .* Carrier), [L, npol]"""
    s = np.asarray(signals, dtype=np.complex128)
    M = 2 ** modorder
    off = math.pi / 4 if modorder > 1 else 0.0
    if freqavg:
        om = np.cumsum(vitvit(s * np.conj(np.roll(s, 1, axis=0)), M, M, freqavg, False), axis=0)
        closest = om[0] + np.round((om[-1] - om[0]) / 2 / math.pi) * 2 * math.pi
        ratio = closest / om[-1]
        om = (om - om[0]) * ratio + om[0]
        demod = s * np.exp(-1j * om)
        theta = vitvit(demod, poworder, M, phasavg, True)
        carrier = np.exp(1j * (-om - theta + off))
    else:
        theta = vitvit(s, poworder, M, phasavg, True)
        carrier = np.exp(1j * (-theta + off))
    return np.angle(s * carrier)


def samp2pat_coherent(phase):
    """samp2pat.m:60-67 -> [first_bit second_bit] per polarization, [L, 2*npol] of 0/1"""
    phase = np.asarray(phase, dtype=np.float64)
    second = phase > 0
    first = np.abs(phase) <= math.pi / 2
    cols = []
    for p in range(phase.shape[1]):
        cols += [first[:, p], second[:, p]]
    return np.stack(cols, axis=1).astype(np.uint8)


def pat_decoder_dqpsk_binary(patmat):
    """pat_decoder(patmat,'dqpsk',struct('binary',true)) (pat_decoder.m:68-82) -> (pat [L], patmat [L,2])"""
    patmat = np.asarray(patmat).astype(np.int64)
    stars = np.zeros(patmat.shape[0], dtype=np.complex128)            # pat2stars.m:51-63
    a, b = patmat[:, 0], patmat[:, 1]
    stars[(a == 0) & (b == 0)] = 1
    stars[(a == 0) & (b == 1)] = 1j
    stars[(a == 1) & (b == 1)] = -1
    stars[(a == 1) & (b == 0)] = -1j
    r = np.conj(stars) * np.roll(stars, 1)                             # conj(stars_t).*fastshift(stars_t,1)
    pat = np.zeros(len(r), dtype=np.int64)                             # stars2pat.m:28-41
    pm = np.zeros((len(r), 2), dtype=np.int64)
    for val, p, bits in ((1, 0, (0, 0)), (1j, 1, (0, 1)), (-1, 3, (1, 1)), (-1j, 2, (1, 0))):
        m = r == val
        pat[m] = p
        pm[m, 0], pm[m, 1] = bits
    return 3 - pat, 1 - pm                                             # pat = 3-pat; patmat = ~patmat


def count_errors_dqpsk(phases_rx, tx_phase):
    """Decision + differential decoding + X/Y swap test + error count (ex20_coherent_polmux.m:160-178,
    ber_estimate.m:118).  phases_rx: [L,2] recovered phases; tx_phase: [L,2] phases of the transmitted symbols."""
    hat = samp2pat_coherent(phases_rx)
    ref = samp2pat_coherent(tx_phase)
    hx, hy = pat_decoder_dqpsk_binary(hat[:, 0:2])[1], pat_decoder_dqpsk_binary(hat[:, 2:4])[1]
    rx, ry = pat_decoder_dqpsk_binary(ref[:, 0:2])[1], pat_decoder_dqpsk_binary(ref[:, 2:4])[1]
    if np.sum(rx != hy) < np.sum(rx != hx):
        hx, hy = hy, hx
    return int(np.sum(rx != hx) + np.sum(ry != hy))
