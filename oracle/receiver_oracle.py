"""NumPy restatement of the reference's coherent receiver front-end (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product never does.

Reference code followed (all under /root/reference):
  myfilter.m:45-152          the filter responses over a frequency vector
  evaldelay.m:28-80          their group delay at f = 0
  receiver_cohmix.m:63-307   channel selection (nind), post-compensating fiber, optical filter, local oscillator,
                             the four mixer outputs per polarization, photodiodes (balanced / single), low-pass filter

PARITY PINNING: tests/test_receiver_oracle.py compares this file with the reference's own myfilter.m, evaldelay.m and
receiver_cohmix.m executed by the mini interpreter (oracle/mini_m) -- goldens under tests/golden/rx/ written by
oracle/make_golden.py rx.
"""
from __future__ import annotations

import math

import numpy as np

CLIGHT = 299792458.0


def myfilter(ftype, f, bw, order=0):
    """myfilter.m:73-152 -> Hf over f(:)"""
    r4p2r2 = 2.61312592975275
    b1, b2, b3 = 3.86370330515627315, 7.4641016151377546, 9.1416201726856413
    b4, b5 = b2, b1
    Bb = 0.3863
    d0, d1, d2, d3, d4 = 945, 945, 420, 105, 15
    x = np.asarray(f, dtype=np.float64).reshape(-1) / bw
    ftype = ftype.lower()
    if ftype == 'movavg':
        return np.sinc(x) + 0j
    if ftype == 'gauss':
        return np.exp(-0.5 * math.log(2) * x * x) + 0j
    if ftype == 'gauss_off':
        return np.exp(-0.5 * math.log(2) * (x - order / bw) * (x - order / bw)) + 0j
    if ftype == 'butt2':
        return 1 / (1 - x * x + 1j * math.sqrt(2) * x)
    if ftype == 'butt4':
        x2 = x * x
        umx2 = 1 - x2
        return 1 / (umx2 * umx2 - math.sqrt(2) * x2 + 1j * r4p2r2 * x * umx2)
    if ftype == 'butt6':
        x2 = x * x
        x3 = x2 * x
        x4 = x3 * x
        x5 = x4 * x
        x6 = x5 * x
        return 1 / (1. - b2 * x2 + b4 * x4 - x6 + 1j * (b1 * x - b3 * x3 + b5 * x5))
    if ftype == 'ideal':
        return (np.abs(x) <= 1) + 0j
    if ftype == 'bessel5':
        om = 2 * math.pi * x * Bb
        om2 = om * om
        om3 = om2 * om
        om4 = om3 * om
        om5 = om4 * om
        pre = d0 - d2 * om2 + d4 * om4
        pim = d1 * om - d3 * om3 + om5
        return d0 / (pre + 1j * pim)
    if ftype == 'rc1':
        return 1 / (1 + 1j * x)
    if ftype == 'rc2':
        return 1 / (1 + 1j * math.sqrt(math.sqrt(2) - 1) * x) ** 2
    if ftype == 'supergauss':
        return np.exp(-0.5 * math.log(2) * x ** (2 * order)) + 0j
    raise ValueError('the filter ftype does not exist.')


def evaldelay(ftype, bw):
    """evaldelay.m:28-80"""
    r4p2r2, b1, Bb = 2.61312592975275, 3.86370330515627315, 0.3863
    table = {'movavg': 0, 'gauss': 0, 'gauss_off': 0, 'ideal': 0, 'supergauss': 0,
             'butt2': 1.11 * math.sqrt(2) / (2 * math.pi * bw), 'butt4': 1.1 * r4p2r2 / (2 * math.pi * bw),
             'butt6': 1.1 * b1 / (2 * math.pi * bw), 'bessel5': Bb / bw, 'rc1': 1 / (2 * math.pi * bw),
             'rc2': (math.sqrt(2) - 1) / (math.pi * bw)}
    if ftype.lower() not in table:
        raise ValueError('the filter ftype does not exist.')
    return float(table[ftype.lower()])


def fastexp(x):
    return np.cos(x) + 1j * np.sin(x)


def receiver_cohmix(gs, ich, x):
    """receiver_cohmix.m:63-307.  gs: an oracle GState (FN, LAMBDA, SYMBOLRATE, NSYMB, NT, NCH, POWER, FIELDX/FIELDY and,
    for b2b, FIELDX_TX/FIELDY_TX); x: dict.  -> (Iric [Nfft, 2 or 4], x updated)"""
    x = dict(x)
    x.setdefault('oord', 0)
    x.setdefault('eord', 0)
    fn = np.asarray(gs.FN, dtype=np.float64).reshape(-1)
    nfft = fn.size
    nfr, nfc = gs.FIELDX.shape
    npoints = np.arange(1, nfft + 1)
    lam = np.asarray(gs.LAMBDA, dtype=np.float64).reshape(-1)
    maxl, minl = lam.max(), lam.min()
    lamc = 2 * maxl * minl / (maxl + minl)
    if nfc != gs.NCH:
        minfreq = fn[1] - fn[0]
        deltafn = CLIGHT * (1 / lamc - 1 / lam[ich - 1])
        ndfn = int(np.round(deltafn / gs.SYMBOLRATE / minfreq))
        nind = np.mod(npoints - ndfn - 1, nfft)          # nmod(npoints-ndfn, Nfft), zero-based
        nch = 1
        if ich == 1:
            ndfnl = nfft // 2
        else:
            dl = CLIGHT * (1 / lamc - 1 / lam[ich - 2])
            ndfnl = int(np.round(dl / gs.SYMBOLRATE / minfreq))
            ndfnl = int(np.round((ndfn - ndfnl) * 0.5))
        if ich == gs.NCH:
            ndfnr = nfft // 2
        else:
            dr = CLIGHT * (1 / lamc - 1 / lam[ich])
            ndfnr = int(np.round(dr / gs.SYMBOLRATE / minfreq))
            ndfnr = int(np.round((ndfnr - ndfn) * 0.5))
    else:
        nind = npoints - 1
        nch, ndfnl, ndfnr = ich, nfft // 2, nfft // 2
    b2b = False
    if 'b2b' in x:
        if x['b2b'] != 'b2b':
            raise ValueError("the b2b field must be 'b2b'")
        b2b = True
        x.pop('dpost', None)
    sigx = (gs.FIELDX_TX if b2b else gs.FIELDX)[:, nch - 1]
    if 'dpost' in x:
        lm = x['lambda']
        b20z = -lm ** 2 / 2 / math.pi / CLIGHT * x['dpost'] * 1e-3
        b30z = (lm / 2 / math.pi / CLIGHT) ** 2 * (2 * lm * x['dpost'] + lm ** 2 * x['slopez']) * 1e-3
        domega_i0 = 2 * math.pi * CLIGHT * (1. / lam[ich - 1] - 1 / lm)
        domega_ic = 2 * math.pi * CLIGHT * (1. / lam[ich - 1] - 1 / lamc)
        domega_c0 = 2 * math.pi * CLIGHT * (1. / lamc - 1 / lm)
        beta1z = b20z * domega_ic + 0.5 * b30z * (domega_i0 ** 2 - domega_c0 ** 2)
        beta2z = b20z + b30z * domega_i0
        omega = 2 * math.pi * gs.SYMBOLRATE * fn
        betat = omega * beta1z + 0.5 * omega ** 2 * beta2z + omega ** 3 * b30z / 6
        x['post_delay'] = gs.SYMBOLRATE * beta1z
        hf = fastexp(-betat)
    else:
        hf = np.ones(nfft)
        x['post_delay'] = 0
    hf = hf * myfilter(x['oftype'], fn, 0.5 * x['obw'], x['oord'])
    pch = float(np.asarray(gs.POWER, dtype=np.float64).reshape(-1)[ich - 1])

    def band_energy(s):
        return (np.sum(np.abs(s[:ndfnl]) ** 2) + np.sum(np.abs(s[nfft - ndfnr:]) ** 2)) / pch / nfft ** 2

    sx = np.fft.fft(sigx)[nind]
    x['avgebx'] = band_energy(sx)
    sx = sx * hf
    # local oscillator
    if x.get('lodetuning'):
        minfreq = gs.SYMBOLRATE * 1e9 / gs.NSYMB
        kdet = math.floor(x['lodetuning'] / minfreq)
        lo_detuning = 2 * math.pi * kdet / nfft * np.arange(1, nfft + 1)
    else:
        lo_detuning = 0
    if 'lophasenoise' in x:
        if len(x['lophasenoise']) != nfft:
            raise ValueError('Incompatible vector.')
        lo_pn = np.asarray(x['lophasenoise'], dtype=np.float64).reshape(-1)
    elif 'lolinewidth' in x:
        raise NotImplementedError('the oracle takes the phase noise from the caller (x.lophasenoise)')
    else:
        lo_pn = np.zeros(nfft)
    ecw = 10 ** (x['lopower'] / 20) if 'lopower' in x else 1
    elo = ecw * fastexp(lo_detuning + lo_pn)
    isy = gs.FIELDY is not None
    if isy:
        if b2b:
            ty = getattr(gs, 'FIELDY_TX', None)
            sy = np.zeros(nfr, dtype=np.complex128) if ty is None else np.fft.fft(ty[:, nch - 1])
        else:
            sy = np.fft.fft(gs.FIELDY[:, nch - 1])
        sy = sy[nind]
        x['avgeby'] = band_energy(sy)
        sy = sy * hf
        sy = np.fft.ifft(sy)
    sx = np.fft.ifft(sx)
    balanced = not (x.get('pdtype') == 'normal')

    def detect(s):
        emix = np.stack([s * 1j + elo * 1j, s - elo, s * 1j - elo, -s + elo * 1j], axis=1)
        iric = np.real(emix * np.conj(emix))
        if balanced:
            return np.stack([iric[:, 0] - iric[:, 1], iric[:, 2] - iric[:, 3]], axis=1)
        return np.stack([iric[:, 0], iric[:, 2]], axis=1)

    he = myfilter(x['eftype'], fn, x['ebw'], x['eord'])[:, None]
    out = np.real(np.fft.ifft(np.fft.fft(detect(sx), axis=0) * he, axis=0))
    if isy:
        outy = np.real(np.fft.ifft(np.fft.fft(detect(sy), axis=0) * he, axis=0))
        out = np.concatenate([out, outy], axis=1)
    return out, x
