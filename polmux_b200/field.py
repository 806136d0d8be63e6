"""create_field: host-side boundary of the path (create_field.m:79-202).

Stays host code as in the reference (O(N) once per run, not in the SSFM loop);
its job is to leave GSTATE.FIELDX / GSTATE.FIELDY in the layout fiber() reads:
[Nfft, nfc] complex columns, nfc = NCH ('sepfields') or 1 ('unique').
"""
from __future__ import annotations

import numpy as np

from .gstate import CONSTANTS, GSTATE


def _shift(v, n):  # fastshift.m:45-58 (n > 0 delays)
    return np.roll(v, int(n), axis=0)


def create_field(ftype, sigx, sigy=None, options=None):
    """create_field(ftype,sigx,sigy,options) -- create_field.m:1.

    options: dict with optional 'power': 'average' (:113-124) and 'delay'
    ('rand' or an [npol, NCH] array, :127-149)."""
    G = GSTATE
    nfft = G.NSYMB * G.NT
    options = dict(options or {})
    for k in options:
        if k not in ('delay', 'power'):
            raise ValueError("unknown option '%s'" % k)                      # checkfields.m:32-40
    sigx = np.asarray(sigx)
    if sigx.size == 0:
        raise ValueError('empty x component')
    sigx = np.array(sigx, dtype=np.complex128).reshape(nfft, -1)
    isy = sigy is not None and np.size(sigy) != 0
    npol = 2 if isy else 1
    if isy:
        sigy = np.array(sigy, dtype=np.complex128).reshape(nfft, -1)
        if sigy.shape != sigx.shape:
            raise ValueError('sigx and sigy must have the same size')
    if sigx.shape[1] != G.NCH:
        raise ValueError('the number of columns of sigx,sigy must be equal to the number of channels')
    power = np.asarray(G.POWER, dtype=np.float64).reshape(-1)
    if str(options.get('power', '')).lower() == 'average':                   # :113-124
        avge = np.mean(np.abs(sigx) ** 2 + (np.abs(sigy) ** 2 if isy else 0.0), axis=0)
        s = np.sqrt(power / avge)
        sigx = sigx * s[None, :]
        if isy:
            sigy = sigy * s[None, :]
        G.POWER = power * power / avge
    if 'delay' in options:                                                   # :127-146
        if isinstance(options['delay'], str) and options['delay'] == 'rand':
            from .gstate import rng
            tau = np.round(rng().random((npol, G.NCH)) * G.NT)
        else:
            d = np.asarray(options['delay'], dtype=np.float64)
            if d.shape != (npol, G.NCH):
                raise ValueError('the delay must be of size [number of polarizations,number of channels]')
            tau = np.round(d * G.NT)
        for kch in range(G.NCH):
            sigx[:, kch] = _shift(sigx[:, kch], tau[0, kch])
            if isy:
                sigy[:, kch] = _shift(sigy[:, kch], tau[1, kch])
        G.DELAY = tau
    else:
        G.DELAY = np.zeros((npol, G.NCH))
    G.DISP = np.zeros((npol, G.NCH))                                         # :151
    ft = ftype.lower()
    if ft == 'sepfields':                                                    # :156-162
        G.FIELDX_TX, G.FIELDX = sigx.copy(), sigx
        if isy:
            G.FIELDY_TX, G.FIELDY = sigy.copy(), sigy
        else:
            G.FIELDY = None
    elif ft == 'unique':                                                     # :164-199
        lamt = np.asarray(G.LAMBDA, dtype=np.float64).reshape(-1)
        maxl, minl = lamt.max(), lamt.min()
        fnyqmin = (CONSTANTS.CLIGHT / minl - CONSTANTS.CLIGHT / maxl) / G.SYMBOLRATE
        if G.NT < fnyqmin and fnyqmin != 0:
            # the reference asks interactively (:171); a library cannot
            raise ValueError('number of samples per symbol is too small')
        lamc = 2 * maxl * minl / (maxl + minl)
        deltafn = CONSTANTS.CLIGHT * (1 / lamc - 1.0 / lamt)
        minfreq = G.FN[1] - G.FN[0]
        ndfn = np.round(deltafn / G.SYMBOLRATE / minfreq).astype(np.int64)
        zx = np.fft.fft(sigx, axis=0)
        fx = np.zeros(nfft, dtype=np.complex128)
        fy = np.zeros(nfft, dtype=np.complex128) if isy else None
        for kch in range(G.NCH):
            fx = fx + _shift(zx[:, kch], -ndfn[kch])
            if isy:
                fy = fy + _shift(np.fft.fft(sigy[:, kch]), -ndfn[kch])
        G.FIELDX = np.fft.ifft(fx)[:, None]
        G.FIELDX_TX = G.FIELDX.copy()
        if isy:
            G.FIELDY = np.fft.ifft(fy)[:, None]
            G.FIELDY_TX = G.FIELDY.copy()
        else:
            G.FIELDY = None
    else:
        raise ValueError("ftype must be 'sepfields' or 'unique'")
