"""create_field: host-side boundary of the path (create_field.m:79-202).

Stays host code as in the reference (O(N) once per run, not in the SSFM loop);
its job is to leave GSTATE.FIELDX / GSTATE.FIELDY in the layout fiber() reads:
[Nfft, nfc] complex columns, nfc = NCH ('sepfields') or 1 ('unique').
"""
from __future__ import annotations

import numpy as np

from .gstate import CONSTANTS, GSTATE

# True: create_field('unique', ...) of a two-polarization field multiplexes the channels on the device
# (pmx_field_mux: one pointwise modulation pass instead of the reference's fft / fastshift / ifft) and leaves
# GSTATE.FIELDX/FIELDY resident in HBM for the first fiber().  False (default): the reference's host arithmetic.
DEVICE_MUX = False


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
    ft = ftype.lower()
    on_device = DEVICE_MUX and ft == 'unique' and isy
    scale = np.ones(G.NCH)
    if str(options.get('power', '')).lower() == 'average':                   # :113-124
        avge = np.mean(np.abs(sigx) ** 2 + (np.abs(sigy) ** 2 if isy else 0.0), axis=0)
        scale = np.sqrt(power / avge)
        if not on_device:
            sigx = sigx * scale[None, :]
            if isy:
                sigy = sigy * scale[None, :]
        G.POWER = power * power / avge
    tau = np.zeros((npol, G.NCH))
    if 'delay' in options:                                                   # :127-146
        if isinstance(options['delay'], str) and options['delay'] == 'rand':
            from .gstate import rng
            tau = np.round(rng().random((npol, G.NCH)) * G.NT)
        else:
            d = np.asarray(options['delay'], dtype=np.float64)
            if d.shape != (npol, G.NCH):
                raise ValueError('the delay must be of size [number of polarizations,number of channels]')
            tau = np.round(d * G.NT)
        if not on_device:
            for kch in range(G.NCH):
                sigx[:, kch] = _shift(sigx[:, kch], tau[0, kch])
                if isy:
                    sigy[:, kch] = _shift(sigy[:, kch], tau[1, kch])
        G.DELAY = tau
    else:
        G.DELAY = np.zeros((npol, G.NCH))
    G.DISP = np.zeros((npol, G.NCH))                                         # :151
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
        if on_device:
            # fastshift(fft(sig), -ndfn) is the modulation exp(-2*pi*i*ndfn*n/Nfft) in time: one pointwise pass
            # over the channels on the device, no transform; the field stays in HBM for the first fiber()
            from . import _lib
            ctx = _lib.default_context()
            fld = _lib.DeviceField(ctx, nfft, 1, 1)
            try:
                _lib.field_mux(ctx, fld, sigx, sigy, ndfn, scale, tau[0], tau[1])
                ox, oy = fld.download()
            except Exception:
                fld.close()
                raise
            G.FIELDX_TX = np.ascontiguousarray(ox[0].T)
            G.FIELDY_TX = np.ascontiguousarray(oy[0].T)
            G.put_device(fld, None, None)
            return
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
