"""ampliflat(x, atype, options): flat-gain amplifier with ASE (ampliflat.m:1,61-148)."""
from __future__ import annotations

import math

import numpy as np

from . import _lib
from . import gstate
from .gstate import CONSTANTS, GSTATE


def ase_sigma(gain: float, f_db, nfc: int) -> np.ndarray:
    """ampliflat.m:91-106: sigma [sqrt(mW)] per column (zeros when options.f is absent or Inf)."""
    if f_db is None or math.isinf(f_db):
        return np.zeros(nfc)
    G = GSTATE
    flin = 10 ** (f_db * 0.1)
    lams = np.asarray(G.LAMBDA, dtype=np.float64).reshape(-1)
    if nfc == 1:
        maxl, minl = lams.max(), lams.min()
        lam = np.array([2 * maxl * minl / (maxl + minl)])
    else:
        lam = lams
    return np.sqrt(flin / 4 * CONSTANTS.HPLANCK * CONSTANTS.CLIGHT / lam * (gain - 1) * G.NT * G.SYMBOLRATE * 1e21)


def _avg_power_abs(G, ctx, ch):
    """avg_power(ch,'abs') for separate channels (avg_power.m:63-76: no filter, every bin counted): the mean over the
    samples of |FIELDX(:,ch)|^2 + |FIELDY(:,ch)|^2, reduced on the device where the field lives."""
    if G.has_y():
        fld, hx, hy = G.take_device(ctx)
        try:
            return float(_lib.field_mean_power(ctx, fld)[0, ch - 1])
        finally:
            G.restore_host(fld, hx, hy)          # (untouched: back as it was, resident or host)
    nfr, nfc = G.field_shape()
    fld = _lib.DeviceField(ctx, nfr, nfc, 1)
    try:
        fld.upload(G.FIELDX, np.zeros_like(G.FIELDX))
        return float(_lib.field_mean_power(ctx, fld)[0, ch - 1])
    finally:
        fld.close()


def ampliflat(x, atype='gain', options=None, ctx=None, seed=None):
    """ampliflat(x,'gain',options) on GSTATE.FIELDX/FIELDY, like the reference.

    options: {'f': noise figure [dB], 'noise': [Nfft, 2*nfc] complex standard normals, 'onepol': 'asex' | 'asey'}.
    Without options.noise the ASE comes from the device's counter-based generator, keyed by `seed`; seed=None (the
    default) takes the next value of the global stream (gstate.seed(k) = randn('state',k)), so successive calls add
    independent noise as the reference's randn does (ampliflat.m:132-135)."""
    G = GSTATE
    atype = atype.lower()
    if atype not in ('gain', 'fixpower'):
        raise ValueError('wrong string atype')                                   # ampliflat.m:74-75
    options = dict(options or {})
    asepol = 3
    if 'onepol' in options:                                                  # ampliflat.m:107-118
        pol = str(options['onepol']).lower()
        if pol not in ('asex', 'asey'):
            raise ValueError("ONEPOL, if exists, must be 'asex' or 'asey'")
        asepol = 1 if pol == 'asex' else 2
    nfr, nfc = G.field_shape()
    if atype == 'fixpower' and nfc != G.NCH:
        raise ValueError("'fixpower' works only for channels separated")         # ampliflat.m:66,70-71
    ctx = ctx or _lib.default_context()
    if atype == 'gain':
        gain = 10 ** (x * 0.1)                                                   # ampliflat.m:62-63
    else:                                                                        # ampliflat.m:65-69
        gain = x / _avg_power_abs(G, ctx, math.ceil(nfc / 2))                    # gain = x/avg_power(midch,'abs')
    sigma = ase_sigma(gain, options.get('f'), nfc) if options else np.zeros(nfc)
    noise = None
    if seed is None:
        seed = gstate.next_ase_seed() if np.any(sigma) and 'noise' not in options else 0
    if np.any(sigma) and 'noise' in options:
        nz = np.asarray(options['noise'], dtype=np.complex128)
        noise = np.ascontiguousarray(nz.T)[None]                 # [1][2*nfc][nfft]
    if G.has_y():
        # two polarizations: works on the field fiber() left in HBM (or uploads it) and leaves it there
        fld, hx, hy = G.take_device(ctx)
        try:
            _lib.ampliflat_exec(ctx, fld, gain, sigma, noise, seed, asepol)
        except Exception:
            G.restore_host(fld, hx, hy)
            raise
        G.put_device(fld, hx, hy)
        return
    # single polarization: host buffers in, host buffers out (ASE creates FIELDY, ampliflat.m:132-146)
    fld = _lib.DeviceField(ctx, nfr, nfc, 1)
    fld.upload(G.FIELDX, np.zeros_like(G.FIELDX))
    _lib.ampliflat_exec(ctx, fld, gain, sigma, noise, seed, asepol)
    ox, oy = fld.download()
    fld.close()
    G.FIELDX = np.ascontiguousarray(ox[0].T)
    if np.any(sigma) and (asepol & 2):
        G.DELAY = np.vstack([G.DELAY[:1], np.zeros((1, G.NCH))])              # ampliflat.m:144
        G.FIELDY = np.ascontiguousarray(oy[0].T)
