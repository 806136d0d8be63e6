"""ampliflat(x, atype, options): flat-gain amplifier with ASE (ampliflat.m:1,61-148)."""
from __future__ import annotations

import math

import numpy as np

from . import _lib
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


def ampliflat(x, atype='gain', options=None, ctx=None, seed=0):
    """Host-buffer form, like the reference: GSTATE.FIELDX/FIELDY in place.

    options: {'f': noise figure [dB], 'noise': [Nfft, 2*nfc] complex standard normals}.
    Without options.noise the ASE comes from the device's counter-based generator (seed)."""
    G = GSTATE
    if atype.lower() != 'gain':
        raise NotImplementedError("ampliflat: only atype 'gain' is built (ampliflat.m:61-63)")
    options = dict(options or {})
    if 'onepol' in options:
        raise NotImplementedError('ampliflat: options.onepol is not built')
    nfr, nfc = G.FIELDX.shape
    gain = 10 ** (x * 0.1)
    sigma = ase_sigma(gain, options.get('f'), nfc) if options else np.zeros(nfc)
    ctx = ctx or _lib.default_context()
    fld = _lib.DeviceField(ctx, nfr, nfc, 1)
    fy = G.FIELDY if G.FIELDY is not None else np.zeros_like(G.FIELDX)
    fld.upload(G.FIELDX, fy)
    noise = None
    if np.any(sigma) and 'noise' in options:
        nz = np.asarray(options['noise'], dtype=np.complex128)
        noise = np.ascontiguousarray(nz.T)[None]                 # [1][2*nfc][nfft]
    _lib.ampliflat_exec(ctx, fld, gain, sigma, noise, seed)
    inplace = (nfc == 1 and G.FIELDY is not None and all(
        a.dtype == np.complex128 and a.flags['C_CONTIGUOUS'] and a.flags['WRITEABLE'] for a in (G.FIELDX, G.FIELDY)))
    if inplace:          # single column: [N,1] is also [1][1][N]; results land in the caller's buffers
        fld.download_into(G.FIELDX, G.FIELDY)
        fld.close()
        return
    ox, oy = fld.download()
    fld.close()
    G.FIELDX = np.ascontiguousarray(ox[0].T)
    if G.FIELDY is not None or np.any(sigma):
        if G.FIELDY is None:
            G.DELAY = np.vstack([G.DELAY[:1], np.zeros((1, G.NCH))])          # ampliflat.m:144
        G.FIELDY = np.ascontiguousarray(oy[0].T)
