"""link(x, flag, nspan, gain_db, options): the span loop of the reference's scripts as one library call.

    for k = 1:Nspan                         ex06_ber.m:110-115, ex20_coherent_polmux.m
        fiber(x, flag);
        ampliflat(Gerbio, 'gain', opt);
    end

Same result, bit for bit, as calling polmux_b200.fiber() and polmux_b200.ampliflat() nspan times (waveplates are
drawn per span from the same stream in the same order, fiber.m:274-276), but the loop runs inside the C ABI
(pmx_link_run): the field goes to the device once and comes back once.  This is the call a MEX gateway without
persistent state binds for multi-span scripts."""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _lib
from .ampliflat import ase_sigma
from .fiber import LAST, apply_side_effects, fiber_setup, setup_to_desc
from . import gstate
from .gstate import GSTATE


def link(x, flag: str, nspan: int, gain_db: Optional[float] = None, options=None, rng=None, ctx=None, seed: Optional[int] = None,
         disp_mode: Optional[str] = None, precision: Optional[str] = None):
    """-> list of the nspan brf structs fiber() would have returned.
    options: ampliflat's ({'f': noise figure [dB], 'noise': list of nspan arrays [Nfft, 2*nfc]}); the ASE of span k
    comes from options['noise'][k] or from the device generator with seed `seed + k` (seed=None: one fresh seed per
    span from the global stream, as ampliflat() draws them)."""
    G = GSTATE
    nspan = int(nspan)
    if nspan < 1:
        raise ValueError('nspan must be at least 1')
    options = dict(options or {})
    asepol = 3
    if 'onepol' in options:                                                              # ampliflat.m:107-118
        pol = str(options['onepol']).lower()
        if pol not in ('asex', 'asey'):
            raise ValueError("ONEPOL, if exists, must be 'asex' or 'asey'")
        asepol = 1 if pol == 'asex' else 2
    setups = [fiber_setup(x, flag, rng) for _ in range(nspan)]          # one plate draw per span, in call order
    s = setups[0]
    if not (s.isv and G.has_y()):
        raise NotImplementedError('link: two-polarization fields only (call fiber/ampliflat for the scalar path)')
    if s.fls[3]:
        raise NotImplementedError('The CNLSE with separate fields is not yet implemented')    # fiber.m:854
    if s.tolflag == 2:
        raise ValueError('adaptive step available in absence of polarization effects')       # fiber.m:372-374
    ctx = ctx or _lib.default_context()
    desc, keep = setup_to_desc(s, disp_mode=disp_mode, precision=precision)
    plates = [np.stack([st.brf[k] for st in setups])[:, None, :] for k in ('db0', 'theta', 'epsilon')]
    gain, sigma, noise = 0.0, None, None
    if gain_db is not None:
        gain = 10 ** (gain_db * 0.1)
        sigma = ase_sigma(gain, options.get('f'), s.nfc) if options else np.zeros(s.nfc)
        if np.any(sigma) and 'noise' in options:
            nz = options['noise']
            if len(nz) != nspan:
                raise ValueError('options.noise: one [Nfft, 2*nfc] array per span')
            noise = np.stack([np.ascontiguousarray(np.asarray(a, dtype=np.complex128).T)[None] for a in nz])
    draws = gain_db is not None and sigma is not None and np.any(sigma) and noise is None
    if seed is None:
        seeds = [gstate.next_ase_seed() if draws else 0 for _ in range(nspan)]
    else:
        seeds = [int(seed) + k for k in range(nspan)]
    ldesc, lkeep = _lib.make_link(nspan, gain, sigma, plates=plates, plate_sets=1, noise=noise,
                                  seeds=seeds, asepol=asepol)
    fx = np.ascontiguousarray(np.asarray(G.FIELDX, dtype=np.complex128).T)[None]             # [1][nfc][nfft]
    fy = np.ascontiguousarray(np.asarray(G.FIELDY, dtype=np.complex128).T)[None]
    io = _lib.complex_field(fx, fy)
    res = _lib.Result(nspan)
    ctx.check(ctx.lib.pmx_link_run(ctx.h, ctypes.byref(desc), ctypes.byref(ldesc), ctypes.byref(io), ctypes.byref(res.c)))
    G.FIELDX = np.ascontiguousarray(fx[0].T)
    G.FIELDY = np.ascontiguousarray(fy[0].T)
    brfs = []
    for st in setups:
        apply_side_effects(st)                                                               # fiber.m:367-369
        brf = dict(st.brf)
        brf['lcorr'] = st.length / st.nplates
        brf['betat'] = st.betat
        brf['db1'] = st.db1
        brfs.append(brf)
    LAST.clear()
    LAST.update(firstdz=float(res.firstdz[-1]), ncycle=int(res.ncycle[-1]), ntot=int(res.ntot[-1]),
                ncycle_per_span=[int(v) for v in res.ncycle])
    return brfs
