"""fiber(x, flag): the reference's optical-fiber operator, SSFM loop on a B200.

Mirror of fiber.m: same struct fields (fiber.m:8-52), same 4-character flag
(:54-71, :157-251), same side effects on GSTATE.FIELDX / FIELDY / DELAY / DISP
(:286, :367-369, :376-388) and the same ``brf`` return (:384).  Everything up
to the dispatch at fiber.m:372 is host scalar set-up and is done here in IEEE
double; the propagation loop itself (matrix_ssfm, :459-555) runs inside the
C-ABI CUDA library through ``pmx_fiber_run`` -- one call per fiber(), host
buffers in, host buffers out.  There is no CPU fallback.

The scalar single-polarization path (scalar_ssfm, nl_step incl. cross-column XPM,
:557-636, :786-803), the local-error adaptive step (x.ltol: scalar_a_ssfm / adaptssfm,
:639-679, :938-1010) and x.dphiadapt (:588-611) run on the device as well; what raises
is what the reference raises for (two polarizations + 'x', :853-854; ltol with
polarization effects, :372-374).
"""
from __future__ import annotations

import dataclasses
import math
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib, simul_out
from .gstate import CONSTANTS, GSTATE, rng as _global_rng

DEF_PLATES = 100  # fiber.m:131
SAFETYFCT = 0.9   # fiber.m:130: safety step-reduction factor of the adaptive step

# flag -> (g, p, s, x) and whether the run is forced to a single linear step
_FLAG_TABLE = {
    '----': (0, 0, 0, 0), 'g---': (1, 0, 0, 0), '-p--': (0, 1, 0, 0), '--s-': (0, 0, 1, 0),
    '---x': (0, 0, 0, 1), 'gp--': (1, 1, 0, 0), 'g-s-': (1, 0, 1, 0), 'g--x': (1, 0, 0, 1),
    '-ps-': (0, 1, 1, 0), '-p-x': (0, 1, 0, 1), '--sx': (0, 0, 1, 1), 'g-sx': (1, 0, 1, 1),
    '-psx': (0, 1, 1, 1), 'gps-': (1, 1, 1, 0), 'gp-x': (1, 1, 0, 1), 'gpsx': (1, 1, 1, 1),
}
_LINEAR = ('----', 'g---', '-p--', 'gp--')
_X_ONLY_MULTI = ('---x', 'g--x', '-p-x', 'gp-x')      # error when nfc == 1
_X_OPTIONAL = ('--sx', 'g-sx', '-psx', 'gpsx')         # x silently dropped when nfc == 1


def _get(x, name, default=None):
    if isinstance(x, dict):
        return x.get(name, default)
    return getattr(x, name, default)


def _has(x, name):
    return (name in x) if isinstance(x, dict) else hasattr(x, name)


@dataclass
class FiberSetup:
    """Everything fiber.m:126-369 computes before the SSFM dispatch."""
    nfft: int
    nfc: int
    fls: tuple
    dphimaxt: float
    dzmaxt: float
    length: float
    alphalin: float
    gam: np.ndarray          # [nfc], before the Manakov 8/9
    betat: np.ndarray        # [nfft, nfc]
    db1: np.ndarray          # [nfft, nfc]
    manakov: bool
    nplates: int
    brf: dict
    isv: bool
    isy: bool
    b1: np.ndarray
    dch: np.ndarray
    scalars: dict            # beta1, beta2, b30, dgdrms, ... (scalar dispersion mode of the C ABI)
    tolflag: int = 0         # x.ltol: 2 = local-error adaptive step throughout, 1 = for the first step only
                             # (x.dphiadapt: it calibrates dphimax), fiber.m:143-155
    trg: Optional[dict] = None


def flag_to_fls(flag: str, nfc: int, x):
    """fiber.m:157-251 -> (fls, dphimaxt, dzmaxt)."""
    f = flag.lower()
    if f not in _FLAG_TABLE:
        raise ValueError("wrong flag. E.g. flag can be 'g---','gp--','-s--', etc")
    if f in _X_ONLY_MULTI and nfc == 1:
        raise ValueError("flag '%s' available only for channels separated" % f)
    g, p, s, xx = _FLAG_TABLE[f]
    if f in _X_OPTIONAL and nfc == 1:
        xx = 0
    single_exact = f in ('--s-', '--sx') and nfc == 1      # exact SPM solution, :172-174, :218-220
    if f in _LINEAR or single_exact:
        return (g, p, s, xx), math.inf, float(_get(x, 'length'))
    return (g, p, s, xx), float(_get(x, 'dphimax')), float(_get(x, 'dzmax'))


_DISP_CACHE = {}


def _dispersion_vectors(G, nfft, nfc, beta1, beta2, b30, dgdrms, pflag):
    """betat(omega), db1(omega) of fiber.m:350-362.  The vectors depend only on a handful of
    scalars and the FN grid, so successive spans with the same fiber reuse them."""
    key = (nfft, G.NSYMB, G.NT, float(G.SYMBOLRATE), tuple(np.asarray(beta1).tolist()),
           tuple(np.asarray(beta2).tolist()), float(b30), float(dgdrms), int(pflag))
    hit = _DISP_CACHE.get(key)
    if hit is not None:
        return hit
    omega = 2 * math.pi * G.SYMBOLRATE * np.asarray(G.FN, dtype=np.float64)  # :352
    betat = np.zeros((nfft, nfc))
    db1 = np.zeros((nfft, nfc))
    for k in range(nfc):                                                    # :354-362
        betat[:, k] = omega * beta1[k] + 0.5 * omega ** 2 * beta2[k] + omega ** 3 * b30 / 6
        if pflag:
            db1[:, k] = dgdrms * omega
    betat.setflags(write=False)
    db1.setflags(write=False)
    if len(_DISP_CACHE) > 8:
        _DISP_CACHE.clear()
    _DISP_CACHE[key] = (betat, db1)
    return betat, db1


def fiber_setup(x, flag: str, rng: Optional[np.random.Generator] = None) -> FiberSetup:
    """Host part of fiber(): parameter checks, PMD plates, unit conversions."""
    G = GSTATE
    if flag is None:
        raise ValueError('Missing propagation type')                        # fiber.m:137
    nfr, nfc = G.field_shape()
    nfft = G.NSYMB * G.NT
    length = float(_get(x, 'length'))
    xd = {k: _get(x, k) for k in ('dzmax', 'dphimax', 'length')}
    if not _has(x, 'dzmax') or xd['dzmax'] > length:                        # :139-141
        xd['dzmax'] = length
    tolflag, trg = 0, None
    if _has(x, 'ltol'):                                                     # :143-155
        if not _has(x, 'dphimax'):
            xd['dphimax'] = math.inf
        adapt_first = _has(x, 'dphiadapt') and bool(_get(x, 'dphiadapt'))           # :147-151
        tolflag, trg = (1 if adapt_first else 2), {'err': float(_get(x, 'ltol')), 'safety': SAFETYFCT}
    fls, dphimaxt, dzmaxt = flag_to_fls(flag, nfc, xd)

    isy = G.has_y()                                                         # :253
    isv = bool(fls[1]) or isy
    brf = {}
    if fls[1]:                                                              # :255-289
        manakov = str(_get(x, 'manakov', 'no')) == 'yes'
        if not _has(x, 'dgd'):
            raise ValueError('Missing DGD in fiber')
        dgd = float(_get(x, 'dgd'))
        given = [_has(x, k) for k in ('db0', 'theta', 'epsilon')]
        if all(given):                                                      # PMF, :264-269
            theta = np.atleast_1d(np.asarray(_get(x, 'theta'), dtype=np.float64)).ravel()
            nplates = theta.size
            brf['db0'] = np.atleast_1d(np.asarray(_get(x, 'db0'), dtype=np.float64)).ravel()
            brf['theta'] = theta
            brf['epsilon'] = np.atleast_1d(np.asarray(_get(x, 'epsilon'), dtype=np.float64)).ravel()
            if brf['db0'].size != nplates or brf['epsilon'].size != nplates:
                raise ValueError('db0, theta and epsilon must have the same length')
            dgdrms = dgd / nplates
        elif not any(given):                                                # random plates, :270-279
            nplates = int(_get(x, 'nplates', DEF_PLATES))
            r = rng if rng is not None else _global_rng()
            brf['db0'] = r.random(nplates) * 2 * np.pi - np.pi
            brf['theta'] = r.random(nplates) * np.pi - 0.5 * np.pi
            brf['epsilon'] = 0.5 * np.arcsin(r.random(nplates) * 2 - 1)
            dgdrms = math.sqrt((3 * math.pi) / 8) * dgd / math.sqrt(nplates)
        else:
            raise ValueError('Missing one of db0, theta or epsilon in fiber')
        brf['dgd'] = dgd
        dgdrms = dgdrms / G.SYMBOLRATE                                      # :284
    else:                                                                   # :290-298
        dgdrms, manakov, nplates = 0.0, False, 1
        brf['db0'] = np.zeros(1)
        brf['theta'] = np.zeros(1)
        brf['epsilon'] = np.zeros(1)

    c0 = CONSTANTS.CLIGHT
    lam = float(_get(x, 'lambda'))
    disp, slope = float(_get(x, 'disp')), float(_get(x, 'slope'))
    alphalin = (math.log(10) * 1e-4) * float(_get(x, 'alphadB'))            # :302
    b20 = -lam ** 2 / 2 / math.pi / c0 * disp * 1e-6                        # :308
    b30 = (lam / 2 / math.pi / c0) ** 2 * (2 * lam * disp + lam ** 2 * slope) * 1e-6
    b30 = b30 * fls[0]                                                      # :311
    lams = np.asarray(G.LAMBDA, dtype=np.float64).reshape(-1)
    maxl, minl = lams.max(), lams.min()
    lamc = 2 * maxl * minl / (maxl + minl)                                  # :315
    w_i0 = 2 * math.pi * c0 * (1.0 / lams - 1 / lam)
    w_ic = 2 * math.pi * c0 * (1.0 / lams - 1 / lamc)
    w_c0 = 2 * math.pi * c0 * (1.0 / lamc - 1 / lam)
    b1 = b20 * w_ic + 0.5 * b30 * (w_i0 ** 2 - w_c0 ** 2)                   # :321
    n2, aeff = float(_get(x, 'n2')), float(_get(x, 'aeff'))
    if nfc == 1:                                                            # :322-329
        beta1 = np.zeros(1)
        w_i0 = np.array([2 * math.pi * c0 * (1.0 / lamc - 1 / lam)])
        gam = np.array([2 * math.pi * n2 / (lamc * aeff) * 1e18])
    else:
        beta1 = b1
        gam = 2 * math.pi * n2 / (lams * aeff) * 1e18
    beta2 = (b20 + b30 * w_i0) * fls[0]                                     # :330-332
    dch = disp + slope * (lams - lam)                                       # :336

    betat, db1 = _dispersion_vectors(G, nfft, nfc, beta1, beta2, b30, dgdrms, fls[1])
    return FiberSetup(nfft=nfft, nfc=nfc, fls=fls, dphimaxt=dphimaxt, dzmaxt=dzmaxt, length=length,
                      alphalin=alphalin, gam=gam, betat=betat, db1=db1, manakov=manakov,
                      nplates=nplates, brf=brf, isv=isv, isy=isy, b1=b1, dch=dch,
                      scalars=dict(nsymb=G.NSYMB, nt=G.NT, symbolrate=float(G.SYMBOLRATE), b30=float(b30),
                                   dgdrms=float(dgdrms), beta1=np.asarray(beta1, dtype=np.float64),
                                   beta2=np.asarray(beta2, dtype=np.float64)),
                      tolflag=tolflag, trg=trg)


DISP_MODE = 'scalar'   # 'scalar': only the field crosses PCIe; 'vector': betat/db1 are uploaded


PRECISION = 'f64'   # arithmetic of the device path: 'f64' (default, the reference's) or 'f32' (reported separately)


def setup_to_desc(s: FiberSetup, batch=1, plate_sets=1, db0=None, theta=None, epsilon=None, disp_mode=None,
                  precision=None, z_start=0.0, dz_first=0.0):
    """FiberSetup -> (pmx_fiber_desc, keep-alive dict)."""
    mode = disp_mode or DISP_MODE
    prec = {'f64': _lib.PMX_F64, 'f32': _lib.PMX_F32}[precision or PRECISION]
    return _lib.make_desc(
        s.nfft, s.nfc, batch, s.length, s.alphalin, s.dzmaxt, s.dphimaxt, s.gam, s.fls, s.manakov, s.nplates,
        s.brf['db0'] if db0 is None else db0, s.brf['theta'] if theta is None else theta,
        s.brf['epsilon'] if epsilon is None else epsilon, s.betat, s.db1 if s.fls[1] else None,
        plate_sets=plate_sets, precision=prec, scalar=s.scalars if mode == 'scalar' else None,
        scalar_field=not s.isv, z_start=z_start, dz_first=dz_first)


def apply_side_effects(s: FiberSetup):
    """GSTATE.DELAY / GSTATE.DISP bookkeeping, fiber.m:367-369."""
    G = GSTATE
    rows = 2 if s.isy else 1
    loc_delay = s.length * G.SYMBOLRATE * s.b1
    G.DELAY = G.DELAY + np.ones((rows, 1)) * loc_delay[None, :]
    G.DISP = G.DISP + np.ones((rows, 1)) * (s.fls[0] * s.dch * s.length * 1e-3)[None, :]


LAST = {}  # firstdz / ncycle / schedule of the most recent fiber(), what the reference prints to simul_out


def _nextstep_host(dzmax, phimax, gam, alphalin, umax):   # (kept for host-side checks; the library has its own)
    """nextstep (fiber.m:693-715) from the per-column maxima of |u|^2."""
    with np.errstate(divide='ignore', invalid='ignore'):
        pmax = np.max(np.asarray(gam, dtype=np.float64) * np.asarray(umax, dtype=np.float64))
        leff = np.float64(phimax) / pmax
        dl = np.float64(alphalin) * leff
        if dl >= 1:
            return float(dzmax)
        step = leff if alphalin == 0 else np.float64(-1.0) / np.float64(alphalin) * np.log(np.float64(1.0) - dl)
        return float(dzmax) if step > dzmax else float(step)


def _scalar_adaptive(s: FiberSetup, ctx: _lib.Context, disp_mode=None):
    """The local-error adaptive dispatches of the scalar path -- scalar_a_ssfm / adaptssfm (x.ltol, fiber.m:639-679,
    938-1010) and scalar_ssfm with x.dphiadapt (tolflag 1, :588-611) -- through pmx_scalar_adaptive_run: the
    accept/reject loop is host logic inside the library, nl_step, lin_step, the error norm and the Richardson
    combination run on the resident field.  -> (firstdz, ncycle)"""
    import ctypes
    G = GSTATE
    desc, keep = setup_to_desc(s, disp_mode=disp_mode)
    fx = np.ascontiguousarray(np.asarray(G.FIELDX, dtype=np.complex128).T)[None]          # [1][nfc][nfft]
    io = _lib.complex_field(fx, None)
    res = _lib.Result(1)
    ctx.check(ctx.lib.pmx_scalar_adaptive_run(ctx.h, ctypes.byref(desc), float(s.trg['err']), float(s.trg['safety']),
                                              1 if s.tolflag == 1 else 0, ctypes.byref(io), ctypes.byref(res.c)))
    G.FIELDX = np.ascontiguousarray(fx[0].T)
    return float(res.firstdz[0]), int(res.ncycle[0])


def fiber(x, flag: str, rng: Optional[np.random.Generator] = None, ctx: Optional[_lib.Context] = None,
          trace: bool = False, disp_mode: Optional[str] = None, precision: Optional[str] = None):
    """zbrf = fiber(x, flag) -- fiber.m:1.  Propagates GSTATE.FIELDX/FIELDY in place.
    precision: 'f64' (default) or 'f32' -- arithmetic of the device path; host arrays stay complex128."""
    G = GSTATE
    s = fiber_setup(x, flag, rng)
    if s.fls[3] and s.isv:
        # matrix_nl_step raises at fiber.m:854 on the first step
        raise NotImplementedError('The CNLSE with separate fields is not yet implemented')
    if s.fls[1] and not s.isy:                                              # :285-289
        G.FIELDY = np.zeros_like(G.FIELDX)
    if s.tolflag == 2 and s.isv:                                            # :372-374
        raise ValueError('adaptive step available in absence of polarization effects')
    ctx = ctx or _lib.default_context()
    # (GSTATE.DELAY / DISP, fiber.m:367-369, are updated once the propagation has succeeded)
    # x.dphiadapt (tolflag 1) is read by scalar_ssfm only (fiber.m:386-387, :588); matrix_ssfm takes no tolflag
    if s.tolflag == 2 or (s.tolflag == 1 and not s.isv):                    # :375-380, :386-387
        if (precision or PRECISION) != 'f64':
            raise NotImplementedError('the local-error adaptive step runs in FP64 only')
        firstdz, ncycle = _scalar_adaptive(s, ctx, disp_mode)
        apply_side_effects(s)
        LAST.clear()
        LAST.update(firstdz=firstdz, ncycle=ncycle, ntot=0)
        simul_out.log_fiber(x, flag, s, firstdz, ncycle)                    # :392-456
        return None
    desc, keep = setup_to_desc(s, disp_mode=disp_mode, precision=precision)
    scalar = not s.isv
    if scalar:
        # single-polarization path (fiber.m:372-380): one call on host buffers, H2D + loop + D2H
        fx = np.ascontiguousarray(np.asarray(G.FIELDX, dtype=np.complex128).T)[None]     # [1][nfc][nfft]
        fy = np.zeros_like(fx)
        io = _lib.complex_field(fx, fy)
        res = _lib.Result(1, trace_cap=4096 if trace else 0)
        import ctypes
        ctx.check(ctx.lib.pmx_fiber_run(ctx.h, ctypes.byref(desc), ctypes.byref(io), ctypes.byref(res.c)))
        G.FIELDX = np.ascontiguousarray(fx[0].T)
    else:
        # two polarizations: the field of the previous in-line device if it is still in HBM, else an upload; the
        # result stays there until GSTATE.FIELDX / FIELDY are read (gstate.RESIDENT)
        fld, hx, hy = G.take_device(ctx, desc.precision)
        try:
            plan = _lib.Plan(ctx, desc, keep)
        except Exception:
            G.restore_host(fld, hx, hy)     # nothing ran: the caller's field is as it was
            raise
        try:
            res = plan.execute(fld, trace_cap=4096 if trace else 0)
        except Exception:
            fld.close()                     # the device copy is part-way through the fiber: not usable
            raise
        finally:
            plan.close()
        G.put_device(fld, hx, hy)
    apply_side_effects(s)
    LAST.clear()
    LAST.update(firstdz=float(res.firstdz[0]), ncycle=int(res.ncycle[0]), ntot=int(res.ntot[0]))
    if trace:
        dz, nt = res.schedule(0)
        LAST.update(trace_dz=dz, trace_ntrunk=nt)
    simul_out.log_fiber(x, flag, s, LAST['firstdz'], LAST['ncycle'])        # :392-456
    if not s.isv:
        return None
    brf = dict(s.brf)
    brf['lcorr'] = s.length / s.nplates                                     # :510
    brf['betat'] = s.betat                                                  # :553
    brf['db1'] = s.db1                                                      # :554
    return brf
