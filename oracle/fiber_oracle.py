"""NumPy restatement of the reference's SSFM fiber channel (TEST INFRASTRUCTURE).

This file restates, operation for operation, the arithmetic of the reference
M-code on the hot path.  It is the checker the CUDA path is compared with; it
is never imported by the product package.

Reference files followed (all under /root/reference):
  reset_all.m:105-112,152-174   CONSTANTS, GSTATE layout, FN grid
  create_field.m:113-124,156-199  power normalisation, 'sepfields', 'unique' mux
  fastshift.m:45-58             circular shift used by the mux
  fiber.m:126-389               front-end (flags, PMD setup, unit conversions)
  fiber.m:459-555               matrix_ssfm
  fiber.m:557-636               scalar_ssfm
  fiber.m:639-679,938-1010      scalar_a_ssfm, adaptssfm
  fiber.m:682-715               nextstep
  fiber.m:718-758               checkstep
  fiber.m:762-804               lin_step, nl_step
  fiber.m:807-874               matrix_nl_step
  fiber.m:877-935               matrix_step
  fastexp.m:28 / fastexp.c:37-44  exp(i*x) = cos(x) + i sin(x)
  ampliflat.m:61-148            flat gain + ASE (options.noise hook :123-129)
  inverse_pmd.m:91-168          inverse Jones matrix (self-check of the PMD step)
  ber_estimate.m:118            integer error count

Third-party arithmetic the reference relies on and that is NOT in the tree:
MATLAB/Octave builtins fft/ifft (FFTW, version unpinned), cos/sin/exp/log
(libm), rand/randn (legacy 'state' streams).  Here: scipy.fft (pocketfft),
numpy ufuncs, numpy Generator streams passed in by the caller.

PARITY PINNING.  The reference ships no golden vectors, no known-answer tests
and no fixtures for this path, and neither Octave nor
MATLAB exists in the build image.  The restatement is pinned two ways:
  (1) tests/golden/ holds outputs produced by executing the reference's own
      fiber.m source text from /root/reference with the minimal M-interpreter
      in oracle/mini_m (script committed: oracle/make_golden.py); the tests
      check this restatement against those fixtures;
  (2) the structural self-checks of SURVEY.md section 4 (linear closed forms,
      exact SPM solution, energy ratio, inverse_pmd round trip, trunk-count
      invariants) and an extended-precision (np.longdouble) instantiation of
      the same code as arbiter.
If (1) is missing from the tree, treat the oracle as "parity unpinned".

Everything is written for a ``real`` dtype chosen by the caller: np.float64
(default, the reference's arithmetic), np.float32 (FP32-mode checker) or
np.longdouble (arbiter).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field as _dc_field
from typing import Optional

import numpy as np
import scipy.fft as _sfft

# reset_all.m:105-112
CLIGHT = 299792458.0
HPLANCK = 6.62606896e-34
ECHARGE = 1.602176487e-19
KBOLTZMANN = 1.3806504e-23

SAFETYFCT = 0.9   # fiber.m:130
DEF_PLATES = 100  # fiber.m:131


def _ctype(real):
    return {np.dtype(np.float32): np.complex64,
            np.dtype(np.float64): np.complex128,
            np.dtype(np.longdouble): np.clongdouble}[np.dtype(real)]


@dataclass
class GState:
    """reset_all.m:152-174 (the fields the path reads or writes)."""
    NSYMB: int
    NT: int
    NCH: int
    FN: np.ndarray                      # [N], FFT order
    SYMBOLRATE: Optional[float] = None  # [GBaud], electricsource sets it
    LAMBDA: Optional[np.ndarray] = None  # [NCH] nm
    POWER: Optional[np.ndarray] = None   # [NCH] mW
    FIELDX: Optional[np.ndarray] = None  # [N, nfc] complex
    FIELDY: Optional[np.ndarray] = None
    FIELDX_TX: Optional[np.ndarray] = None
    FIELDY_TX: Optional[np.ndarray] = None
    DELAY: Optional[np.ndarray] = None   # [npol, NCH]
    DISP: Optional[np.ndarray] = None
    PRINT: bool = False
    real: type = np.float64
    log: dict = _dc_field(default_factory=dict)  # firstdz / ncycle / schedule of the last fiber()


def reset_all(nsymb: int, nt: int, nch: int, real=np.float64) -> GState:
    """reset_all.m:152-156: FN = fftshift(-Nt/2 : 1/Nsymb : Nt/2 - 1/Nsymb)."""
    n = nsymb * nt
    stepf = 1.0 / nsymb
    # MATLAB colon: a + k*step (exact here whenever 1/Nsymb is a power of two)
    fn = np.fft.fftshift(-nt / 2.0 + np.arange(n, dtype=np.float64) * stepf)
    return GState(NSYMB=nsymb, NT=nt, NCH=nch, FN=fn.astype(real), real=real)


# --------------------------------------------------------------------------
# helpers
def fastexp(x):
    """fastexp.m:28 / fastexp.c:37-44: cos(x) + i*sin(x) for real x."""
    out = np.empty(np.shape(x), dtype=_ctype(np.asarray(x).dtype))
    out.real = np.cos(x)
    out.imag = np.sin(x)
    return out


def fft(u):
    """Column-wise forward FFT, MATLAB convention exp(-2*pi*i*k*n/N)."""
    return _sfft.fft(u, axis=0)


def ifft(u):
    """Column-wise inverse FFT, scaled 1/N."""
    return _sfft.ifft(u, axis=0)


def fastshift(x, n: int):
    """fastshift.m:45-58: y(k) = x(k-n) circularly (n>0 delays), along axis 0."""
    return np.roll(x, int(n), axis=0)


# --------------------------------------------------------------------------
def create_field(gs: GState, ftype: str, sigx, sigy=None, power_average=False):
    """create_field.m:79-202 without the 'delay' option.

    sigx, sigy: [N, NCH] complex.  ``power_average`` is options.power='average'
    (:113-124).  'sepfields' copies (:156-162); 'unique' multiplexes in the
    frequency domain (:164-199).
    """
    real = gs.real
    ct = _ctype(real)
    sigx = np.array(sigx, dtype=ct, copy=True).reshape(gs.NSYMB * gs.NT, -1)
    isy = sigy is not None and np.size(sigy) != 0
    if isy:
        sigy = np.array(sigy, dtype=ct, copy=True).reshape(gs.NSYMB * gs.NT, -1)
        if sigy.shape != sigx.shape:
            raise ValueError('sigx and sigy must have the same size')
    if sigx.shape[1] != gs.NCH:
        raise ValueError('the number of columns of sigx,sigy must be equal to the number of channels')
    npol = 2 if isy else 1
    power = np.asarray(gs.POWER, dtype=real).reshape(-1)
    if power_average:                                    # :113-124
        if isy:
            avge = np.mean(np.abs(sigx) ** 2 + np.abs(sigy) ** 2, axis=0)
        else:
            avge = np.mean(np.abs(sigx) ** 2, axis=0)
        scale = np.sqrt(power / avge)
        sigx = sigx * scale[None, :]
        if isy:
            sigy = sigy * scale[None, :]
        gs.POWER = power * power / avge
    gs.DELAY = np.zeros((npol, gs.NCH))                  # :146
    gs.DISP = np.zeros((npol, gs.NCH))                   # :151
    ftype = ftype.lower()
    if ftype == 'sepfields':                             # :156-162
        gs.FIELDX_TX = sigx.copy()
        gs.FIELDX = sigx
        if isy:
            gs.FIELDY_TX = sigy.copy()
            gs.FIELDY = sigy
        else:
            gs.FIELDY = None
    elif ftype == 'unique':                              # :164-199
        lamt = np.asarray(gs.LAMBDA, dtype=np.float64).reshape(-1)
        maxl, minl = lamt.max(), lamt.min()
        lamc = 2 * maxl * minl / (maxl + minl)
        deltafn = CLIGHT * (1 / lamc - 1.0 / lamt)
        minfreq = float(gs.FN[1] - gs.FN[0])
        ndfn = np.round(deltafn / gs.SYMBOLRATE / minfreq).astype(np.int64)
        n = gs.NSYMB * gs.NT
        fx = np.zeros((n, 1), dtype=ct)
        zx = fft(sigx)
        fy = np.zeros((n, 1), dtype=ct) if isy else None
        for kch in range(gs.NCH):
            fx[:, 0] = fx[:, 0] + fastshift(zx[:, kch], -ndfn[kch])
            if isy:
                zy = fft(sigy[:, kch])
                fy[:, 0] = fy[:, 0] + fastshift(zy, -ndfn[kch])
        gs.FIELDX = ifft(fx)
        gs.FIELDX_TX = gs.FIELDX.copy()
        if isy:
            gs.FIELDY = ifft(fy)
            gs.FIELDY_TX = gs.FIELDY.copy()
        else:
            gs.FIELDY = None
    else:
        raise ValueError("ftype must be 'sepfields' or 'unique'")


# --------------------------------------------------------------------------
_FLAGS = {
    # flag: (fls, nonlinear-step rule: 'lin' | 'nl' | 'nl_if_multi', needs nfc != 1)
    '----': ((0, 0, 0, 0), 'lin', False),
    'g---': ((1, 0, 0, 0), 'lin', False),
    '-p--': ((0, 1, 0, 0), 'lin', False),
    '--s-': ((0, 0, 1, 0), 'nl_if_multi', False),
    '---x': ((0, 0, 0, 1), 'nl', True),
    'gp--': ((1, 1, 0, 0), 'lin', False),
    'g-s-': ((1, 0, 1, 0), 'nl', False),
    'g--x': ((1, 0, 0, 1), 'nl', True),
    '-ps-': ((0, 1, 1, 0), 'nl', False),
    '-p-x': ((0, 1, 0, 1), 'nl', True),
    '--sx': ((0, 0, 1, 0), 'nl_if_multi', False),   # x added when nfc != 1
    'g-sx': ((1, 0, 1, 0), 'nl', False),
    '-psx': ((0, 1, 1, 0), 'nl', False),
    'gps-': ((1, 1, 1, 0), 'nl', False),
    'gp-x': ((1, 1, 0, 1), 'nl', True),
    'gpsx': ((1, 1, 1, 0), 'nl', False),
}


def parse_flag(flag: str, nfc: int, x: dict):
    """fiber.m:157-251 -> (fls, dphimaxt, dzmaxt)."""
    f = flag.lower()
    if f not in _FLAGS:
        raise ValueError("wrong flag. E.g. flag can be 'g---','gp--','-s--', etc")
    fls, rule, need_multi = _FLAGS[f]
    fls = list(fls)
    if need_multi and nfc == 1:
        raise ValueError("flag '%s' available only for channels separated" % f)
    if f in ('--sx', 'g-sx', '-psx', 'gpsx') and nfc != 1:
        fls[3] = 1
    if rule == 'lin' or (rule == 'nl_if_multi' and nfc == 1):
        dphimaxt, dzmaxt = math.inf, x['length']
    else:
        dphimaxt, dzmaxt = x['dphimax'], x['dzmax']
    return fls, dphimaxt, dzmaxt


def fiber(gs: GState, x: dict, flag: str, rng: Optional[np.random.Generator] = None):
    """fiber.m:126-389.  Mutates gs.FIELDX/FIELDY/DELAY/DISP; returns brf (dict)
    for two-polarization runs (fiber.m:384) else None.  ``rng`` replaces the
    interpreter's global rand stream for the random-plate draw (:274-276)."""
    real = gs.real
    x = dict(x)
    nfr, nfc = gs.FIELDX.shape
    nfft = gs.NSYMB * gs.NT
    if 'dzmax' not in x or x['dzmax'] > x['length']:      # :139-141
        x['dzmax'] = x['length']
    if 'ltol' in x:                                       # :143-155
        if 'dphimax' not in x:
            x['dphimax'] = math.inf
        trg = {'err': x['ltol'], 'safety': SAFETYFCT}
        tolflag = 1 if x.get('dphiadapt', False) else 2
    else:
        trg, tolflag = None, 0
    fls, dphimaxt, dzmaxt = parse_flag(flag, nfc, x)

    isy = gs.FIELDY is not None and np.size(gs.FIELDY) != 0  # :253
    isv = fls[1] == 1 or isy
    brf = {}
    if fls[1] == 1:                                       # :255-289
        if 'manakov' not in x:
            x['manakov'] = 'no'
        if 'dgd' not in x:
            raise ValueError('Missing DGD in fiber')
        ispmf = ('db0' in x) + ('theta' in x) + ('epsilon' in x)
        if ispmf == 3:
            theta = np.atleast_1d(np.asarray(x['theta'], dtype=np.float64))
            x['nplates'] = len(theta)
            brf['db0'] = np.atleast_1d(np.asarray(x['db0'], dtype=np.float64))
            brf['theta'] = theta
            brf['epsilon'] = np.atleast_1d(np.asarray(x['epsilon'], dtype=np.float64))
            dgdrms = x['dgd'] / x['nplates']
        elif ispmf == 0:
            if 'nplates' not in x:
                x['nplates'] = DEF_PLATES
            if rng is None:
                raise ValueError('random plates need an rng')
            npl = int(x['nplates'])
            brf['db0'] = rng.random(npl) * 2 * np.pi - np.pi
            brf['theta'] = rng.random(npl) * np.pi - 0.5 * np.pi
            brf['epsilon'] = 0.5 * np.arcsin(rng.random(npl) * 2 - 1)
            dgdrms = math.sqrt((3 * math.pi) / 8) * x['dgd'] / math.sqrt(x['nplates'])
        else:
            raise ValueError('Missing one of db0, theta or epsilon in fiber')
        brf['dgd'] = x['dgd']
        dgdrms = dgdrms / gs.SYMBOLRATE
        if not isy:
            gs.FIELDY = np.zeros((nfr, nfc), dtype=_ctype(real))
    else:                                                 # :290-298
        dgdrms = 0.0
        brf['db0'] = np.zeros(1)
        brf['theta'] = np.zeros(1)
        brf['epsilon'] = np.zeros(1)
        x['manakov'] = 'no'
        x['nplates'] = 1

    # conversions :302-362 (host scalars, always IEEE double like the reference)
    alphalin = (math.log(10) * 1e-4) * x['alphadB']
    lam = x['lambda']
    b20 = -lam ** 2 / 2 / math.pi / CLIGHT * x['disp'] * 1e-6
    b30 = (lam / 2 / math.pi / CLIGHT) ** 2 * (2 * lam * x['disp'] + lam ** 2 * x['slope']) * 1e-6
    b30 = b30 * fls[0]
    lambdas = np.asarray(gs.LAMBDA, dtype=np.float64).reshape(-1)
    maxl, minl = lambdas.max(), lambdas.min()
    lamc = 2 * maxl * minl / (maxl + minl)
    domega_i0 = 2 * math.pi * CLIGHT * (1.0 / lambdas - 1 / lam)
    domega_ic = 2 * math.pi * CLIGHT * (1.0 / lambdas - 1 / lamc)
    domega_c0 = 2 * math.pi * CLIGHT * (1.0 / lamc - 1 / lam)
    b1 = b20 * domega_ic + 0.5 * b30 * (domega_i0 ** 2 - domega_c0 ** 2)
    if nfc == 1:
        beta1 = np.zeros(1)
        domega_i0 = np.array([2 * math.pi * CLIGHT * (1.0 / lamc - 1 / lam)])
        gam = np.array([2 * math.pi * x['n2'] / (lamc * x['aeff']) * 1e18])
    else:
        beta1 = b1
        gam = 2 * math.pi * x['n2'] / (lambdas * x['aeff']) * 1e18
    beta2 = b20 + b30 * domega_i0
    beta2 = beta2 * fls[0]
    dch = x['disp'] + x['slope'] * (lambdas - lam)

    omega = 2 * math.pi * gs.SYMBOLRATE * np.asarray(gs.FN, dtype=np.float64)   # :352
    betat = np.zeros((nfft, nfc))
    db1 = np.zeros((nfft, nfc))
    for kch in range(nfc):                                # :354-362
        betat[:, kch] = omega * beta1[kch] + 0.5 * omega ** 2 * beta2[kch] + omega ** 3 * b30 / 6
        if fls[1] == 1:
            db1[:, kch] = dgdrms * omega

    # :367-369 (isy is the value from before FIELDY was auto-created)
    loc_delay = x['length'] * gs.SYMBOLRATE * b1
    rows = 2 if isy else 1
    gs.DELAY = gs.DELAY + np.ones((rows, 1)) * loc_delay[None, :]
    gs.DISP = gs.DISP + np.ones((rows, 1)) * (fls[0] * dch * x['length'] * 1e-3)[None, :]

    betat_r = betat.astype(real)
    db1_r = db1.astype(real)
    if tolflag == 2:                                      # :372-378
        if isv:
            raise ValueError('adaptive step available in absence of polarization effects')
        firstdz, ncycle, gs.FIELDX = scalar_a_ssfm(gs.FIELDX, betat_r, dzmaxt, dphimaxt, gam, alphalin,
                                                   nfft, nfc, x['length'], trg, fls, real)
        gs.log = {'firstdz': firstdz, 'ncycle': ncycle}
        return None
    if isv:                                               # :380-384
        brf_in = {k: np.asarray(v, dtype=np.float64) if k != 'dgd' else v for k, v in brf.items()}
        firstdz, ncycle, gs.FIELDX, gs.FIELDY, brf, sched = matrix_ssfm(
            gs.FIELDX, gs.FIELDY, betat_r, db1_r, dzmaxt, dphimaxt, gam, alphalin, nfc,
            x['length'], int(x['nplates']), x['manakov'], fls, brf_in, real)
        brf['betat'] = betat
        brf['db1'] = db1
        gs.log = {'firstdz': firstdz, 'ncycle': ncycle, 'schedule': sched,
                  'gam': gam, 'alphalin': alphalin}
        return brf
    firstdz, ncycle, gs.FIELDX = scalar_ssfm(gs.FIELDX, betat_r, dzmaxt, dphimaxt, gam, alphalin,
                                             nfft, nfc, x['length'], fls, tolflag, trg, real)
    gs.log = {'firstdz': firstdz, 'ncycle': ncycle, 'gam': gam, 'alphalin': alphalin}
    return None


# --------------------------------------------------------------------------
def nextstep(dzmax, phimax, gam, alphalin, ux, uy, isv, real=np.float64):
    """fiber.m:693-715.  Returns dz (python float / real scalar)."""
    if isv:
        umax = np.max(ux.real ** 2 + ux.imag ** 2 + uy.real ** 2 + uy.imag ** 2, axis=0)
    else:
        umax = np.max(ux.real ** 2 + ux.imag ** 2, axis=0)
    R = np.dtype(real).type
    with np.errstate(divide='ignore', invalid='ignore'):
        pmax = np.max(np.asarray(gam, dtype=real) * umax.astype(real))
        leff = R(phimax) / pmax
        dl = R(alphalin) * leff
        if dl >= 1:
            return R(dzmax)
        if alphalin == 0:
            step = leff
        else:
            step = R(-1) / R(alphalin) * np.log(R(1) - dl)
        if step > dzmax:
            return R(dzmax)
        return step


def checkstep(zprop, dz, lcorr, dz_miss, nz_old):
    """fiber.m:739-758 -> (dzb list, dz_miss, nmem, ntrunk)."""
    nz = zprop / lcorr
    nzc = int(math.ceil(nz))
    if dz_miss == 0:
        nmem = 0
        ntrunk = nzc - nz_old
        dzlast = dz - lcorr * (ntrunk - 1)
        dzb = [lcorr] * (ntrunk - 1) + [dzlast]
        dz_miss = lcorr - dzlast
    else:
        nmem = 1
        ntrunk = nzc - nz_old + 1
        if ntrunk == 1:
            dzb = [dz]
            dz_miss = dz_miss - dz
        else:
            dzlast = dz - dz_miss - lcorr * (ntrunk - 2)
            dzb = [dz_miss] + [lcorr] * (ntrunk - 2) + [dzlast]
            dz_miss = lcorr - dzlast
    return dzb, dz_miss, nmem, ntrunk


def _leff(alphalin, dz, real):
    R = np.dtype(real).type
    if alphalin == 0:
        return R(dz)
    return (R(1) - np.exp(-R(alphalin) * R(dz))) / R(alphalin)


def matrix_nl_step(ismanakov, alphalin, gam, dz, ux, uy, nfc, spm, xpm, real=np.float64):
    """fiber.m:827-874."""
    leff = _leff(alphalin, dz, real)
    R = np.dtype(real).type
    for kch in range(nfc):
        if spm:
            power = (ux[:, kch].real ** 2 + ux[:, kch].imag ** 2 +
                     uy[:, kch].real ** 2 + uy[:, kch].imag ** 2)
            gamleff = R(gam[kch]) * leff
            nlscalar = fastexp(-gamleff * power)
            ux[:, kch] = ux[:, kch] * nlscalar
            uy[:, kch] = uy[:, kch] * nlscalar
            if not ismanakov:
                s3 = 2 * (ux[:, kch].real * uy[:, kch].imag - ux[:, kch].imag * uy[:, kch].real)
                e = fastexp(gamleff * s3 / 3)
                cosphi, sinphi = e.real, e.imag
                uxx = cosphi * ux[:, kch] + sinphi * uy[:, kch]
                uyy = -sinphi * ux[:, kch] + cosphi * uy[:, kch]
                ux[:, kch] = uxx
                uy[:, kch] = uyy
        if xpm:
            raise ValueError('The CNLSE with separate fields is not yet implemented')  # :854
    return ux, uy


def _mat_r(theta, epsilon, real):
    """fiber.m:910-912: matR = (cos(th)*I - sin(th)*[0 1;-1 0]) * complex(cos(e)*I, sin(e)*[0 1;1 0])."""
    ct = _ctype(real)
    c, s = math.cos(theta), math.sin(theta)
    ce, se = math.cos(epsilon), math.sin(epsilon)
    rth = np.array([[c, -s], [s, c]], dtype=real)
    rep = np.array([[ce, 1j * se], [1j * se, ce]], dtype=ct)
    return rth.astype(ct) @ rep


def matrix_step(betat, db1, dzb, ntrunk, ux, uy, brf, ntot, nmem, lcorr, real=np.float64):
    """fiber.m:904-935."""
    R = np.dtype(real).type
    ux = fft(ux)
    uy = fft(uy)
    for k in range(1, ntrunk + 1):
        n = ntot + k - nmem                       # 1-based trunk number
        if n > len(brf['theta']):
            raise IndexError('trunk index %d exceeds nplates %d (fiber.m:910 index error)'
                             % (n, len(brf['theta'])))
        m = _mat_r(float(brf['theta'][n - 1]), float(brf['epsilon'][n - 1]), real)
        uux = np.conj(m[0, 0]) * ux + np.conj(m[1, 0]) * uy
        uuy = np.conj(m[0, 1]) * ux + np.conj(m[1, 1]) * uy
        combeta = betat * R(dzb[k - 1])
        deltabeta = R(0.5) * (db1 + R(brf['db0'][n - 1])) * R(dzb[k - 1]) / R(lcorr)
        uux = fastexp(-(combeta + deltabeta)) * uux
        uuy = fastexp(-(combeta - deltabeta)) * uuy
        ux = m[0, 0] * uux + m[0, 1] * uuy
        uy = m[1, 0] * uux + m[1, 1] * uuy
    ux = ifft(ux)
    uy = ifft(uy)
    return ux, uy


def matrix_ssfm(ux, uy, betat, db1, dzmaxt, dphimaxt, gam, alphalin, nfc, lf, nplates,
                manakov, fls, brf, real=np.float64):
    """fiber.m:498-554.  Also returns the per-step schedule for test comparison."""
    R = np.dtype(real).type
    ct = _ctype(real)
    ux = np.array(ux, dtype=ct, copy=True)
    uy = np.array(uy, dtype=ct, copy=True)
    gam = np.asarray(gam, dtype=np.float64).copy()
    if manakov == 'yes':
        gam = gam * 8 / 9
        ismanakov = True
    else:
        ismanakov = False
    ncycle = 1
    lcorr = lf / nplates
    dz_miss = 0.0
    brf = dict(brf)
    brf['lcorr'] = lcorr
    dz = float(nextstep(dzmaxt, dphimaxt, gam, alphalin, ux, uy, True, real))
    halfalpha = 0.5 * alphalin
    ntot = 0
    firstdz = dz
    zprop = dz
    sched = []
    while zprop < lf:
        ux, uy = matrix_nl_step(ismanakov, alphalin, gam, dz, ux, uy, nfc, fls[2], fls[3], real)
        dzb, dz_miss, nmem, ntrunk = checkstep(zprop, dz, lcorr, dz_miss, ntot)
        sched.append({'dz': dz, 'zprop': zprop, 'ntrunk': ntrunk, 'nmem': nmem, 'ntot': ntot, 'dzb': list(dzb)})
        ux, uy = matrix_step(betat, db1, dzb, ntrunk, ux, uy, brf, ntot, nmem, lcorr, real)
        ntot = ntot + ntrunk - nmem
        a = np.exp(R(-halfalpha * dz))
        ux = ux * a
        uy = uy * a
        dz = float(nextstep(dzmaxt, dphimaxt, gam, alphalin, ux, uy, True, real))
        zprop = zprop + dz
        ncycle += 1
    last_step = lf - zprop + dz
    ux, uy = matrix_nl_step(ismanakov, alphalin, gam, last_step, ux, uy, nfc, fls[2], fls[3], real)
    dzb, dz_miss, nmem, ntrunk = checkstep(lf, last_step, lcorr, dz_miss, ntot)
    sched.append({'dz': last_step, 'zprop': lf, 'ntrunk': ntrunk, 'nmem': nmem, 'ntot': ntot, 'dzb': list(dzb)})
    ux, uy = matrix_step(betat, db1, dzb, ntrunk, ux, uy, brf, ntot, nmem, lcorr, real)
    a = np.exp(R(-halfalpha * last_step))
    ux = ux * a
    uy = uy * a
    return firstdz, ncycle, ux, uy, brf, sched


# --------------------------------------------------------------------------
def lin_step(betaxdz, u):
    """fiber.m:771-773."""
    return ifft(fft(u) * fastexp(-betaxdz))


def nl_step(alphalin, gam, dz, u, nfc, spm, xpm, real=np.float64):
    """fiber.m:787-804.  gam is the Nfft x nfc replicated matrix (or broadcastable row)."""
    leff = _leff(alphalin, dz, real)
    powr = u.real ** 2 + u.imag ** 2
    if xpm:
        tot = np.sum(powr, axis=1, keepdims=True) * np.ones((1, nfc), dtype=powr.dtype)
        if spm:
            powr = 2 * tot - powr
        else:
            powr = 2 * (tot - powr)
    elif not spm:
        return u
    return u * fastexp(-gam * powr * leff)


def scalar_ssfm(u, betat, dzmaxt, dphimaxt, gam, alphalin, nfft, nfc, lf, fls, tolflag, trg,
                real=np.float64):
    """fiber.m:584-636."""
    R = np.dtype(real).type
    u = np.array(u, dtype=_ctype(real), copy=True)
    gam = np.asarray(gam, dtype=np.float64)
    gamrep = np.asarray(gam, dtype=real).reshape(1, -1) * np.ones((nfft, 1), dtype=real)
    dz = float(nextstep(dzmaxt, dphimaxt, gam, alphalin, u, None, False, real))
    halfalpha = 0.5 * alphalin
    ncycle = 1
    if tolflag == 1:                                  # :588-611
        if dz >= dzmaxt:
            umax = np.max(u.real ** 2 + u.imag ** 2, axis=0)
            maxpow = float(np.max(gam * umax))
            dphimaxt = maxpow * dz if alphalin == 0 else maxpow * (1 - math.exp(-alphalin * dz)) / alphalin
        dzini = dz
        zdone = 0.0
        while zdone == 0:
            u, ncycle, nrej, zdone, dz = adaptssfm(u, zdone, dz, alphalin, gamrep, nfc, fls, betat,
                                                   halfalpha, trg, 0, 0, real)
        if dz > dzmaxt:
            dz = dzmaxt
        dphimaxt = dphimaxt * (1 - math.exp(-alphalin * zdone)) / (1 - math.exp(-alphalin * dzini))
        firstdz = zdone
        zprop = zdone + dz
        ncycle += 1
    else:
        firstdz = dz
        zprop = dz
    while zprop < lf:
        u = nl_step(alphalin, gamrep, dz, u, nfc, fls[2], fls[3], real)
        u = lin_step(betat * R(dz), u)
        u = u * np.exp(R(-halfalpha * dz))
        dz = float(nextstep(dzmaxt, dphimaxt, gam, alphalin, u, None, False, real))
        zprop = zprop + dz
        ncycle += 1
    last_step = lf - zprop + dz
    u = nl_step(alphalin, gamrep, last_step, u, nfc, fls[2], fls[3], real)
    u = lin_step(betat * R(last_step), u)
    u = u * np.exp(R(-halfalpha * last_step))
    return firstdz, ncycle, u


def adaptssfm(u, zdone, dz, alphalin, gamrep, nfc, fls, betat, halfalpha, trg, nrej, ncycle,
              real=np.float64):
    """fiber.m:966-1009 (Richardson-extrapolated symmetric step)."""
    R = np.dtype(real).type
    dz2, dz4 = 0.5 * dz, 0.25 * dz
    ustack = u
    uh = u
    u = nl_step(alphalin, gamrep, dz2, u, nfc, fls[2], fls[3], real)
    u = u * np.exp(R(-halfalpha * dz2))
    u = lin_step(betat * R(dz), u)
    u = nl_step(alphalin, gamrep, dz2, u, nfc, fls[2], fls[3], real)
    u = u * np.exp(R(-halfalpha * dz2))
    uh = nl_step(alphalin, gamrep, dz4, uh, nfc, fls[2], fls[3], real)
    uh = uh * np.exp(R(-halfalpha * dz4))
    uh = lin_step(betat * R(dz2), uh)
    uh = nl_step(alphalin, gamrep, dz2, uh, nfc, fls[2], fls[3], real)
    uh = uh * np.exp(R(-halfalpha * dz2))
    uh = lin_step(betat * R(dz2), uh)
    uh = nl_step(alphalin, gamrep, dz4, uh, nfc, fls[2], fls[3], real)
    uh = uh * np.exp(R(-halfalpha * dz4))
    d = u - uh
    est_err = float(np.max(np.sqrt(d.real ** 2 + d.imag ** 2))) / dz
    if est_err > trg['err']:
        dz = trg['safety'] * math.sqrt(trg['err'] / est_err) * dz
        u = ustack
        nrej += 1
    else:
        u = R(4) / R(3) * uh - R(1) / R(3) * u
        zdone = zdone + dz
        dz = trg['safety'] * math.sqrt(trg['err'] / est_err) * dz
        ncycle += 1
    return u, ncycle, nrej, zdone, dz


def scalar_a_ssfm(u, betat, dzmaxt, dphimaxt, gam, alphalin, nfft, nfc, lf, trg, fls, real=np.float64):
    """fiber.m:664-679."""
    u = np.array(u, dtype=_ctype(real), copy=True)
    gam = np.asarray(gam, dtype=np.float64)
    gamrep = np.asarray(gam, dtype=real).reshape(1, -1) * np.ones((nfft, 1), dtype=real)
    ncycle = 1
    dz = float(nextstep(dzmaxt, dphimaxt, gam, alphalin, u, None, False, real))
    halfalpha = 0.5 * alphalin
    firstdz = dz
    zdone = 0.0
    nrej = 0
    while zdone < lf:
        if zdone + dz > lf:
            dz = lf - zdone
        u, ncycle, nrej, zdone, dz = adaptssfm(u, zdone, dz, alphalin, gamrep, nfc, fls, betat,
                                               halfalpha, trg, nrej, ncycle, real)
        if dz > dzmaxt:
            dz = dzmaxt
    return firstdz, ncycle, u


# --------------------------------------------------------------------------
def ampliflat_sigma(gs: GState, gain_db: float, f_db: float, nfc: int, gain=None):
    """ampliflat.m:62,91-106: linear gain (10^(gain_db/10) unless given) and ASE sigma [sqrt(mW)] per column."""
    if gain is None:
        gain = 10 ** (gain_db * 0.1)
    if f_db is None or math.isinf(f_db):
        return gain, np.zeros(nfc)
    flin = 10 ** (f_db * 0.1)
    lambdas = np.asarray(gs.LAMBDA, dtype=np.float64).reshape(-1)
    if nfc == 1:
        maxl, minl = lambdas.max(), lambdas.min()
        lamc = np.array([2 * maxl * minl / (maxl + minl)])
    else:
        lamc = lambdas
    sigma = np.sqrt(flin / 4 * HPLANCK * CLIGHT / lamc * (gain - 1) * gs.NT * gs.SYMBOLRATE * 1e21)
    return gain, sigma


def avg_power_abs(gs: GState, ich: int) -> float:
    """avg_power(ich,'abs') for separate channels (avg_power.m:63,83-86,90-98,110-132 with nfc == NCH: ndfn = 0,
    ndfnl = ndfnr = Nfft/2, no filter): E = sum |fft(FIELD(:,ich))|^2 / Nfft^2, X plus Y."""
    nfft, nfc = gs.FIELDX.shape
    if ich > gs.NCH:
        raise ValueError('The channel does not exist')
    if nfc != gs.NCH:
        raise NotImplementedError('avg_power: separate channels only')
    h = nfft // 2

    def one(col):
        x = np.fft.fft(np.asarray(col, dtype=np.complex128))
        return (np.sum(np.abs(x[:h]) ** 2) + np.sum(np.abs(x[nfft - h:]) ** 2)) / nfft ** 2

    e = one(gs.FIELDX[:, ich - 1])
    if gs.FIELDY is not None:
        e = e + one(gs.FIELDY[:, ich - 1])
    return float(e)


def ampliflat(gs: GState, gain_db: float, f_db: Optional[float] = None, noise=None, onepol=None, atype='gain'):
    """ampliflat.m:61-148.  atype 'gain': gain_db [dB]; 'fixpower': gain_db is the output power [mW] of the middle
    channel (:65-72).  ``noise`` is options.noise: [N, 2*nfc] complex standard normals (X columns first, then Y),
    :123-129; ``onepol``: 'asex' / 'asey' (:107-118)."""
    nfr, nfc = gs.FIELDX.shape
    atype = atype.lower()
    if atype == 'gain':
        gain, sigma = ampliflat_sigma(gs, gain_db, f_db, nfc)
    elif atype == 'fixpower':
        if nfc != gs.NCH:
            raise ValueError("'fixpower' works only for channels separated")
        gain = gain_db / avg_power_abs(gs, math.ceil(nfc / 2))
        gain, sigma = ampliflat_sigma(gs, 10 * math.log10(gain), f_db, nfc, gain=gain)
    else:
        raise ValueError('wrong string atype')
    R = np.dtype(gs.real).type
    sg = R(math.sqrt(gain))
    gs.FIELDX = gs.FIELDX * sg
    isy = gs.FIELDY is not None
    if isy:
        gs.FIELDY = gs.FIELDY * sg
    if np.any(sigma):
        if noise is None:
            raise ValueError('the oracle takes ASE noise from the caller (options.noise)')
        noise = np.asarray(noise)
        sig = sigma.astype(gs.real)[None, :]
        if onepol is not None and str(onepol).lower() not in ('asex', 'asey'):
            raise ValueError("ONEPOL, if exists, must be 'asex' or 'asey'")
        asepol = (True, True) if onepol is None else (str(onepol).lower() == 'asex', str(onepol).lower() == 'asey')
        if asepol[0]:
            gs.FIELDX = gs.FIELDX + sig * noise[:, :nfc]
        if asepol[1]:
            ny = sig * noise[:, nfc:]
            gs.FIELDY = gs.FIELDY + ny if isy else ny


# --------------------------------------------------------------------------
def inverse_pmd_matrix(brf_list, nfft, mat=None, gvd=True):
    """inverse_pmd.m:73-131: U(omega) of a chain of fibers and its inverse.  ``mat`` is options.mat (a change of the
    reference system, applied first, :87-89 -- update_U keeps its first row only and completes it to the form
    [a b; -b* a*], :158-161); ``gvd`` False is options.gvd = 'no' (:79,124-128)."""
    u = np.zeros((2, 2, nfft), dtype=np.complex128)
    u[0, 0, :] = 1
    u[1, 1, :] = 1
    allgvd = np.zeros(nfft)

    def update(l1, l2, mr, uold):
        t11, t12 = l1 * mr[0, 0], l1 * mr[0, 1]
        un = np.empty_like(uold)
        un[0, 0] = t11 * uold[0, 0] + t12 * uold[1, 0]
        un[0, 1] = t11 * uold[0, 1] + t12 * uold[1, 1]
        un[1, 0] = -np.conj(un[0, 1])
        un[1, 1] = np.conj(un[0, 0])
        return un

    one = np.ones(nfft)
    if mat is not None:
        u = update(one, one, np.asarray(mat, dtype=np.complex128), u)
    for brf in brf_list:
        th, ep, db0 = brf['theta'], brf['epsilon'], brf['db0']
        db1 = brf['db1'][:, 0]
        ntr = len(th)
        mr = _mat_r(th[0], ep[0], np.float64)
        l1 = fastexp(-(0.5 * (db1 + db0[0])))
        u = update(l1, 1 / l1, mr.conj().T, u)
        for k in range(1, ntr):
            m1 = _mat_r(th[k - 1], ep[k - 1], np.float64)
            m2 = _mat_r(th[k], ep[k], np.float64)
            l1 = fastexp(-(0.5 * (db1 + db0[k])))
            u = update(l1, 1 / l1, m2.conj().T @ m1, u)
        u = update(one, one, _mat_r(th[-1], ep[-1], np.float64), u)
        allgvd = allgvd + brf['betat'][:, 0] * brf['lcorr'] * ntr
    if gvd:
        u = u * fastexp(-allgvd)[None, None, :]
    uinv = np.empty_like(u)
    uinv[0, 0], uinv[0, 1] = np.conj(u[0, 0]), np.conj(u[1, 0])
    uinv[1, 0], uinv[1, 1] = np.conj(u[0, 1]), np.conj(u[1, 1])
    return uinv, u


def inverse_pmd(gs: GState, brf_list, options=None):
    """inverse_pmd.m:60-141 (single column) -> (Uinv, U), both [2, 2, Nfft].  The field is transformed when
    options.apply is absent -- or equal to 'n', the one value the test at :135 lets through; any other value,
    the documented 'no' included, leaves the field alone."""
    if gs.FIELDX.shape[1] > 1:
        raise ValueError('inverse_pmd can be used only with a unique field.')
    options = options or {}
    nfft = gs.NSYMB * gs.NT
    uinv, u = inverse_pmd_matrix(brf_list, nfft, mat=options.get('mat'), gvd=options.get('gvd') != 'no')
    if 'apply' not in options or options['apply'] == 'n':
        fx, fy = fft(gs.FIELDX)[:, 0], fft(gs.FIELDY)[:, 0]
        gs.FIELDX = ifft((uinv[0, 0] * fx + uinv[0, 1] * fy)[:, None])
        gs.FIELDY = ifft((uinv[1, 0] * fx + uinv[1, 1] * fy)[:, None])
        gs.DISP = np.zeros((2, gs.NCH))
    return uinv, u


def count_errors(pat_hat, pat) -> int:
    """ber_estimate.m:118: err = sum(sum(pat ~= pat_hat)) (integer, exact)."""
    return int(np.count_nonzero(np.asarray(pat) != np.asarray(pat_hat)))


def rel_l2(ux, uy, rx, ry) -> float:
    """SURVEY 8c error metric: ||[ux;uy]-[rx;ry]||_2 / ||[rx;ry]||_2 over all columns."""
    num = np.sum(np.abs(np.asarray(ux, dtype=np.clongdouble) - rx) ** 2) + \
        np.sum(np.abs(np.asarray(uy, dtype=np.clongdouble) - ry) ** 2)
    den = np.sum(np.abs(np.asarray(rx, dtype=np.clongdouble)) ** 2) + \
        np.sum(np.abs(np.asarray(ry, dtype=np.clongdouble)) ** 2)
    return float(np.sqrt(num / den))
