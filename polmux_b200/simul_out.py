"""The text log of a simulation: GSTATE.DIR/simul_out, written when GSTATE.PRINT is set.

reset_all(Nsymb,Nt,Nch,outdir) opens the log (reset_all.m:176-225), every fiber() appends its summary block
(fiber.m:392-456).  The formats are the reference's fprintf strings; tests/golden/simul_out_*.json hold the text the
reference's own source produces for three cases (captured from the mini interpreter's fprintf)."""
from __future__ import annotations

import datetime
import math
import os
import re
import socket

import numpy as np

from .gstate import CONSTANTS, GSTATE

MAXBYTES = 50e6   # reset_all.m:118: warning size of simul_out


def _f(fmt: str, *vals) -> str:
    """C-style formatting with the interpreter's spelling of non-finite numbers"""
    out = fmt % vals
    if any(isinstance(v, float) and not math.isfinite(v) for v in vals):
        out = re.sub(r'\binf\b', 'Inf', out)
        out = re.sub(r'\bnan\b', 'NaN', out)
    return out


def _append(text: str):
    os.makedirs(GSTATE.DIR, exist_ok=True)
    with open(os.path.join(GSTATE.DIR, 'simul_out'), 'a') as f:
        f.write(text)


def reset_all_block(nsymb: int, nt: int, nch: int) -> str:
    """reset_all.m:193-214"""
    now = datetime.datetime.now()
    t = '++++++++++++++++++++++++++++++++++++++++\n'
    t += '++++       START OF SIMULATION      ++++\n'
    t += '++++                                ++++\n'
    t += '++++ Hostname: %s\n' % socket.gethostname()
    t += '++++ Date: %s %.2d:%.2d:%.2d\t\n' % (now.strftime('%d-%b-%Y'), now.hour, now.minute, now.second)
    t += '++++++++++++++++++++++++++++++++++++++++\n\n\n'
    t += '========================================\n'
    t += '===             reset_all            ===\n'
    t += '========================================\n\n'
    t += 'Global variable GSTATE initialized\n\n'
    t += 'Nsymb = %6d\t (number of symbols)\n' % nsymb
    t += 'Nt   = %6d\t (points x symbol)\n' % nt
    t += 'Nch  = %6d\t (number of channels)\n' % nch
    t += 'Output directory = %s\n' % GSTATE.DIR
    t += '\n****************************************\n\n'
    return t


def open_log(nsymb: int, nt: int, nch: int) -> bool:
    """reset_all.m:176-225: output directories, header block; -> True when simul_out has grown past MAXBYTES"""
    d = GSTATE.DIR
    base = os.path.basename(os.path.normpath(d))
    for sub in ('', base + '.MOD', base + '.ANG'):                           # :178-186
        os.makedirs(os.path.join(d, sub), exist_ok=True)
    _append(reset_all_block(nsymb, nt, nch))
    return os.path.getsize(os.path.join(d, 'simul_out')) > MAXBYTES          # :218-225


def fiber_block(x, flag: str, s, firstdz: float, ncycle: int) -> str:
    """fiber.m:392-456.  s: the FiberSetup of the call; GSTATE.DELAY / DISP already updated (:367-369)."""
    from .fiber import DEF_PLATES, _get, _has
    G = GSTATE
    nch = G.NCH
    length = float(_get(x, 'length'))
    leff = length if s.alphalin == 0 else (1 - math.exp(-s.alphalin * length)) / s.alphalin      # :303-307
    gam = np.asarray(s.gam, dtype=np.float64).reshape(-1)
    gamprint = gam * np.ones(nch) if s.nfc == 1 else gam                                         # :394-398
    lam = float(_get(x, 'lambda'))
    dch = np.asarray(s.dch, dtype=np.float64).reshape(-1)
    with np.errstate(divide='ignore'):
        ld = np.where(dch != 0, 1.0 / (G.SYMBOLRATE ** 2 * np.abs(lam ** 2 / 2 / math.pi / CONSTANTS.CLIGHT * dch * 1e-6)),
                      np.inf)                                                                     # :339-341
        lnl = 1.0 / (gam * np.asarray(G.POWER, dtype=np.float64).reshape(-1))                    # :364
    b30 = float(s.scalars.get('b30', 0.0))
    lds = 1.0 / (G.SYMBOLRATE ** 3 * abs(b30)) if b30 != 0 else math.inf                          # :343-347
    loc_delay = length * G.SYMBOLRATE * np.asarray(s.b1, dtype=np.float64).reshape(-1)           # :366
    t = '========================================\n'
    t += '===              fiber               ===\n'
    t += '========================================\n\n'
    t += 'Fiber parameters:\n\n'
    t += _f('Length:%17.3f  [km]\n', length * 1e-3)
    t += _f('Attenuation:%12.2f  [dB/km] (Leff = %7.3f [km])\n', float(_get(x, 'alphadB')), leff * 1e-3)
    t += _f('lambda of Dc:%11.2f  [nm]\n', lam)
    t += _f('Dc:%21.4f  [ps/nm/km]\n', float(_get(x, 'disp')))
    t += _f('Slope:%18.4f  [ps/nm^2/km]\n', float(_get(x, 'slope')))
    t += _f('n2:%21.2e  [m^2/W]\n', float(_get(x, 'n2')))
    t += _f('Aeff:%19.2f  [um^2]\n\n', float(_get(x, 'aeff')))
    if s.fls[1]:
        ispmf = all(_has(x, k) for k in ('db0', 'theta', 'epsilon'))
        t += _f('DGD:%12.4f  [bits]\n', float(_get(x, 'dgd')))
        t += '# plates:%d  \n' % (s.nplates if ispmf else int(_get(x, 'nplates', DEF_PLATES)))
        t += 'Manakov Equation: %s\n' % str(_get(x, 'manakov', 'no'))
        if ispmf:
            t += _f('db0 = %8.2f, theta = %3.2f*pi, epsilon = %3.2f*pi\n', float(s.brf['db0'][0]),
                    float(s.brf['theta'][0]) / math.pi, float(s.brf['epsilon'][0]) / math.pi)
        else:
            t += 'Random birefringence\n'
    t += "Propagation type: '%s'\n\n" % flag
    if s.tolflag:
        t += _f('Local error x step: %.1e\n', float(_get(x, 'ltol')))
    t += _f('Max NL phase rotation x step: %-6.2g  [rad]\n', float(s.dphimaxt))
    t += _f('Max step: %.2e  [m]\n', float(s.dzmaxt))
    t += _f('Initial step: %.2e  (num. steps: %d)\n\n', float(firstdz), int(ncycle))
    t += 'Channel properties (Ld: disp. length. Lnl: NL length):\n\n'
    for k in range(nch):
        t += _f('ch. #%.2d: Dc = %.4f  [ps/nm/km]   (Ld = %3.2e [km])\n', k + 1, float(dch[k]), float(ld[k]) * 1e-3)
        t += _f('\t gamma = %.3e [1/mW/km] (Lnl = %3.2e [km])\n', float(gamprint[k]) * 1e3, float(lnl[k]) * 1e-3)
        t += _f('\t sqrt(Ld/Lnl) = %.4f\n', math.sqrt(float(ld[k]) / float(lnl[k])))
        t += _f('\t local delay = %.3f\n', float(loc_delay[k]))
    t += _f('\nSlope length Lds: %-3.2e  [km]\n', lds * 1e-3)
    t += '\nGlobal  delay (ch.1 -> %d)\n' % nch
    # GSTATE.DELAY(kch) is a LINEAR index into the [npol x NCH] matrix (:445-447): column-major, first NCH entries
    dl = np.asarray(G.DELAY, dtype=np.float64).ravel(order='F')
    t += ''.join(_f('%.3f  ', float(dl[k])) for k in range(nch))
    t += '\nGlobal cumulated dispersion [ps/nm] '
    t += '(ch.1 -> %d)\n' % nch
    dp = np.asarray(G.DISP, dtype=np.float64).ravel(order='F')
    t += ''.join(_f('%.3f  ', float(dp[k])) for k in range(nch))
    t += '\n****************************************\n\n'
    return t


def log_fiber(x, flag: str, s, firstdz: float, ncycle: int):
    if getattr(GSTATE, 'PRINT', False):
        _append(fiber_block(x, flag, s, firstdz, ncycle))
