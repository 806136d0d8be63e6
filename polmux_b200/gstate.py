"""Global simulation state, mirroring the reference's GSTATE / CONSTANTS structs.

reset_all.m:105-112 (CONSTANTS) and :152-174 (GSTATE fields).  Like the
reference, the state is a process-wide object that the Tx functions fill and
the in-line devices (fiber, ampliflat) mutate in place, so scripts written
against the reference read the same here:

    reset_all(Nsymb, Nt, Nch); ...; create_field('unique', Ex, Ey); fiber(fib, 'gps-')
"""
from __future__ import annotations

import numpy as np


class _Struct:
    def __repr__(self):
        return '%s(%s)' % (type(self).__name__, ', '.join(sorted(self.__dict__)))


class _Constants(_Struct):
    CLIGHT = 299792458.0          # speed of light in vacuum [m/s]      reset_all.m:106
    HPLANCK = 6.62606896e-34      # Planck's constant [J*s]              reset_all.m:107
    ECHARGE = 1.602176487e-19     # electron's charge [C]                reset_all.m:111
    KBOLTZMANN = 1.3806504e-23    # Boltzmann's constant [J/K]           reset_all.m:112


CONSTANTS = _Constants()
GSTATE = _Struct()

# stream standing in for the interpreter's global rand/randn state
_rng = np.random.default_rng(0)


def seed(value: int):
    """Equivalent of rand('state',k) / randn('state',k): reseed the global stream."""
    global _rng
    _rng = np.random.Generator(np.random.PCG64(int(value)))


def rng() -> np.random.Generator:
    return _rng


def reset_all(Nsymb: int, Nt: int, Nch: int, *opts):
    """reset_all(Nsymb,Nt,Nch[,outdir[,'noprint']]) -- reset_all.m:114-174.

    Printing to simul_out is not built (GSTATE.PRINT is always False)."""
    if len(opts) > 2:
        raise ValueError('Invalid number of inputs')
    for k in list(GSTATE.__dict__):
        delattr(GSTATE, k)
    GSTATE.PRINT = False
    if opts:
        if not isinstance(opts[0], str):
            raise ValueError('directory name must be a string')
        GSTATE.DIR = opts[1] if (len(opts) == 2 and opts[0] == 'noprint') else opts[0]
    stepf = 1.0 / Nsymb
    n = int(Nsymb) * int(Nt)
    # fftshift(-Nt/2 : 1/Nsymb : Nt/2-1/Nsymb)                         reset_all.m:153
    GSTATE.FN = np.fft.fftshift(-Nt / 2.0 + np.arange(n) * stepf)
    GSTATE.NSYMB = int(Nsymb)
    GSTATE.NT = int(Nt)
    GSTATE.NCH = int(Nch)
    GSTATE.SYMBOLRATE = None
    GSTATE.FIELDX = None
    GSTATE.FIELDY = None
    GSTATE.FIELDX_TX = None
    GSTATE.FIELDY_TX = None
    GSTATE.DELAY = None
    GSTATE.DISP = None
    GSTATE.LAMBDA = None
    GSTATE.POWER = None
    return GSTATE
