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


# True: FIELDX/FIELDY stay in HBM between the in-line devices (fiber, ampliflat, inverse_pmd) of a two-polarization
# link and reach the host when somebody reads them; False: every call copies the field up and down, as a MEX call does
RESIDENT = True


class _Resident:
    """A two-polarization field living in HBM after an in-line device ran, plus the host arrays it came from
    (results land in those again when they can, as the reference overwrites GSTATE.FIELDX/FIELDY)."""

    def __init__(self, field, hostx, hosty):
        self.field, self.hostx, self.hosty = field, hostx, hosty


class _GState(_Struct):
    """GSTATE (reset_all.m:152-174).  FIELDX / FIELDY read and assign like the reference's arrays.  Behind them the
    field of a two-polarization link may be resident on the device: fiber() -> ampliflat() -> fiber() ... chains of
    the reference's span loops (ex06_ber.m:110-115, ex20_coherent_polmux.m) then cross PCIe once in and once out
    instead of twice per call.  Reading either attribute downloads the field and gives the device copy up (the host
    array is handed out and may be written in place); assigning does the same before it replaces the array."""

    def __repr__(self):
        return 'GSTATE(%s)' % ', '.join(sorted(k for k in self.__dict__ if not k.startswith('_')) + ['FIELDX', 'FIELDY'])

    # -- host view
    def _materialize(self):
        res = self.__dict__.pop('_res', None)
        if res is None:
            return
        fld = res.field
        n, nfc = fld.nfft, fld.nfc
        bufs = (res.hostx, res.hosty)
        # Value semantics like the interpreter's: an array the caller may still hold (x0 = GSTATE.FIELDX before the
        # fiber) is never overwritten -- unless it is PINNED memory, which a caller allocates precisely to have the
        # transfers land in it (bench e2e; a staging copy would defeat it).
        from . import _lib
        if nfc == 1 and all(isinstance(a, np.ndarray) and a.dtype == np.complex128 and a.shape == (n, 1)
                            and a.flags['C_CONTIGUOUS'] and a.flags['WRITEABLE'] and _lib.host_is_pinned(a)
                            for a in bufs):
            fld.download_into(res.hostx, res.hosty)              # [N,1] is also [1][1][N]; pinned stays pinned
            hx, hy = res.hostx, res.hosty
        else:
            ox, oy = fld.download()
            hx, hy = np.ascontiguousarray(ox[0].T), np.ascontiguousarray(oy[0].T)
        fld.close()
        self.__dict__['_hx'], self.__dict__['_hy'] = hx, hy

    @property
    def FIELDX(self):
        self._materialize()
        return self.__dict__.get('_hx')

    @FIELDX.setter
    def FIELDX(self, value):
        self._materialize()
        self.__dict__['_hx'] = value

    @property
    def FIELDY(self):
        self._materialize()
        return self.__dict__.get('_hy')

    @FIELDY.setter
    def FIELDY(self, value):
        self._materialize()
        self.__dict__['_hy'] = value

    # -- what the in-line devices use (no transfer just to look at the shape)
    def field_shape(self):
        res = self.__dict__.get('_res')
        if res is not None:
            return res.field.nfft, res.field.nfc
        return np.shape(self.__dict__.get('_hx'))

    def has_y(self):
        """~isempty(GSTATE.FIELDY) (fiber.m:253)"""
        if self.__dict__.get('_res') is not None:
            return True
        hy = self.__dict__.get('_hy')
        return hy is not None and np.size(hy) != 0

    def is_resident(self):
        return self.__dict__.get('_res') is not None

    def take_device(self, ctx, precision=None):
        """-> (DeviceField [1][nfc][N] holding FIELDX/FIELDY, hostx, hosty): the resident copy when there is one on
        this context in this precision (None: whatever is resident, FP64 for an upload), else an upload of the host
        arrays.  The caller owns the field until it hands it back with put_device()."""
        from . import _lib
        res = self.__dict__.get('_res')
        if res is not None and (res.field.ctx is not ctx or (precision is not None and res.field.precision != precision)):
            self._materialize()
            res = None
        if res is not None:
            del self.__dict__['_res']
            self.__dict__['_taken_resident'] = True
            # the host arrays of a resident field are stale: only pinned ones are kept, as download targets
            return res.field, res.hostx, res.hosty
        hx, hy = self.__dict__.get('_hx'), self.__dict__.get('_hy')
        n, nfc = np.shape(hx)
        self.__dict__['_taken_resident'] = False
        fld = _lib.DeviceField(ctx, n, nfc, 1, precision=_lib.PMX_F64 if precision is None else precision)
        try:
            fld.upload(hx, hy)
        except Exception:
            fld.close()
            raise
        return fld, hx, hy

    def put_device(self, fld, hostx, hosty):
        """The field an in-line device leaves behind.  With RESIDENT it stays in HBM until it is read."""
        self.__dict__.pop('_taken_resident', None)
        self.__dict__['_res'] = _Resident(fld, hostx, hosty)
        self.__dict__['_hx'] = self.__dict__['_hy'] = None
        if not RESIDENT:
            self._materialize()

    def restore_host(self, fld, hostx, hosty):
        """An in-line device failed before it touched the field: give the caller's state back as it was (the device
        copy stays resident when it was the only copy)."""
        if self.__dict__.pop('_taken_resident', False):
            self.__dict__['_res'] = _Resident(fld, hostx, hosty)
            return
        fld.close()
        self.__dict__['_hx'], self.__dict__['_hy'] = hostx, hosty

    def drop_device(self):
        res = self.__dict__.pop('_res', None)
        if res is not None:
            res.field.close()


CONSTANTS = _Constants()
GSTATE = _GState()

# stream standing in for the interpreter's global rand/randn state
_rng = np.random.default_rng(0)


def seed(value: int):
    """Equivalent of rand('state',k) / randn('state',k): reseed the global stream."""
    global _rng
    _rng = np.random.Generator(np.random.PCG64(int(value)))


def next_ase_seed() -> int:
    """Seed of the next ASE draw when the caller names none: taken from the global stream, so that -- like randn in
    ampliflat.m:132-135 -- every ampliflat() call adds fresh, independent noise and seed(k) makes a run repeatable."""
    return int(_rng.integers(0, 1 << 63, dtype=np.int64))


def rng() -> np.random.Generator:
    return _rng


def reset_all(Nsymb: int, Nt: int, Nch: int, *opts):
    """reset_all(Nsymb,Nt,Nch[,outdir[,'noprint']]) -- reset_all.m:114-174.

    With an output directory (and without 'noprint') GSTATE.PRINT is set and the log GSTATE.DIR/simul_out is opened
    (reset_all.m:176-225); fiber() appends its summary block to it (polmux_b200/simul_out.py)."""
    if len(opts) > 2:
        raise ValueError('Invalid number of inputs')
    GSTATE.drop_device()
    for k in list(GSTATE.__dict__):
        del GSTATE.__dict__[k]
    GSTATE.PRINT = False
    if len(opts) == 1:                                                       # reset_all.m:125-133
        if not isinstance(opts[0], str):
            raise ValueError('directory name must be a string')
        if opts[0] == 'noprint':
            raise ValueError("The output directory cannot be called 'noprint'")
        GSTATE.DIR = opts[0]
        GSTATE.PRINT = True
    elif len(opts) == 2:                                                     # :134-150
        if opts[0] == 'noprint':
            if not isinstance(opts[1], str):
                raise ValueError('directory name must be a string')
            if opts[1] == 'noprint':
                raise ValueError("The output directory cannot be called 'noprint'")
            GSTATE.DIR = opts[1]
        else:
            if not isinstance(opts[0], str):
                raise ValueError('directory name must be a string')
            GSTATE.DIR = opts[0]
            if opts[1] != 'noprint':
                raise ValueError("Use 'noprint' to avoid printing to file")
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
    if GSTATE.PRINT:
        from . import simul_out
        if simul_out.open_log(int(Nsymb), int(Nt), int(Nch)):
            import warnings
            warnings.warn('The output file simul_out is very big')           # reset_all.m:220-222
    return GSTATE
