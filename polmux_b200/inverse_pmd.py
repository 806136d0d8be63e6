"""inverse_pmd(brf) -- inverse_pmd.m:1.  Undoes the GVD and PMD of a chain of fibers on GSTATE.FIELDX/FIELDY.

The reference builds U(omega) = prod_fibers R_last * prod_k [D_k * R_k' R_(k-1)] * R_1' * exp(-i*betat*L) on the
host and multiplies the spectrum by inv(U) (inverse_pmd.m:91-141).  inv(R_n D_n R_n' ... R_1 D_1 R_1') is the same
product taken backwards with every phase negated, i.e. the linear step of a fiber with the plates in reverse
order, db0 -> -db0, db1 -> -db1, betat -> -betat and no loss: one single-step run of the SSFM kernels per fiber
(pass B applies the whole-trunk Jones product), the field staying on the device in between.
options.mat (a change of the reference system, inverse_pmd.m:87-89) multiplies U from the right, so its inverse is a
constant Jones matrix applied after the chain (pmx_field_jones); the [Uinv, U] outputs come from pmx_pmd_matrix, one
thread per frequency in the reference's order of operations."""
from __future__ import annotations

import math

import numpy as np

from . import _lib
from .fiber import FiberSetup, setup_to_desc
from .gstate import GSTATE


def inverse_pmd(brf, options=None, ctx=None, nargout=0):
    """inverse_pmd(brf[, options]); with nargout = 1 / 2 returns Uinv / (Uinv, U), [2, 2, Nfft] each (inverse_pmd.m:21-31).
    options: 'gvd' ('no': PMD only, :79), 'mat' ([2,2] unitary, :87-89), 'apply' (:135 -- the field is transformed when
    the key is absent or equal to 'n'; every other value, the documented 'no' included, leaves it alone)."""
    G = GSTATE
    options = dict(options or {})
    unknown = set(options) - {'gvd', 'mat', 'apply'}
    if unknown:
        raise ValueError('inverse_pmd: unknown options %s' % sorted(unknown))
    keep_gvd = str(options.get('gvd', 'yes')) != 'no'          # options.gvd = 'no': PMD only (inverse_pmd.m:79,124)
    apply = 'apply' not in options or options['apply'] == 'n'  # inverse_pmd.m:135
    mat = None
    if 'mat' in options:
        mat = np.asarray(options['mat'], dtype=np.complex128)
        if mat.shape != (2, 2):
            raise ValueError('options.mat must be a [2,2] matrix')
    brfs = [brf] if isinstance(brf, dict) else list(brf)
    nfr, nfc = G.field_shape()
    if nfc > 1:
        raise ValueError('inverse_pmd can be used only with a unique field.')     # inverse_pmd.m:63
    n = G.NSYMB * G.NT
    ctx = ctx or _lib.default_context()
    out = ()
    if nargout >= 1:
        uinv, u = _lib.pmd_matrix(ctx, n, brfs, mat=mat, gvd=keep_gvd, want_u=nargout >= 2)
        out = (uinv,) if nargout == 1 else (uinv, u)
    if not apply:
        return out[0] if nargout == 1 else (out or None)
    if not G.has_y():
        raise ValueError('inverse_pmd needs both polarizations')
    fld, hx, hy = G.take_device(ctx, _lib.PMX_F64)      # the field fiber() left in HBM, or an upload
    for b in reversed(brfs):
        th = np.asarray(b['theta'], dtype=np.float64).ravel()
        ep = np.asarray(b['epsilon'], dtype=np.float64).ravel()
        db0 = np.asarray(b['db0'], dtype=np.float64).ravel()
        ntr = th.size
        length = float(b['lcorr']) * ntr
        betat = -np.asarray(b['betat'], dtype=np.float64).reshape(n, 1) * (1.0 if keep_gvd else 0.0)
        db1 = -np.asarray(b['db1'], dtype=np.float64).reshape(n, 1)
        inv = FiberSetup(nfft=n, nfc=1, fls=(1, 1, 0, 0), dphimaxt=math.inf, dzmaxt=length, length=length,
                         alphalin=0.0, gam=np.zeros(1), betat=betat, db1=db1, manakov=False, nplates=ntr,
                         brf={'db0': -db0[::-1], 'theta': th[::-1], 'epsilon': ep[::-1]}, isv=True, isy=True,
                         b1=np.zeros(1), dch=np.zeros(1), scalars={})
        desc, keep = setup_to_desc(inv, disp_mode='vector')
        try:
            plan = _lib.Plan(ctx, desc, keep)
        except Exception:
            if b is brfs[-1]:
                G.restore_host(fld, hx, hy)                 # nothing ran yet: the caller's field is as it was
            else:
                fld.close()
            raise
        try:
            plan.execute(fld)
        except Exception:
            fld.close()
            raise
        finally:
            plan.close()
    if mat is not None:
        # U = chain * [m11 m12; -m12* m11*] (update_U keeps the first row of options.mat): Uinv = [m11* -m12; m12* m11] * chain'
        m11, m12 = mat[0, 0], mat[0, 1]
        try:
            _lib.field_jones(ctx, fld, [[np.conj(m11), -m12], [np.conj(m12), m11]])
        except Exception:
            fld.close()
            raise
    G.put_device(fld, hx, hy)
    G.DISP = np.zeros((2, G.NCH))                         # inverse_pmd.m:141
    return out[0] if nargout == 1 else (out or None)
