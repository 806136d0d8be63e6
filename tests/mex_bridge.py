"""`ssfm_mex` for the mini M interpreter (oracle/mini_m), two ways:

  real_gateway()    the compiled gateway itself -- mex/ssfm_mex.c built against the mex.h stand-in as
                    mex/build/libssfm_mex_shim.so -- called through ctypes: the interpreter's values become mxArrays
                    (split real / imaginary planes), mexFunction runs, the plhs[] come back as numpy arrays.  Arrays
                    the gateway returned keep their mxArray alive, so handing them back unchanged hands the same data
                    pointers back, as MATLAB / Octave do for an unmodified variable: that is what the gateway's
                    resident-field check looks at.
  oracle_gateway()  a stand-in with the same argument list whose propagation is the numpy oracle (CPU tests of the
                    front-end matlab/fiber.m: everything but the device loop)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, 'mex', 'build', 'libssfm_mex_shim.so')


class _MxArray(C.Structure):
    _fields_ = [('m', C.c_size_t), ('n', C.c_size_t), ('pr', C.POINTER(C.c_double)), ('pi', C.POINTER(C.c_double)),
                ('str', C.c_char_p)]


class RealGateway:
    def __init__(self):
        lib = C.CDLL(SHIM)
        P = C.POINTER(_MxArray)
        lib.mxCreateDoubleMatrix.restype = P
        lib.mxCreateDoubleMatrix.argtypes = [C.c_size_t, C.c_size_t, C.c_int]
        lib.mxCreateString.restype = P
        lib.mxCreateString.argtypes = [C.c_char_p]
        lib.mxDestroyArray.argtypes = [P]
        lib.mex_shim_call.argtypes = [C.c_int, C.POINTER(P), C.c_int, C.POINTER(P)]
        lib.mex_shim_last_error.restype = C.c_char_p
        self.lib, self.P = lib, P
        self.alive = {}          # id(numpy array handed to the interpreter) -> (array, mxArray*) it is a view of

    def _to_mx(self, v, tmp):
        lib = self.lib
        if isinstance(v, str):
            a = lib.mxCreateString(v.encode())
            tmp.append(a)
            return a
        hit = self.alive.get(id(v))
        if hit is not None and hit[0] is v and np.array_equal(v, hit[2]):
            return hit[1]                                  # an array we returned, still untouched: same pointers
        a = np.asarray(v)
        if a.dtype == bool:
            a = a.astype(np.float64)
        if a.ndim != 2:
            a = a.reshape(1, -1) if a.ndim == 1 else a.reshape(a.shape[0], -1)
        m, n = a.shape
        cplx = np.iscomplexobj(a)
        mx = lib.mxCreateDoubleMatrix(m, n, 1 if cplx else 0)
        tmp.append(mx)
        if a.size:
            flat = np.asfortranarray(a)
            re = np.ascontiguousarray(flat.real.T).ravel()
            C.memmove(mx.contents.pr, re.ctypes.data, re.nbytes)
            if cplx:
                im = np.ascontiguousarray(flat.imag.T).ravel()
                C.memmove(mx.contents.pi, im.ctypes.data, im.nbytes)
        return mx

    def _from_mx(self, mx):
        c = mx.contents
        m, n = int(c.m), int(c.n)
        if m * n == 0:
            self.lib.mxDestroyArray(mx)
            return np.zeros((0, 0))
        re = np.ctypeslib.as_array(c.pr, shape=(n, m)).T
        if c.pi:
            out = np.empty((m, n), dtype=np.complex128)
            out.real = re
            out.imag = np.ctypeslib.as_array(c.pi, shape=(n, m)).T
        else:
            out = np.array(re, dtype=np.float64)
        # (the interpreter works on a copy in numpy's layout; the mxArray stays alive for the pointer-identity check)
        self.alive[id(out)] = (out, mx, out.copy())        # (the copy tells an in-place edit by M code apart)
        if len(self.alive) > 8:
            for k in list(self.alive)[:-4]:
                self.lib.mxDestroyArray(self.alive.pop(k)[1])
        return out

    def __call__(self, it, args, nargout):
        from oracle.mini_m.interp import MError
        tmp = []
        prhs = (self.P * len(args))(*[self._to_mx(a, tmp) for a in args])
        nl = max(nargout, 1)
        plhs = (self.P * 4)()
        rc = self.lib.mex_shim_call(nargout, plhs, len(args), prhs)
        for t in tmp:
            self.lib.mxDestroyArray(t)
        if rc:
            raise MError(self.lib.mex_shim_last_error().decode())
        return [self._from_mx(plhs[k]) for k in range(nl) if plhs[k]]

    def stats(self):
        tmp = []
        prhs = (self.P * 1)(self._to_mx('stats', tmp))
        plhs = (self.P * 4)()
        self.lib.mex_shim_call(1, plhs, 1, prhs)
        out = [float(plhs[0].contents.pr[k]) for k in range(3)]
        self.lib.mxDestroyArray(plhs[0])
        self.lib.mxDestroyArray(tmp[0])
        return dict(uploads=out[0], downloads=out[1], resident_hits=out[2])


def real_gateway():
    return RealGateway()


def oracle_gateway(log=None):
    """ssfm_mex('fiber', ux, uy, betat, db1, P, gam, fls, plates, scal, opt) with the oracle's loops behind it."""
    import math
    import oracle.fiber_oracle as orc

    def call(it, a, nargout):
        cmd = a[0]
        if cmd == 'ampliflat':                             # [ux,uy] = ssfm_mex('ampliflat', ux, uy, gain, sigma, noise, asepol, opt)
            ux, uy, gain, sigma, noise, asepol = a[1:7]
            g = float(np.asarray(gain).ravel()[0])
            sg = np.asarray(sigma, dtype=np.float64).reshape(1, -1)
            ox, oy = np.asarray(ux) * math.sqrt(g), np.asarray(uy) * math.sqrt(g)
            nz = np.asarray(noise)
            nfc = ox.shape[1]
            if np.any(sg) and nz.size == 2 * ox.size:      # options.noise (ampliflat.m:123-129)
                pol = int(np.asarray(asepol).ravel()[0])
                if pol & 1:
                    ox = ox + sg * nz[:, :nfc]
                if pol & 2:
                    oy = oy + sg * nz[:, nfc:]
            elif np.any(sg):
                raise NotImplementedError('the oracle-backed gateway takes injected noise only')
            return [ox, oy][:max(nargout, 1)]
        if cmd == 'invpmd':    # [ux,uy,Uinv4,U4] = ssfm_mex('invpmd', ux, uy, plates, ntr, lcorr, betat, db1, mat, flags)
            ux, uy, plates, ntr, lcorr, betat, db1, mat, flags = a[1:10]
            plates = np.asarray(plates, dtype=np.float64)
            ntr = [int(v) for v in np.asarray(ntr).ravel()]
            lc = [float(v) for v in np.asarray(lcorr).ravel()]
            brfs, row = [], 0
            for k, nt in enumerate(ntr):
                brfs.append({'db0': plates[row:row + nt, 0], 'theta': plates[row:row + nt, 1], 'epsilon': plates[row:row + nt, 2],
                             'betat': np.asarray(betat)[:, k:k + 1], 'db1': np.asarray(db1)[:, k:k + 1], 'lcorr': lc[k]})
                row += nt
            gvd, apply = [bool(v) for v in np.asarray(flags).ravel()]
            n = np.asarray(ux).shape[0]
            uinv, u = orc.inverse_pmd_matrix(brfs, n, mat=(np.asarray(mat) if np.size(mat) else None), gvd=gvd)
            ox, oy = np.asarray(ux), np.asarray(uy)
            if apply:
                fx, fy = np.fft.fft(ox[:, 0]), np.fft.fft(oy[:, 0])
                ox = np.fft.ifft(uinv[0, 0] * fx + uinv[0, 1] * fy)[:, None]
                oy = np.fft.ifft(uinv[1, 0] * fx + uinv[1, 1] * fy)[:, None]
            four = lambda m: np.reshape(m, (4, n), order='F')      # column n = [M11; M21; M12; M22]
            return [ox, oy, four(uinv), four(u)][:max(nargout, 1)]
        if cmd == 'cohmix':    # [Iric,avgeb] = ssfm_mex('cohmix', sigx, sigy, Hopt, Hel, lo, lophase, band, opt)
            sigx, sigy, hopt, hel, lo, lophase, band = a[1:8]
            n = np.size(sigx)
            ecw, det, balanced = [float(v) for v in np.asarray(lo).ravel()]
            ndfn, ndfnl, ndfnr = [int(v) for v in np.asarray(band).ravel()]
            nind = np.mod(np.arange(1, n + 1) - ndfn - 1, n)
            elo = ecw * np.exp(1j * (det * np.arange(1, n + 1) + (np.asarray(lophase).ravel() if np.size(lophase) else 0.0)))
            hopt, hel = np.asarray(hopt).ravel(), np.asarray(hel).ravel()
            cols, avg = [], []
            for sg in (sigx, sigy):
                if not np.size(sg):
                    avg.append(0.0)
                    continue
                sp = np.fft.fft(np.asarray(sg).ravel())[nind]
                avg.append((np.sum(np.abs(sp[:ndfnl]) ** 2) + np.sum(np.abs(sp[n - ndfnr:]) ** 2)) / n ** 2)
                t = np.fft.ifft(sp * hopt)
                e = [1j * t + 1j * elo, t - elo, 1j * t - elo, -t + 1j * elo]
                i = [np.real(v * np.conj(v)) for v in e]
                pair = [i[0] - i[1], i[2] - i[3]] if balanced else [i[0], i[2]]
                cols += [np.real(np.fft.ifft(np.fft.fft(c) * hel)) for c in pair]
            return [np.stack(cols, axis=1), np.array([avg])][:max(nargout, 1)]
        if cmd != 'fiber':
            raise NotImplementedError(cmd)
        ux, uy, betat, db1, P, gam, fls, plates, scal = a[1:10]
        opt = np.asarray(a[10], dtype=np.float64).ravel() if len(a) > 10 else np.zeros(0)
        o = lambda k, d=0.0: float(opt[k]) if opt.size > k else d
        dzmaxt, dphimaxt, alphalin, lf, nplates, manakov = [float(v) for v in np.asarray(P).ravel()]
        fls = [int(v) for v in np.asarray(fls).ravel()]
        nfft, nfc = ux.shape
        gam = np.asarray(gam, dtype=np.float64).ravel()
        if np.size(betat) == 0:                            # scalar dispersion mode: fiber.m:350-362 from the scalars
            s = np.asarray(scal, dtype=np.float64).ravel()
            rate, nsymb, nt, b30, dgdrms = s[:5]
            b1, b2 = s[5:5 + nfc], s[5 + nfc:5 + 2 * nfc]
            fn = np.fft.fftshift(-nt / 2.0 + np.arange(nfft) / nsymb)
            omega = 2 * math.pi * rate * fn
            betat = np.stack([omega * b1[k] + 0.5 * omega ** 2 * b2[k] + omega ** 3 * b30 / 6 for k in range(nfc)], axis=1)
            db1 = np.stack([dgdrms * omega * (1 if fls[1] else 0) for _ in range(nfc)], axis=1)
        if log is not None:
            log.append(dict(scalar_field=o(0), precision=o(1), resident=o(2), tolflag=o(3), fls=fls, nplates=nplates))
        if o(0):                                           # scalar dispatches
            trg = {'err': o(4), 'safety': o(5, 0.9)}
            if int(o(3)) == 2:
                firstdz, ncycle, u = orc.scalar_a_ssfm(ux, betat, dzmaxt, dphimaxt, gam, alphalin, nfft, nfc, lf, trg, fls)
            else:
                firstdz, ncycle, u = orc.scalar_ssfm(ux, betat, dzmaxt, dphimaxt, gam, alphalin, nfft, nfc, lf, fls,
                                                     int(o(3)), trg)
            outs = [u, np.zeros((0, 0))]
        else:
            pl = np.asarray(plates, dtype=np.float64)
            brf = ({'db0': pl[:, 0], 'theta': pl[:, 1], 'epsilon': pl[:, 2]} if pl.size
                   else {'db0': np.zeros(1), 'theta': np.zeros(1), 'epsilon': np.zeros(1)})
            firstdz, ncycle, x1, y1, _, _ = orc.matrix_ssfm(ux, uy, betat, db1, dzmaxt, dphimaxt, gam, alphalin, nfc, lf,
                                                            int(nplates), 'yes' if manakov else 'no', fls, brf)
            outs = [x1, y1]
        outs += [np.array([[float(firstdz)]]), np.array([[float(ncycle)]])]
        return outs[:max(nargout, 1)]
    return call
