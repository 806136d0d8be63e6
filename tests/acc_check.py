"""Accuracy of the CUDA path vs the oracle on a few cases (prints rel-L2 and step-length deviation)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this script lives in tests/: it uses the oracle as checker)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import oracle.fiber_oracle as orc
import polmux_b200 as pmx
from common import base_fiber, make_tx, rel_l2
for lg, man in ((12, 'no'), (14, 'no'), (16, 'no'), (16, 'yes'), (18, 'yes')):
    fib = base_fiber(length=1e5, dgd=1.0, nplates=10, manakov=man)
    gs = make_tx(1 << (lg - 4), 16)
    orc.fiber(gs, fib, 'gps-', rng=np.random.Generator(np.random.PCG64(1000)))
    pmx.fiber(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(1000)), trace=True)
    G = pmx.GSTATE
    dz_o = np.array([s['dz'] for s in gs.log['schedule']])
    dz_g = pmx.FIBER_LAST['trace_dz']
    print('N=2^%d manakov=%s  rel_l2=%.2e  ncycle %d/%d  max|ddz/dz|=%.2e' % (
        lg, man, rel_l2(G.FIELDX, G.FIELDY, gs.FIELDX, gs.FIELDY), pmx.FIBER_LAST['ncycle'], gs.log['ncycle'],
        np.max(np.abs(dz_g - dz_o) / dz_o) if len(dz_g) == len(dz_o) else -1))
# extended-precision arbiter (the same restatement in np.longdouble): how far are the float64 oracle and the CUDA path from it
for lg, man in ((12, 'no'), (14, 'yes')):
    fib = base_fiber(length=1e5, dgd=1.0, nplates=10, manakov=man)
    gl = make_tx(1 << (lg - 4), 16, real=np.longdouble)
    orc.fiber(gl, fib, 'gps-', rng=np.random.Generator(np.random.PCG64(1000)))
    gs = make_tx(1 << (lg - 4), 16)
    orc.fiber(gs, fib, 'gps-', rng=np.random.Generator(np.random.PCG64(1000)))
    pmx.fiber(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(1000)))
    G = pmx.GSTATE
    print('N=2^%d manakov=%s  vs longdouble arbiter: oracle(f64) %.2e   CUDA %.2e   (CUDA vs oracle %.2e)' % (
        lg, man, rel_l2(gs.FIELDX, gs.FIELDY, gl.FIELDX, gl.FIELDY), rel_l2(G.FIELDX, G.FIELDY, gl.FIELDX, gl.FIELDY),
        rel_l2(G.FIELDX, G.FIELDY, gs.FIELDX, gs.FIELDY)))
# FP32 mode on C1 (2^16 samples, CNLSE, 10 plates, DGD 1 symbol, 100 km) and on a C2-like Manakov span at 2^16
for man, dgd, npl, L in (('no', 1.0, 10, 1e5), ('yes', 0.1, 100, 8e4), ('no', 0.1, 10, 1e5), ('no', 1.0, 10, 2.5e4)):
    fib = base_fiber(length=L, dgd=dgd, nplates=npl, manakov=man)
    out = {}
    for prec in ('f64', 'f32'):
        make_tx(1 << 12, 16)
        pmx.fiber(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(1000)), precision=prec)
        out[prec] = (np.array(pmx.GSTATE.FIELDX), np.array(pmx.GSTATE.FIELDY), pmx.FIBER_LAST['ncycle'])
    print('FP32 vs FP64  N=2^16 manakov=%s dgd=%.1f nplates=%d L=%.0f km: rel_l2=%.2e  ncycle %d/%d' % (
        man, dgd, npl, L * 1e-3, rel_l2(out['f32'][0], out['f32'][1], out['f64'][0], out['f64'][1]), out['f32'][2], out['f64'][2]))
