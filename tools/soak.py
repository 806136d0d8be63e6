"""Stress run over sizes, batches, flags and precisions: every execute must return, with finite fields and a sane
step count (looks for hangs / races in the persistent kernels), and the FP32 result must agree with the FP64 one to
1e-4 -- the two precisions take different code paths through the per-bin physics (FP64: difference recurrence, K-form
boundary matrices, common-scalar extraction; FP32: per-bin phasors, plain matrices), so their agreement over all sizes,
column counts, flags and batches is an independent cross-check of both.
Usage: python tools/soak.py [repeats]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import polmux_b200 as pmx
from polmux_b200 import _lib, synth, mc
from polmux_b200.fiber import fiber_setup, setup_to_desc

REP = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ctx = _lib.Context(0)
t00 = time.time()
nrun = 0
for lg in (6, 8, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22):
    nsymb, nt = 1 << (lg - 3), 8
    N = nsymb * nt
    for nch, ftype in ((1, 'unique'), (3, 'sepfields'), (8, 'sepfields')):
        if (nch == 3 and lg > 18) or (nch == 8 and lg > 12):
            continue
        ex, ey, _, _ = synth.pdm_qpsk(nsymb, nt, nch)
        pmx.reset_all(nsymb, nt, nch)
        G = pmx.GSTATE
        G.SYMBOLRATE, G.LAMBDA, G.POWER = bench.RATE, synth.wdm_lambdas(nch, 1550.0, 0.4), np.full(nch, bench.PAVG)
        pmx.create_field(ftype, ex, ey, {'power': 'average'})
        for flag, man, npl in (('gps-', 'yes', 20), ('gps-', 'no', 7), ('gp--', 'no', 40), ('g-s-', 'no', 1)):
            fib = bench.fiber_params(4e4, npl)
            fib['manakov'] = man
            setup = fiber_setup(fib, flag, rng=np.random.Generator(np.random.PCG64(lg)))
            nfc = setup.nfc
            for batch in ((1, 3, 8) if lg <= 20 else (1, 2)):
                fields = {}
                for prec in ('f64', 'f32'):
                    pc = _lib.PMX_F64 if prec == 'f64' else _lib.PMX_F32
                    d = [mc.draw_plates(7 + b, setup.nplates) for b in range(batch)]
                    pl = [np.stack([x[i] for x in d]) for i in range(3)]
                    desc, keep = setup_to_desc(setup, batch=batch, plate_sets=batch, db0=pl[0], theta=pl[1], epsilon=pl[2],
                                               precision=prec)
                    plan = _lib.Plan(ctx, desc, keep)
                    tx = _lib.DeviceField(ctx, N, nfc, 1, precision=pc)
                    tx.upload(G.FIELDX, G.FIELDY)
                    work = _lib.DeviceField(ctx, N, nfc, batch, precision=pc)
                    for r in range(REP):
                        work.broadcast_from(tx)
                        res = plan.execute(work)
                        nrun += 1
                    x, y = work.download(0, 1)
                    ok = np.isfinite(x).all() and np.isfinite(y).all() and (res.ncycle > 0).all() and (res.ncycle < 500).all()
                    if not ok:
                        print('BAD', lg, nch, flag, man, batch, prec, res.ncycle.tolist(), flush=True)
                        sys.exit(1)
                    fields[prec] = (x, y)
                    plan.close()
                    del work, tx
                a, b = fields['f64'], fields['f32']
                err = np.sqrt((np.abs(a[0] - b[0]) ** 2).sum() + (np.abs(a[1] - b[1]) ** 2).sum()) / np.sqrt(
                    (np.abs(a[0]) ** 2).sum() + (np.abs(a[1]) ** 2).sum())
                worst = max(globals().get('worst', 0.0), err)
                if err > 1e-4:
                    print('FP32/FP64 DISAGREE', lg, nch, flag, man, batch, err, flush=True)
                    sys.exit(1)
    print('N=2^%d ok (%d executes so far, %.0f s; largest FP32-FP64 rel-L2 %.1e)' % (lg, nrun, time.time() - t00, worst), flush=True)
print('soak passed: %d executes' % nrun, flush=True)
