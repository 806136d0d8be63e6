"""Per-pass device time for different flag sets (what each part of a pass costs).
Usage: python tools/pass_breakdown.py [batch] [log2N] [number of flag sets]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import polmux_b200 as pmx
from polmux_b200 import _lib, synth, mc
from polmux_b200.fiber import fiber_setup, setup_to_desc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
LG = int(sys.argv[2]) if len(sys.argv) > 2 else 20
NF = int(sys.argv[3]) if len(sys.argv) > 3 else 6
PREC = sys.argv[4] if len(sys.argv) > 4 else 'f64'
PC = {'f64': _lib.PMX_F64, 'f32': _lib.PMX_F32}[PREC]
SAB = 32 if PREC == 'f64' else 16
nsymb, nt = 1 << (LG - 4), 16
N = nsymb * nt
ex, ey, _, _ = synth.pdm_qpsk(nsymb, nt, 1)
ctx = _lib.Context(0)
for flag, man in (('gps-', 'yes'), ('gps-', 'no'), ('g-s-', 'no'), ('--s-', 'no'), ('g---', 'no'), ('gp--', 'no'))[:NF]:
    pmx.reset_all(nsymb, nt, 1)
    G = pmx.GSTATE
    G.SYMBOLRATE, G.LAMBDA, G.POWER = bench.RATE, np.array([1550.0]), np.array([bench.PAVG])
    pmx.create_field('unique', ex, ey, {'power': 'average'})
    fib = bench.fiber_params(bench.SPAN_KM * 1e3, bench.NPLATES)
    fib['manakov'] = man
    if flag in ('--s-',):
        G.FIELDX = np.tile(G.FIELDX, (1, 2)); G.FIELDY = np.tile(G.FIELDY, (1, 2)); G.NCH = 2
        G.LAMBDA, G.POWER = np.array([1549.8, 1550.2]), np.array([bench.PAVG] * 2)
        G.DELAY, G.DISP = np.zeros((2, 2)), np.zeros((2, 2))
    setup = fiber_setup(fib, flag, rng=np.random.Generator(np.random.PCG64(0)))
    nfc = setup.nfc
    bb = max(1, B // nfc)
    d = [mc.draw_plates(1000 + b, setup.nplates) for b in range(bb)]
    pl = [np.stack([x[i] for x in d]) for i in range(3)]
    desc, keep = setup_to_desc(setup, batch=bb, plate_sets=bb, db0=pl[0], theta=pl[1], epsilon=pl[2], precision=PREC)
    plan = _lib.Plan(ctx, desc, keep)
    tx = _lib.DeviceField(ctx, N, nfc, 1, precision=PC)
    tx.upload(G.FIELDX, G.FIELDY)
    work = _lib.DeviceField(ctx, N, nfc, bb, precision=PC)
    for rep in range(2):
        work.broadcast_from(tx)
        ctx.profile(rep == 1)
        res = plan.execute(work)
    ms, n = ctx.profile_read()
    ctx.profile(False)
    sa = float(res.ncycle.sum()) * N * nfc
    print('%s manakov=%-3s nfc=%d ncycle=%3d  GB/s: A %6.0f  B %6.0f  C %6.0f   ps/Sa: A %.1f B %.1f C %.1f' % (
        flag, man, nfc, int(res.ncycle[0]), *[2 * SAB * sa / (ms[i] * 1e-3) / 1e9 for i in range(3)],
        *[ms[i] * 1e-3 / sa * 1e12 for i in range(3)]))
