"""Small fields: one resident fiber (100 km, CNLSE, 10 plates), on-chip kernel against the three passes (2^12 only).
python tools/small_times.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import polmux_b200 as pmx
from polmux_b200 import _lib, synth, mc
from polmux_b200.fiber import fiber_setup, setup_to_desc

ctx = _lib.Context(0)
for lg, B, env in ((8, 1, '0'), (10, 1, '0'), (10, 148, '0'), (12, 1, '0'), (12, 1, '1'), (12, 148, '0'), (12, 148, '1'), (12, 1184, '0'), (12, 1184, '1')):
    os.environ['PMX_NO_ONCHIP'] = env
    nsymb, nt = (1 << lg) // 16, 16
    N = nsymb * nt
    ex, ey, _, _ = synth.pdm_qpsk(nsymb, nt, 1)
    pmx.reset_all(nsymb, nt, 1)
    G = pmx.GSTATE
    G.SYMBOLRATE, G.LAMBDA, G.POWER = 28.0, np.array([1550.0]), np.array([2.0])
    pmx.create_field('unique', ex, ey, {'power': 'average'})
    fib = dict(synth.SMF); fib.update(length=1e5, dgd=1.0, nplates=10, manakov='no')
    setup = fiber_setup(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(0)))
    d = [mc.draw_plates(1000 + b, 10) for b in range(B)]
    pl = [np.stack([x[i] for x in d]) for i in range(3)]
    desc, keep = setup_to_desc(setup, batch=B, plate_sets=B, db0=pl[0], theta=pl[1], epsilon=pl[2])
    plan = _lib.Plan(ctx, desc, keep)
    tx = _lib.DeviceField(ctx, N, 1, 1); tx.upload(G.FIELDX, G.FIELDY)
    work = _lib.DeviceField(ctx, N, 1, B)
    best = 1e9
    for rep in range(4):
        work.broadcast_from(tx); ctx.sync()
        t0 = time.perf_counter(); res = plan.execute(work); ctx.sync()
        if rep: best = min(best, time.perf_counter() - t0)
    sa = float(res.ncycle.sum()) * N
    print('N=2^%-2d batch %-4d %-10s ncycle %d: %.3f ms per fiber call, %.2f us per step, %.2f GSa*steps/s' % (
        lg, B, 'three-pass' if env == '1' else 'on-chip', int(res.ncycle[0]), best * 1e3, best * 1e6 / int(res.ncycle[0]), sa / best / 1e9), flush=True)
    plan.close(); tx.close(); work.close()
