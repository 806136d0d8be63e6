"""Throughput of one span of every BASELINE.json configuration on a resident batch (host clock around
pmx_fiber_exec, best of REPS):  python tools/config_times.py [reps]
C1 Run_my_PDM_QPSK (N=2^16, 100 km CNLSE, 10 plates), C2 ex20 (N=2^20, 80 km Manakov, 100 plates), C3 ex24_pmd
(N=2^20, 200 plates, DGD 0.5: the reference's linear 'gp--' and 'gps-'), C4 nine-channel WDM (N=2^22, 80 km Manakov)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import polmux_b200 as pmx
from polmux_b200 import _lib, synth, mc
from polmux_b200.fiber import fiber_setup, setup_to_desc

REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ctx = _lib.Context(0)
CASES = [  # name, nsymb, nt, nch, pavg per channel, fiber overrides, flag, batch
    ('C1 batch 1', 1 << 12, 16, 1, 2.0, dict(length=1e5, dgd=1.0, nplates=10, manakov='no'), 'gps-', 1),
    ('C1 batch 64', 1 << 12, 16, 1, 2.0, dict(length=1e5, dgd=1.0, nplates=10, manakov='no'), 'gps-', 64),
    ('C2 batch 1', 1 << 16, 16, 1, 2.0, dict(length=8e4, dgd=0.1, nplates=100, manakov='yes'), 'gps-', 1),
    ('C2 batch 8', 1 << 16, 16, 1, 2.0, dict(length=8e4, dgd=0.1, nplates=100, manakov='yes'), 'gps-', 8),
    ('C3 gp-- batch 8', 1 << 16, 16, 1, 2.0, dict(length=8e4, dgd=0.5, nplates=200), 'gp--', 8),
    ('C3 gps- batch 8', 1 << 16, 16, 1, 2.0, dict(length=8e4, dgd=0.5, nplates=200, manakov='yes'), 'gps-', 8),
    ('C4 batch 2', 1 << 16, 64, 9, 1.0, dict(length=8e4, dgd=0.1, nplates=100, manakov='yes'), 'gps-', 2),
]
for name, nsymb, nt, nch, pw, over, flag, B in CASES:
    N = nsymb * nt
    ex, ey, _, _ = synth.pdm_qpsk(nsymb, nt, nch)
    pmx.reset_all(nsymb, nt, nch)
    G = pmx.GSTATE
    G.SYMBOLRATE, G.LAMBDA, G.POWER = 28.0, synth.wdm_lambdas(nch, 1550.0, 0.4), np.full(nch, pw)
    pmx.create_field('unique', ex, ey, {'power': 'average'})
    fib = dict(synth.SMF)
    fib.update(over)
    setup = fiber_setup(fib, flag, rng=np.random.Generator(np.random.PCG64(0)))
    d = [mc.draw_plates(1000 + b, setup.nplates) for b in range(B)]
    pl = [np.stack([x[i] for x in d]) for i in range(3)]
    desc, keep = setup_to_desc(setup, batch=B, plate_sets=B, db0=pl[0], theta=pl[1], epsilon=pl[2])
    plan = _lib.Plan(ctx, desc, keep)
    tx = _lib.DeviceField(ctx, N, 1, 1)
    tx.upload(G.FIELDX, G.FIELDY)
    work = _lib.DeviceField(ctx, N, 1, B)
    work.broadcast_from(tx)
    res = plan.execute(work)
    ctx.sync()
    best = 1e9
    for _ in range(REPS):
        work.broadcast_from(tx)
        ctx.sync()
        t0 = time.perf_counter()
        res = plan.execute(work)
        ctx.sync()
        best = min(best, time.perf_counter() - t0)
    steps = float(res.ncycle.sum())
    sa = steps * N
    trunks = setup.nplates * B
    print('%-16s N=2^%d  ncycle %d  %.2f ms per span  %.2f GSa*steps/s  (%.1f GSa*trunks/s)  %.0f%% of the 192 B/Sa*step HBM roofline'
          % (name, int(np.log2(N)), int(res.ncycle[0]), best * 1e3, sa / best / 1e9, trunks * N / best / 1e9,
             100 * 192 * sa / best / 1e9 / 6554.2), flush=True)
    for f in (tx, work):
        f.close()
    plan.close()
