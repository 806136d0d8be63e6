"""Do two half-batches on two streams (half-size persistent grids) beat one full batch?  The passes are bound by
different resources (A/C: TMA/HBM, B: FP64), so CTAs of different passes sharing an SM could complement each other.
Usage: PMX_GRID_DIV=2 python tools/two_stream.py 2    |   python tools/two_stream.py 1"""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import polmux_b200 as pmx
from polmux_b200 import _lib, synth, mc
from polmux_b200.fiber import fiber_setup, setup_to_desc

NS = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B, LG = 8, 20
nsymb, nt = 1 << (LG - 4), 16
N = nsymb * nt
ex, ey, _, _ = synth.pdm_qpsk(nsymb, nt, 1)
pmx.reset_all(nsymb, nt, 1)
G = pmx.GSTATE
G.SYMBOLRATE, G.LAMBDA, G.POWER = bench.RATE, np.array([1550.0]), np.array([bench.PAVG])
pmx.create_field('unique', ex, ey, {'power': 'average'})
setup = fiber_setup(bench.fiber_params(8e4, 100), 'gps-', rng=np.random.Generator(np.random.PCG64(0)))
bb = B // NS
jobs = []
for s in range(NS):
    ctx = _lib.Context(0)
    d = [mc.draw_plates(1000 + s * bb + b, 100) for b in range(bb)]
    pl = [np.stack([x[i] for x in d]) for i in range(3)]
    desc, keep = setup_to_desc(setup, batch=bb, plate_sets=bb, db0=pl[0], theta=pl[1], epsilon=pl[2])
    plan = _lib.Plan(ctx, desc, keep)
    tx = _lib.DeviceField(ctx, N, 1, 1); tx.upload(G.FIELDX, G.FIELDY)
    work = _lib.DeviceField(ctx, N, 1, bb)
    jobs.append((ctx, plan, tx, work))
res = [None] * NS
def run(i, reps):
    ctx, plan, tx, work = jobs[i]
    for _ in range(reps):
        work.broadcast_from(tx)
        res[i] = plan.execute(work)
    ctx.sync()
for reps in (1, 3):
    th = [threading.Thread(target=run, args=(i, reps)) for i in range(NS)]
    t0 = time.perf_counter()
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
sa = sum(float(r.ncycle.sum()) for r in res) * N * 3
print('%d stream(s) x batch %d: %.1f ms per pass over the span, %.2f GSa*steps/s (host wall clock)' % (NS, bb, dt / 3 * 1e3, sa / dt / 1e9))
