"""Wall-clock throughput of one C2 span on a resident batch (no per-pass events, so realization groups may
overlap): python tools/span_time.py [batch] [log2N] [reps] [f64|f32]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import polmux_b200 as pmx
from polmux_b200 import _lib, synth, mc
from polmux_b200.fiber import fiber_setup, setup_to_desc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
LG = int(sys.argv[2]) if len(sys.argv) > 2 else 20
REPS = int(sys.argv[3]) if len(sys.argv) > 3 else 4
PREC = sys.argv[4] if len(sys.argv) > 4 else 'f64'
PC = {'f64': _lib.PMX_F64, 'f32': _lib.PMX_F32}[PREC]
nsymb, nt = 1 << (LG - 4), 16
N = nsymb * nt
ex, ey, _, _ = synth.pdm_qpsk(nsymb, nt, 1)
pmx.reset_all(nsymb, nt, 1)
G = pmx.GSTATE
G.SYMBOLRATE, G.LAMBDA, G.POWER = bench.RATE, np.array([1550.0]), np.array([bench.PAVG])
pmx.create_field('unique', ex, ey, {'power': 'average'})
setup = fiber_setup(bench.fiber_params(8e4, 100), 'gps-', rng=np.random.Generator(np.random.PCG64(0)))
ctx = _lib.Context(0)
d = [mc.draw_plates(1000 + b, 100) for b in range(B)]
pl = [np.stack([x[i] for x in d]) for i in range(3)]
desc, keep = setup_to_desc(setup, batch=B, plate_sets=B, db0=pl[0], theta=pl[1], epsilon=pl[2], precision=PREC)
plan = _lib.Plan(ctx, desc, keep)
tx = _lib.DeviceField(ctx, N, 1, 1, precision=PC); tx.upload(G.FIELDX, G.FIELDY)
work = _lib.DeviceField(ctx, N, 1, B, precision=PC)
work.broadcast_from(tx); res = plan.execute(work); ctx.sync()
best = 1e9
for _ in range(REPS):
    work.broadcast_from(tx); ctx.sync()
    t0 = time.perf_counter()
    res = plan.execute(work); ctx.sync()
    best = min(best, time.perf_counter() - t0)
sa = float(res.ncycle.sum()) * N
print(PREC, 'batch %d N=2^%d groups=%s: %.2f ms per span, %.2f GSa*steps/s, %.1f ps/Sa*step (best of %d, host clock)' % (
    B, LG, os.environ.get('PMX_GROUPS', '1'), best * 1e3, sa / best / 1e9, best / sa * 1e12, REPS), flush=True)
