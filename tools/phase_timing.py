"""PMX_TIMING builds only: where pass-B CTAs spend their cycles (thread 0 of every CTA)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import polmux_b200 as pmx
from polmux_b200 import _lib, synth, mc
from polmux_b200.fiber import fiber_setup, setup_to_desc
B, LG = 8, 20
nsymb, nt = 1 << (LG - 4), 16
N = nsymb * nt
ex, ey, _, _ = synth.pdm_qpsk(nsymb, nt, 1)
pmx.reset_all(nsymb, nt, 1)
G = pmx.GSTATE
G.SYMBOLRATE, G.LAMBDA, G.POWER = bench.RATE, np.array([1550.0]), np.array([bench.PAVG])
pmx.create_field('unique', ex, ey, {'power': 'average'})
setup = fiber_setup(bench.fiber_params(8e4, 100), 'gps-', rng=np.random.Generator(np.random.PCG64(0)))
d = [mc.draw_plates(1000 + b, 100) for b in range(B)]
pl = [np.stack([x[i] for x in d]) for i in range(3)]
ctx = _lib.Context(0)
lib = ctx.lib
lib.pmx_debug_timing.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
buf = np.zeros(24, dtype=np.int64)
lib.pmx_debug_timing(ctx.h, None, 1)       # allocate + zero
desc, keep = setup_to_desc(setup, batch=B, plate_sets=B, db0=pl[0], theta=pl[1], epsilon=pl[2])
plan = _lib.Plan(ctx, desc, keep)
tx = _lib.DeviceField(ctx, N, 1, 1); tx.upload(G.FIELDX, G.FIELDY)
work = _lib.DeviceField(ctx, N, 1, B)
work.broadcast_from(tx); plan.execute(work)
lib.pmx_debug_timing(ctx.h, None, 1)
work.broadcast_from(tx); res = plan.execute(work)
lib.pmx_debug_timing(ctx.h, buf.ctypes.data_as(ctypes.c_void_p), 0)
namesB = ['tile top (plates, ctl, tables)', 'mbarrier wait (TMA)', 'tile LDS + barrier + issue', 'forward FFT', 'Jones + phases', 'inverse FFT', 'twiddle + store + barrier', 'pkg wait + data-independent phasors (before the tile wait)']
namesAC = ['tile top (ctl, tables)', 'mbarrier wait (TMA)', 'tile LDS + barrier + issue', 'NL step (A) / -', 'FFT', 'twiddle/scale + staging STS', 'barrier + TMA store issue (+ max publish)', '-']
for kind, title, names in ((0, 'pass A', namesAC), (1, 'pass B', namesB), (2, 'pass C', namesAC)):
    t = buf[8 * kind:8 * kind + 8].astype(float)
    print('%s phases, share of CTA cycles (thread 0), total %.0f Mcycles:' % (title, t.sum() / 1e6))
    for n_, v in zip(names, t):
        print('  %-44s %5.1f%%' % (n_, 100 * v / max(t.sum(), 1)))
