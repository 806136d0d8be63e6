import cProfile, pstats, sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
import polmux_b200 as pmx
from polmux_b200 import _lib, synth
NS, NT = bench.NSYMB, bench.NT
N = NS * NT
ex, ey, _, _ = synth.pdm_qpsk(NS, NT, 1)
pmx.reset_all(NS, NT, 1)
G = pmx.GSTATE
G.SYMBOLRATE, G.LAMBDA, G.POWER = bench.RATE, np.array([1550.0]), np.array([bench.PAVG])
pmx.create_field('unique', ex, ey, {'power': 'average'})
fib = bench.fiber_params(bench.SPAN_KM * 1e3, bench.NPLATES)
ctx = _lib.default_context()
pinx = torch.empty((N, 1), dtype=torch.complex128).pin_memory()
piny = torch.empty((N, 1), dtype=torch.complex128).pin_memory()
txx, txy = np.array(G.FIELDX_TX), np.array(G.FIELDY_TX)
def step(sid, nspan=10):
    hx, hy = pinx.numpy(), piny.numpy()
    hx[...] = txx; hy[...] = txy
    G.FIELDX, G.FIELDY = hx, hy
    G.DELAY, G.DISP = np.zeros((2, 1)), np.zeros((2, 1))
    sa = 0
    for k in range(nspan):
        pmx.fiber(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(1000 + 100000 * k)), ctx=ctx)
        sa += pmx.FIBER_LAST['ncycle'] * N
        pmx.ampliflat(bench.GAIN_DB, 'gain', {'f': bench.NF_DB}, ctx=ctx, seed=sid * 64 + k)
    assert G.FIELDX is hx
    return sa
step(0); step(1)
t0 = time.perf_counter(); sa = step(2); ctx.sync(); dt = time.perf_counter() - t0
print('e2e %.2f GSa*steps/s, %.1f ms per span' % (sa / dt / 1e9, dt / 10 * 1e3))
pr = cProfile.Profile(); pr.enable(); step(3); ctx.sync(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(30)
