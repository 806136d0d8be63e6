"""Where the Monte-Carlo leg of bench.py spends its host time (one GPU): python tools/mc_profile.py"""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
import polmux_b200 as pmx
from polmux_b200 import _lib, synth, mc
from polmux_b200.fiber import fiber_setup
NS, NT = bench.NSYMB, bench.NT
ex, ey, symx, symy = synth.pdm_qpsk(NS, NT, 1)
pmx.reset_all(NS, NT, 1)
G = pmx.GSTATE
G.SYMBOLRATE, G.LAMBDA, G.POWER = bench.RATE, np.array([1550.0]), np.array([bench.PAVG])
pmx.create_field('unique', ex, ey, {'power': 'average'})
fib = bench.fiber_params(bench.SPAN_KM * 1e3, bench.NPLATES)
setup = fiber_setup(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(0)))
ctx = _lib.Context(0)
sym = np.stack([symx[:, 0], symy[:, 0]]).astype(np.uint8)
B = 8
def go():
    return mc.run_mc(ctx, setup, G.FIELDX_TX, G.FIELDY_TX, sym, NS, NT, bench.NSPAN, bench.GAIN_DB, bench.NF_DB, B, B, 0, 1, ase_seed=7)
go(); ctx.sync()
for i in range(3):
    t0 = time.perf_counter(); c, sa = go(); ctx.sync(); dt = time.perf_counter() - t0
    print('run_mc: %.3f s for %d realizations (%.1f /s), link alone would be %.3f s at 20.6 GSa*steps/s' % (dt, B, B / dt, sa / 20.6e9), flush=True)
pr = cProfile.Profile(); pr.enable(); go(); ctx.sync(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
