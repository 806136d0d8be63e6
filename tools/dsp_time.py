"""Monte-Carlo realizations per second with the genie receiver and with the blind DSP core (C2 link, batch 16)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import polmux_b200 as pmx
from polmux_b200 import _lib, synth, mc
from polmux_b200.fiber import fiber_setup

NSYMB, NT = bench.NSYMB, bench.NT
ex, ey, sx, sy = synth.pdm_qpsk(NSYMB, NT, 1)
pmx.reset_all(NSYMB, NT, 1)
G = pmx.GSTATE
G.SYMBOLRATE, G.LAMBDA, G.POWER = bench.RATE, np.array([1550.0]), np.array([bench.PAVG])
pmx.create_field('unique', ex, ey, {'power': 'average'})
setup = fiber_setup(bench.fiber_params(8e4, 100), 'gps-', rng=np.random.Generator(np.random.PCG64(0)))
sym = np.stack([sx[:, 0], sy[:, 0]]).astype(np.uint8)
ctx = _lib.Context(0)
for rec in ('genie', 'blind'):
    r = mc.McRunner(ctx, setup, G.FIELDX_TX, G.FIELDY_TX, sym, NSYMB, NT, bench.NSPAN, bench.GAIN_DB, bench.NF_DB, 32, 16,
                    receiver=rec)
    r.run(ase_seed=3)
    t0 = time.perf_counter()
    counts, _ = r.run(ase_seed=3)
    dt = time.perf_counter() - t0
    print('%s receiver: 32 realizations in %.2f s = %.1f realizations/s, errors %s, CMA passes %s' % (
        rec, dt, 32 / dt, counts.tolist()[:8], [p.tolist()[:4] for p in r.passes[-1:]]), flush=True)
    r.close()
