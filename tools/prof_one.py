"""Short driver for ncu: one C2 span (2^20 samples, 'gps-' Manakov, 100 plates) on a batch of
realizations resident in HBM.  Usage: python tools/prof_one.py [batch] [log2N]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import polmux_b200 as pmx  # noqa: E402
from polmux_b200 import _lib, synth  # noqa: E402
from polmux_b200.fiber import fiber_setup, setup_to_desc  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
PREC = sys.argv[3] if len(sys.argv) > 3 else "f64"
LG = int(sys.argv[2]) if len(sys.argv) > 2 else 20
nsymb, nt = 1 << (LG - 4), 16
N = nsymb * nt
ex, ey, _, _ = synth.pdm_qpsk(nsymb, nt, 1)
pmx.reset_all(nsymb, nt, 1)
G = pmx.GSTATE
G.SYMBOLRATE, G.LAMBDA, G.POWER = bench.RATE, np.array([1550.0]), np.array([bench.PAVG])
pmx.create_field('unique', ex, ey, {'power': 'average'})
fib = bench.fiber_params(bench.SPAN_KM * 1e3, bench.NPLATES)
setup = fiber_setup(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(0)))
from polmux_b200 import mc  # noqa: E402
d = [mc.draw_plates(1000 + b, bench.NPLATES) for b in range(B)]
pl = [np.stack([x[i] for x in d]) for i in range(3)]
ctx = _lib.Context(0)
desc, keep = setup_to_desc(setup, batch=B, plate_sets=B, db0=pl[0], theta=pl[1], epsilon=pl[2], precision=PREC)
PC = {'f64': _lib.PMX_F64, 'f32': _lib.PMX_F32}[PREC]
plan = _lib.Plan(ctx, desc, keep)
tx = _lib.DeviceField(ctx, N, 1, 1, precision=PC)
tx.upload(G.FIELDX, G.FIELDY)
work = _lib.DeviceField(ctx, N, 1, B, precision=PC)
work.broadcast_from(tx)
res = plan.execute(work)
ctx.sync()
print('ncycle', res.ncycle.tolist(), 'launches', ctx.launches)
