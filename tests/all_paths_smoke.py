"""A few small propagations through every device path in one process (a quick check after a kernel change; the box
this was developed on does not allow compute-sanitizer): one reference-style fiber() call (batch of one, programmatic
dependent launches), a resident batch of five realizations in two groups, FP32, the scalar XPM path, a three-span link
with ASE, the device multiplex and the error counter.  Prints the parity of the first against the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))   # lives under tests/: it uses the oracle as checker
import oracle.fiber_oracle as orc
import polmux_b200 as pmx
from polmux_b200 import _lib, mc, field as fmod
from polmux_b200.fiber import fiber_setup, setup_to_desc
from common import base_fiber, make_tx, rel_l2

fib = base_fiber(length=3e4, dgd=0.5, nplates=8, manakov='no')
gs = make_tx(1 << 8, 16)
orc.fiber(gs, fib, 'gps-', rng=np.random.Generator(np.random.PCG64(1)))
pmx.fiber(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(1)))
G = pmx.GSTATE
print('fiber gps- CNLSE N=2^12: rel_l2 %.2e' % rel_l2(G.FIELDX, G.FIELDY, gs.FIELDX, gs.FIELDY), flush=True)
pmx.ampliflat(5.0, 'gain', {'f': 5.0}, seed=3)
pmx.fiber(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(2)), precision='f32')
print('f32 after ampliflat: finite', bool(np.isfinite(G.FIELDX).all()), flush=True)
# resident batch, two realization groups
make_tx(1 << 9, 16)
fibm = base_fiber(length=3e4, dgd=0.5, nplates=8, manakov='yes')
setup = fiber_setup(fibm, 'gps-', rng=np.random.Generator(np.random.PCG64(0)))
B = 5
d = [mc.draw_plates(10 + b, setup.nplates) for b in range(B)]
pl = [np.stack([x[i] for x in d]) for i in range(3)]
ctx = _lib.default_context()
desc, keep = setup_to_desc(setup, batch=B, plate_sets=B, db0=pl[0], theta=pl[1], epsilon=pl[2])
os.environ.setdefault('PMX_GROUPS', '2')
plan = _lib.Plan(ctx, desc, keep)
work = _lib.DeviceField(ctx, setup.nfft, 1, B)
work.upload(np.stack([np.asarray(G.FIELDX).T] * B), np.stack([np.asarray(G.FIELDY).T] * B))
res = plan.execute(work)
print('batch of 5 in two groups: ncycle', res.ncycle.tolist(), flush=True)
sym = np.zeros((2, 1 << 9), dtype=np.uint8)
import ctypes
cnt = ctypes.c_void_p()
import torch
buf = torch.zeros(B, dtype=torch.int64, device='cuda')
_lib.qpsk_count(ctx, work, sym, 1 << 9, 16, buf.data_ptr())
print('qpsk_count', buf.cpu().tolist(), flush=True)
# link call, device mux, scalar XPM
make_tx(1 << 8, 16)
pmx.link(fibm, 'gps-', 3, 4.0, {'f': 5.0}, rng=np.random.Generator(np.random.PCG64(5)), seed=11)
print('link: finite', bool(np.isfinite(pmx.GSTATE.FIELDX).all()), flush=True)
fmod.DEVICE_MUX = True
make_tx(1 << 8, 64, nch=3)
fmod.DEVICE_MUX = False
pmx.fiber(fibm, 'gps-', rng=np.random.Generator(np.random.PCG64(6)))
print('device mux + fiber: finite', bool(np.isfinite(pmx.GSTATE.FIELDX).all()), flush=True)
gs = make_tx(1 << 8, 16, nch=3, ftype='sepfields')
pmx.GSTATE.FIELDY = None
pmx.fiber(base_fiber(length=2e4), 'g-sx')
print('scalar XPM: finite', bool(np.isfinite(pmx.GSTATE.FIELDX).all()), flush=True)
