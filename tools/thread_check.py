import faulthandler, sys, threading, os
faulthandler.enable()
sys.path.insert(0, os.getcwd())
import numpy as np
import bench
import polmux_b200 as pmx
from polmux_b200 import _lib, synth, mc
from polmux_b200.fiber import fiber_setup, setup_to_desc
nsymb, nt = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 12), 16
N = nsymb * nt
ex, ey, _, _ = synth.pdm_qpsk(nsymb, nt, 1)
pmx.reset_all(nsymb, nt, 1)
G = pmx.GSTATE
G.SYMBOLRATE, G.LAMBDA, G.POWER = bench.RATE, np.array([1550.0]), np.array([bench.PAVG])
pmx.create_field('unique', ex, ey, {'power': 'average'})
setup = fiber_setup(bench.fiber_params(8e4, 100), 'gps-', rng=np.random.Generator(np.random.PCG64(0)))
def job(tag):
    ctx = _lib.Context(0)
    d = [mc.draw_plates(1000 + b, 100) for b in range(4)]
    pl = [np.stack([x[i] for x in d]) for i in range(3)]
    desc, keep = setup_to_desc(setup, batch=4, plate_sets=4, db0=pl[0], theta=pl[1], epsilon=pl[2])
    plan = _lib.Plan(ctx, desc, keep)
    tx = _lib.DeviceField(ctx, N, 1, 1); tx.upload(G.FIELDX, G.FIELDY)
    work = _lib.DeviceField(ctx, N, 1, 4)
    for _ in range(3):
        work.broadcast_from(tx); res = plan.execute(work)
    ctx.sync()
    print(tag, 'ncycle', res.ncycle.tolist(), flush=True)
job('main')
t = threading.Thread(target=job, args=('thread',)); t.start(); t.join()
ts = [threading.Thread(target=job, args=('thread%d' % i,)) for i in range(3)]
[t.start() for t in ts]; [t.join() for t in ts]
print('done', flush=True)
