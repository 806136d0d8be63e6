import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np
import oracle.fiber_oracle as orc
import polmux_b200 as pmx
from polmux_b200 import gstate
from common import base_fiber, make_tx, rel_l2
fib = base_fiber(length=1e5, dgd=1.0, nplates=10, manakov='no')
for resident in (True, False):
    gstate.RESIDENT = resident
    ts = []
    for i in range(6):
        gs = make_tx(1 << 12, 16)
        G = pmx.GSTATE
        t0 = time.perf_counter()
        pmx.fiber(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(1000)))
        x = G.FIELDX
        ts.append(time.perf_counter() - t0)
    print('C1 fiber() call incl. H2D/D2H, resident=%s: best %.2f ms, median %.2f ms, ncycle %d' % (resident, min(ts) * 1e3, sorted(ts)[3] * 1e3, pmx.FIBER_LAST['ncycle']))
t0 = time.perf_counter()
orc.fiber(gs, fib, 'gps-', rng=np.random.Generator(np.random.PCG64(1000)))
print('numpy oracle, one core: %.2f s; rel_l2 %.1e' % (time.perf_counter() - t0, rel_l2(G.FIELDX, G.FIELDY, gs.FIELDX, gs.FIELDY)))
