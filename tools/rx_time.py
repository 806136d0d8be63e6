"""Receive chain on the C2 link (2^16 symbols x 16 samples, batch 16): time of its parts per batch, and Monte-Carlo
realizations per second with the three receivers (genie / blind / cohmix)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import polmux_b200 as pmx
from polmux_b200 import _lib, synth, mc, dsp
from polmux_b200.fiber import fiber_setup
import torch

NSYMB, NT, B = bench.NSYMB, bench.NT, 16
ex, ey, sx, sy = synth.pdm_qpsk(NSYMB, NT, 1)
pmx.reset_all(NSYMB, NT, 1)
G = pmx.GSTATE
G.SYMBOLRATE, G.LAMBDA, G.POWER = bench.RATE, np.array([1550.0]), np.array([bench.PAVG])
pmx.create_field('unique', ex, ey, {'power': 'average'})
setup = fiber_setup(bench.fiber_params(8e4, 100), 'gps-', rng=np.random.Generator(np.random.PCG64(0)))
sym = np.stack([sx[:, 0], sy[:, 0]]).astype(np.uint8)
ctx = _lib.Context(0)
nreal = int(sys.argv[1]) if len(sys.argv) > 1 else 32


def timed(fn, n=1):
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    ctx.sync()
    return (time.perf_counter() - t0) / n


r = mc.McRunner(ctx, setup, G.FIELDX_TX, G.FIELDY_TX, sym, NSYMB, NT, bench.NSPAN, bench.GAIN_DB, bench.NF_DB, B, B,
                receiver='cohmix', pipeline=False)
r.link.retarget(0)
r.work.broadcast_from(r.tx)
t_link = timed(lambda: r.link.run(r.work, 3))
t_cd = timed(lambda: r.link.cd_compensate(r.work))
R, S = r.rx, r.rx['S']
keep = _lib.DeviceField(ctx, setup.nfft, 1, B)
_lib.field_copy_cols(keep, 0, r.work, 0, B)
t_fo = timed(lambda: R['fo'].execute(r.work), 3)
_lib.field_copy_cols(r.work, 0, keep, 0, B)
R['fo'].execute(r.work)
t_mix = timed(lambda: _lib.cohmix_exec(ctx, r.work, S.ecw, S.detune, S.lophase, S.balanced))
t_fe = timed(lambda: R['fe'].execute(r.work))
buf = torch.zeros(B, dtype=torch.int64, device='cuda')
for mu, mp in ((1 / 6000, 0), (1 / 6000, 1)):
    passes = []
    t = timed(lambda: passes.append(dsp.dsp_count(ctx, r.work, NSYMB, NT, r.ref_patmat, buf.data_ptr(), sample_shift=R['shift'],
                                                  peak=R['peak'], mu=mu, max_passes=mp)))
    print('dsp_count mu=%g max_passes=%d: %.1f ms per batch of %d, CMA passes %s, errors %s' % (
        mu, mp, t * 1e3, B, passes[-1].tolist()[:6], buf.cpu().numpy().tolist()[:6]), flush=True)
print('per batch of %d: link %.1f ms, CD compensation %.2f ms, optical filter %.2f ms, LO mixing + photodiodes %.2f ms, '
      'low-pass filter %.2f ms' % (B, t_link * 1e3, t_cd * 1e3, t_fo * 1e3, t_mix * 1e3, t_fe * 1e3), flush=True)
keep.close()
r.close()
for rec, dp, pipe in (('genie', None, True), ('blind', None, False), ('blind', None, True), ('cohmix', None, False),
                      ('cohmix', None, True), ('cohmix', dict(applyeasi=True), True),
                      ('cohmix', dict(applyeasi=True, applypol=False), True)):
    r = mc.McRunner(ctx, setup, G.FIELDX_TX, G.FIELDY_TX, sym, NSYMB, NT, bench.NSPAN, bench.GAIN_DB, bench.NF_DB, nreal, B,
                    receiver=rec, dsp_params=dp, pipeline=pipe)
    r.run(ase_seed=3)
    t0 = time.perf_counter()
    counts, _ = r.run(ase_seed=3)
    dt = time.perf_counter() - t0
    tag = rec + ('' if not dp else (' (combo: EASI + CMA)' if dp.get('applypol', True) else ' (EASI)')) + \
        ('' if rec == 'genie' else (', chain beside the next link' if pipe else ', chain after the link'))
    print('%s receiver: %d realizations in %.2f s = %.1f realizations/s, errors in total %d, realizations with more than '
          'NSYMB/2 errors %d, first counts %s' % (tag, nreal, dt, nreal / dt, int(counts.sum()), int((counts > NSYMB // 2).sum()),
                                                    counts.tolist()[:8]), flush=True)
    r.close()
