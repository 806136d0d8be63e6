"""Monte-Carlo layer: ber_estimate recursion pinned against the interpreted reference, sharding,
the world_size-2 integer all-reduce (gloo on CPU), and on the GPU the link + equaliser + error
counter against the oracle with identical (file-backed) ASE."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle.fiber_oracle as orc
import polmux_b200 as pmx
from polmux_b200 import mc, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'


def test_shard_is_a_contiguous_partition():
    for n in (1, 7, 1024, 1000):
        for w in (1, 2, 4, 8):
            parts = [mc.shard(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))


def test_ber_recursion_closed_form():
    """cumulative mean of the block means = total errors / total bits; stops like ber_estimate"""
    g = np.random.default_rng(3)
    counts = g.binomial(4096, 0.01, size=200)
    st = mc.BerState()
    for e in counts:
        cond, avg, nruns, std = mc.ber_update(st, int(e), 4096, stop=(0.0, 68.0), nmin=10 ** 12)
    assert cond and nruns == 200 * 4096
    assert abs(avg - counts.sum() / (200 * 4096)) < 1e-15
    rep = mc.ber_replay(counts, 4096, stop=(0.1, 95.0), nmin=100)
    assert rep['converged'] and rep['realizations_used'] < 200


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present (GPU box)')
def test_ber_recursion_matches_reference_source():
    """ber_estimate.m executed by the mini interpreter, block by block, same counts"""
    from oracle.mini_m.interp import Interp, MStruct, to_m
    it = Interp(REF)
    g = np.random.default_rng(5)
    M = 512
    pat = g.integers(0, 2, size=(M, 1)).astype(float)
    st = mc.BerState()
    x = MStruct({'stop': to_m(np.array([[0.1, 95.0]])), 'nmin': to_m(50.0)})
    for blk in range(60):
        hat = pat.copy()
        flips = g.random(M) < 0.02
        hat[flips, 0] = 1 - hat[flips, 0]
        cond_r, avg_r, n_r, std_r = it.call('ber_estimate', [to_m(hat), to_m(pat), x], 4)
        cond, avg, n, std = mc.ber_update(st, int(flips.sum()), M, stop=(0.1, 95.0), nmin=50)
        assert abs(float(avg_r.flat[0]) - avg) <= 1e-15 * max(avg, 1e-300)
        assert float(n_r.flat[0]) == n
        assert abs(float(std_r.flat[0]) - std) <= 1e-12 * max(std, 1e-300)
        assert bool(cond_r.flat[0]) == cond
        if not cond:
            break
    assert not cond


WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from polmux_b200 import mc
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo', rank=rank, world_size=world)
nreal = 37
truth = np.random.default_rng(11).integers(0, 50, size=nreal)
r0, r1 = mc.shard(nreal, rank, world)
full = mc.allreduce_counts(truth[r0:r1], r0, nreal).numpy()
assert np.array_equal(full, truth), (rank, full, truth)
rep = mc.ber_replay(full, 4096, stop=(0.1, 68.0), nmin=10)
ref = mc.ber_replay(truth, 4096, stop=(0.1, 68.0), nmin=10)
assert rep == ref
dist.destroy_process_group()
print('ok', rank)
'''


def test_count_allreduce_world2_gloo(tmp_path):
    script = tmp_path / 'w.py'
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29577', WORLD_SIZE='2')
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e
        assert 'ok' in o


def test_philox_known_answer():
    """Random123 known-answer vectors for Philox4x32-10."""
    from oracle.philox import philox4x32_10
    r = philox4x32_10([0], [0], [0], [0], 0, 0)
    assert [int(v[0]) for v in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    r = philox4x32_10([0xffffffff], [0xffffffff], [0xffffffff], [0xffffffff], 0xffffffff, 0xffffffff)
    assert [int(v[0]) for v in r] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    r = philox4x32_10([0x243f6a88], [0x85a308d3], [0x13198a2e], [0x03707344], 0xa4093822, 0x299f31d0)
    assert [int(v[0]) for v in r] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_device_ase_generator_matches_restatement(ctx):
    """ampliflat on a zero field with sigma = 1 exposes the device's Philox/Box-Muller stream"""
    from oracle.philox import ase_normals
    from polmux_b200 import _lib
    n, batch, seed = 4096, 3, 0x1234_5678_9abc
    f = _lib.DeviceField(ctx, n, 1, batch)
    f.upload(np.zeros((batch, 1, n), complex), np.zeros((batch, 1, n), complex))
    _lib.ampliflat_exec(ctx, f, 1.0, [1.0], None, seed=seed)
    x, y = f.download()
    for b in range(batch):
        nx, ny = ase_normals(n, 0, b, seed)
        np.testing.assert_allclose(x[b, 0], nx, rtol=0, atol=1e-12)
        np.testing.assert_allclose(y[b, 0], ny, rtol=0, atol=1e-12)
    assert abs(np.mean(np.abs(x) ** 2) - 2.0) < 0.1      # randn + i*randn: variance 2


@pytest.mark.gpu
def test_mc_counts_match_oracle(ctx):
    """2 spans x (Manakov fiber + noisy amplifier), 3 realizations, identical noise on both sides:
    the integer error counts of the CUDA path equal the oracle's bit for bit."""
    import polmux_b200 as pmx
    from polmux_b200 import _lib, synth
    from polmux_b200.fiber import fiber_setup
    from common import base_fiber
    nsymb, nt, nspan, batch = 256, 16, 2, 3
    n = nsymb * nt
    ex, ey, sx, sy = synth.pdm_qpsk(nsymb, nt, 1)
    pmx.reset_all(nsymb, nt, 1)
    G = pmx.GSTATE
    G.SYMBOLRATE, G.LAMBDA, G.POWER = 28.0, np.array([1550.0]), np.array([2.0])
    pmx.create_field('unique', ex, ey, {'power': 'average'})
    fib = base_fiber(length=8e4, dgd=0.3, nplates=20, manakov='yes')
    setup = fiber_setup(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(0)))
    gain_db, nf_db = 16.0, 38.0          # absurd noise figure: a few percent of the bits flip
    noise = np.random.default_rng(9).standard_normal((nspan, batch, 2, n, 2)).view(np.complex128)[..., 0]
    link = mc.Link(ctx, setup, nspan, batch, gain_db, nf_db, first_realization=5)
    tx = _lib.DeviceField(ctx, n, 1, 1)
    tx.upload(G.FIELDX, G.FIELDY)
    work = _lib.DeviceField(ctx, n, 1, batch)
    work.broadcast_from(tx)
    link.run(work, ase_seed=1, noise_fn=lambda k: noise[k])
    link.equalize(work)
    import torch
    buf = torch.zeros(batch, dtype=torch.int64, device='cuda')
    torch.cuda.synchronize()
    sym = np.stack([sx[:, 0], sy[:, 0]]).astype(np.uint8)
    _lib.qpsk_count(ctx, work, sym, nsymb, nt, buf.data_ptr())
    got = buf.cpu().numpy()

    want = []
    for b in range(batch):
        gs = orc.reset_all(nsymb, nt, 1)
        gs.SYMBOLRATE, gs.LAMBDA, gs.POWER = 28.0, np.array([1550.0]), np.array([2.0])
        orc.create_field(gs, 'unique', ex, ey, power_average=True)
        brfs = []
        for k in range(nspan):
            rng = np.random.Generator(np.random.PCG64(mc.plate_seed(5 + b, k)))
            brfs.append(orc.fiber(gs, fib, 'gps-', rng=rng))
            orc.ampliflat(gs, gain_db, nf_db, noise=noise[k, b].T)
        for brf in reversed(brfs):                        # ideal linear equaliser, span by span
            orc.inverse_pmd(gs, [brf])
        errs = 0
        for fld, s in ((gs.FIELDX[:, 0], sx[:, 0]), (gs.FIELDY[:, 0], sy[:, 0])):
            r = fld[::nt]
            ref = (2 * (s & 1) - 1) + 1j * (2 * (s >> 1) - 1)
            r = r * np.conj(np.sum(r * np.conj(ref)))
            d = (r.real > 0).astype(int) | ((r.imag > 0).astype(int) << 1)
            errs += int(np.sum(((d ^ s) & 1) + (((d ^ s) >> 1) & 1)))
        want.append(errs)
    assert got.tolist() == want
    assert min(want) > 0


@pytest.mark.gpu
def test_native_mc_run_matches_the_torch_driven_path(ctx):
    """pmx_mc_run (the Monte-Carlo job behind the C ABI: host threads + contexts + NCCL bound at run time inside the
    library) gives the error counts of mc.run_mc bit for bit, on one GPU and -- when the box has more -- sharded over
    two with the counts all-reduced by NCCL"""
    from polmux_b200 import _lib
    from polmux_b200.fiber import fiber_setup
    nsymb, nt, nspan, nreal, batch = 1 << 10, 16, 3, 6, 2
    ex, ey, sx, sy = synth.pdm_qpsk(nsymb, nt, 1)
    pmx.reset_all(nsymb, nt, 1)
    G = pmx.GSTATE
    G.SYMBOLRATE, G.LAMBDA, G.POWER = 28.0, np.array([1550.0]), np.array([8.0])
    pmx.create_field('unique', ex, ey, {'power': 'average'})
    fib = dict(synth.SMF)
    fib.update(length=8e4, dgd=0.1, nplates=20, manakov='yes')
    setup = fiber_setup(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(0)))
    sym = np.stack([sx[:, 0], sy[:, 0]]).astype(np.uint8)
    ref, sa_ref = mc.run_mc(ctx, setup, G.FIELDX_TX, G.FIELDY_TX, sym, nsymb, nt, nspan, 16.0, 18.0, nreal, batch, ase_seed=5)
    got, sa = mc.run_mc_native(setup, G.FIELDX_TX, G.FIELDY_TX, sym, nsymb, nt, nspan, 16.0, 18.0, nreal, batch,
                               devices=(0,), ase_seed=5)
    assert np.array_equal(got, ref) and sa == sa_ref and ref.sum() > 0
    assert _lib.load().pmx_mc_nccl_available() == 1
    # the ASE generator is keyed by the global realization index: another grouping of the same realizations, same counts
    got3, _ = mc.run_mc_native(setup, G.FIELDX_TX, G.FIELDY_TX, sym, nsymb, nt, nspan, 16.0, 18.0, nreal, 3,
                               devices=(0,), ase_seed=5)
    assert np.array_equal(got3, ref)
    import torch
    if torch.cuda.device_count() >= 2:
        got2, sa2 = mc.run_mc_native(setup, G.FIELDX_TX, G.FIELDY_TX, sym, nsymb, nt, nspan, 16.0, 18.0, nreal, batch,
                                     devices=(0, 1), ase_seed=5)
        assert np.array_equal(got2, ref) and sa2 == sa_ref


@pytest.mark.gpu
def test_mc_with_the_blind_receiver(ctx):
    """McRunner(receiver='blind'): link -> dispersion compensation -> CMA + Viterbi & Viterbi + differential decision
    (pmx_dsp_count) instead of the genie equaliser; at a comfortable OSNR both receivers count (almost) no errors, and
    the blind one does not read the plates"""
    from polmux_b200.fiber import fiber_setup
    nsymb, nt, nspan, nreal, batch = 1 << 11, 16, 2, 4, 2
    ex, ey, sx, sy = synth.pdm_qpsk(nsymb, nt, 1)
    pmx.reset_all(nsymb, nt, 1)
    G = pmx.GSTATE
    G.SYMBOLRATE, G.LAMBDA, G.POWER = 28.0, np.array([1550.0]), np.array([1.0])
    pmx.create_field('unique', ex, ey, {'power': 'average'})
    fib = dict(synth.SMF)
    fib.update(length=4e4, dgd=0.3, nplates=20, manakov='yes')
    setup = fiber_setup(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(0)))
    sym = np.stack([sx[:, 0], sy[:, 0]]).astype(np.uint8)
    out = {}
    for rec in ('genie', 'blind', 'cohmix'):
        r = mc.McRunner(ctx, setup, G.FIELDX_TX, G.FIELDY_TX, sym, nsymb, nt, nspan, 8.0, 5.0, nreal, batch, receiver=rec,
                        dsp_params=dict(mu=1 / 2000, freqavg=200))
        out[rec], _ = r.run(ase_seed=9)
        if rec != 'genie':
            assert len(r.passes) == nreal // batch and all(p.min() >= 1 for p in r.passes)
            # the chain beside the next group's propagation (default) or after its own link: the same counts
            q = mc.McRunner(ctx, setup, G.FIELDX_TX, G.FIELDY_TX, sym, nsymb, nt, nspan, 8.0, 30.0, nreal + 1, batch, receiver=rec,
                            dsp_params=dict(mu=1 / 2000, freqavg=200), pipeline=False)
            seq, _ = q.run(ase_seed=9)
            q.close()
            q = mc.McRunner(ctx, setup, G.FIELDX_TX, G.FIELDY_TX, sym, nsymb, nt, nspan, 8.0, 30.0, nreal + 1, batch, receiver=rec,
                            dsp_params=dict(mu=1 / 2000, freqavg=200))
            par, _ = q.run(ase_seed=9)
            par2, _ = q.run(ase_seed=9)
            q.close()
            assert q.pipelined and np.array_equal(seq, par) and np.array_equal(par, par2)
        r.close()
    assert out['genie'].sum() == 0
    assert out['blind'].sum() <= 8          # differential decoding doubles isolated errors; none expected here
    # the full front-end of receiver_cohmix (gauss 1.9 / bessel5 0.65, ex20_coherent_polmux.m:47-50) in front of the same DSP
    assert out['cohmix'].sum() <= 8


@pytest.mark.gpu
@pytest.mark.parametrize('nf_db', [5.0, 37.0])
def test_native_mc_run_with_the_receive_chain(ctx, nf_db):
    """pmx_mc_run with pmx_mc_desc.rx (receiver_cohmix front-end + sampler + DSP core inside the library's Monte-Carlo
    job) against McRunner(receiver='cohmix') driving the same pieces from Python: equal counts per realization"""
    from polmux_b200.fiber import fiber_setup
    nsymb, nt, nspan, nreal, batch = 1 << 11, 16, 2, 5, 2
    ex, ey, sx, sy = synth.pdm_qpsk(nsymb, nt, 1)
    pmx.reset_all(nsymb, nt, 1)
    G = pmx.GSTATE
    G.SYMBOLRATE, G.LAMBDA, G.POWER = 28.0, np.array([1550.0]), np.array([1.0])
    pmx.create_field('unique', ex, ey, {'power': 'average'})
    fib = dict(synth.SMF)
    fib.update(length=4e4, dgd=0.3, nplates=20, manakov='yes')
    setup = fiber_setup(fib, 'gps-', rng=np.random.Generator(np.random.PCG64(0)))
    sym = np.stack([sx[:, 0], sy[:, 0]]).astype(np.uint8)
    x = {'oftype': 'gauss', 'obw': 1.9, 'eftype': 'bessel5', 'ebw': 0.65, 'lopower': 0.0}
    dspp = dict(mu=1 / 2000, freqavg=200)
    r = mc.McRunner(ctx, setup, G.FIELDX_TX, G.FIELDY_TX, sym, nsymb, nt, nspan, 8.0, nf_db, nreal, batch, receiver='cohmix',
                    dsp_params=dspp, rx_params=x)
    want, sa = r.run(ase_seed=9)
    r.close()
    got, sa2 = mc.run_mc_native(setup, G.FIELDX_TX, G.FIELDY_TX, sym, nsymb, nt, nspan, 8.0, nf_db, nreal, batch, devices=(0,),
                                ase_seed=9, receiver=x, dsp_params=dspp)
    assert 0 < sa2 <= sa          # (McRunner also counts the padded slot of the ragged last group)
    assert np.array_equal(got, want), (got, want)
    assert (want.sum() > 0) == (nf_db > 20)
