"""Generate tests/golden/*.npz by EXECUTING the reference's own M source (fiber.m, fastexp.m,
reset_all.m, create_field.m, fastshift.m, nmod.m, checkfields.m, ampliflat.m under
/root/reference) with the minimal interpreter in oracle/mini_m.  Run in the build container
(the reference tree does not exist on the GPU box):

    python oracle/make_golden.py [/root/reference]

Inputs are the seeded synthetic PDM-QPSK fields of polmux_b200.synth; random plates come
from numpy Generator(PCG64(seed)) in the order fiber.m:274-276 draws them.  Each fixture
stores inputs, parameters and the reference outputs, so the tests need nothing else.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.mini_m.interp import Interp, MCell, MStruct, from_m, to_m  # noqa: E402
from polmux_b200 import synth  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
OUT = os.path.join(ROOT, 'tests', 'golden')

SMF = dict(synth.SMF)


def new_interp(seed):
    it = Interp(REF, rng=np.random.Generator(np.random.PCG64(seed)))
    return it


def tx_through_reference(it, nsymb, nt, nch, rate, pavg, ftype, two_pol=True, spac=0.4):
    """reset_all.m + (lasersource/electricsource stand-ins) + create_field.m, all interpreted."""
    ex, ey, _, _ = synth.pdm_qpsk(nsymb, nt, nch)
    it.call('reset_all', [to_m(nsymb), to_m(nt), to_m(nch)], 0)
    G = it.globals['GSTATE']
    G = G.copy()
    G['SYMBOLRATE'] = to_m(rate)                                  # electricsource.m sets it
    G['LAMBDA'] = to_m(synth.wdm_lambdas(nch, 1550.0, spac).reshape(1, -1))   # lasersource.m:163
    G['POWER'] = to_m(np.full((1, nch), float(pavg)))
    it.globals['GSTATE'] = G
    opts = MStruct({'power': 'average'})
    args = [ftype, to_m(ex), to_m(ey) if two_pol else np.zeros((0, 0)), opts]
    it.call('create_field', args, 0)
    return ex, ey


def snapshot(it):
    G = it.globals['GSTATE']
    out = {'FIELDX': np.array(G['FIELDX'])}
    fy = G['FIELDY']
    out['FIELDY'] = np.array(fy) if isinstance(fy, np.ndarray) and fy.size else np.zeros((0, 0))
    out['DELAY'] = np.array(G['DELAY'])
    out['DISP'] = np.array(G['DISP'])
    out['POWER'] = np.array(G['POWER'])
    return out


def run_case(name, nsymb, nt, nch, ftype, fib, flag, seed=1000, rate=28.0, pavg=2.0, two_pol=True, want_brf=True,
             amp=None):
    it = new_interp(seed)
    ex, ey = tx_through_reference(it, nsymb, nt, nch, rate, pavg, ftype, two_pol)
    pre = snapshot(it)
    x = to_m(fib)
    res = it.call('fiber', [x, flag], 1 if want_brf else 0)
    post = snapshot(it)
    data = {'in_ex': ex, 'in_ey': ey, 'tx_FIELDX': pre['FIELDX'], 'tx_FIELDY': pre['FIELDY'], 'tx_POWER': pre['POWER'],
            'out_FIELDX': post['FIELDX'], 'out_FIELDY': post['FIELDY'], 'out_DELAY': post['DELAY'],
            'out_DISP': post['DISP']}
    meta = {'name': name, 'nsymb': nsymb, 'nt': nt, 'nch': nch, 'ftype': ftype, 'fiber': fib, 'flag': flag,
            'seed': seed, 'rate': rate, 'pavg': pavg, 'two_pol': two_pol, 'warnings': it.warnings}
    if want_brf and res:
        brf = from_m(res[0])
        for k in ('db0', 'theta', 'epsilon'):
            data['brf_' + k] = np.asarray(brf[k]).ravel()
        data['brf_lcorr'] = np.asarray(brf['lcorr']).ravel()
        data['brf_betat'] = np.asarray(brf['betat'])
        data['brf_db1'] = np.asarray(brf['db1'])
    if amp is not None:                       # span boundary: ampliflat with file-backed noise (options.noise)
        g = np.random.Generator(np.random.PCG64(seed + 7))
        n = nsymb * nt
        nfc = post['FIELDX'].shape[1]
        noise = (g.standard_normal((n, 2 * nfc)) + 1j * g.standard_normal((n, 2 * nfc)))
        opt = MStruct({'f': to_m(amp['f']), 'noise': to_m(noise)})
        it.call('ampliflat', [to_m(amp['gain']), amp.get('atype', 'gain'), opt], 0)   # ('fixpower': gain = mW)
        a = snapshot(it)
        data.update(amp_noise=noise, amp_FIELDX=a['FIELDX'], amp_FIELDY=a['FIELDY'])
        meta['amp'] = amp
    data['meta'] = np.array(json.dumps(meta))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **data)
    print('%-28s N=%-6d flag=%s  |out|=%.6e' % (name, nsymb * nt, flag, np.linalg.norm(post['FIELDX'])))


def run_print_case(name, nsymb, nt, nch, ftype, fib, flag, seed=1000, rate=28.0, pavg=2.0, two_pol=True):
    """the simul_out block fiber.m:392-456 writes when GSTATE.PRINT is set, captured from the interpreted reference
    (fprintf of the mini interpreter) -> tests/golden/simul_out_<name>.json {params, firstdz/ncycle are in the text}"""
    it = new_interp(seed)
    tx_through_reference(it, nsymb, nt, nch, rate, pavg, ftype, two_pol)
    G = it.globals['GSTATE'].copy()
    G['PRINT'] = to_m(True)
    G['DIR'] = 'sim'
    it.globals['GSTATE'] = G
    it.printed = []
    it.call('fiber', [to_m(fib), flag], 0)
    text = ''.join(it.printed)
    meta = {'name': name, 'nsymb': nsymb, 'nt': nt, 'nch': nch, 'ftype': ftype, 'fiber': fib, 'flag': flag, 'seed': seed,
            'rate': rate, 'pavg': pavg, 'two_pol': two_pol, 'text': text}
    with open(os.path.join(OUT, 'simul_out_' + name + '.json'), 'w') as f:
        json.dump(meta, f, indent=1)
    print('simul_out_%s: %d characters' % (name, len(text)))


def fib(**kw):
    f = dict(SMF)
    f.update(kw)
    return f


def fixpower_cases():
    """ampliflat(x,'fixpower',options): output power of the middle channel = x [mW], through the interpreted ampliflat.m
    and avg_power.m; separate channels, two polarizations and the scalar (FIELDY empty) field"""
    run_case('fixpower_sep3_manakov', 64, 16, 3, 'sepfields', fib(length=4e4, dgd=0.3, nplates=8, manakov='yes',
                                                                  slope=0.057), 'gps-',
             amp={'gain': 1.5, 'f': 5.0, 'atype': 'fixpower'})
    run_case('fixpower_scalar_sep3_gsx', 64, 16, 3, 'sepfields', fib(length=3e4, slope=0.057), 'g-sx', two_pol=False,
             want_brf=False, amp={'gain': 0.8, 'f': 6.0, 'atype': 'fixpower'})


def inverse_pmd_cases(name='invpmd_two_fibers'):
    """inverse_pmd.m interpreted, after two fibers 'gp--' (5 and 8 plates): without options, with options.mat +
    options.gvd = 'no', and with options.apply = 'no' (matrices only) -> tests/golden/invpmd/<name>.npz"""
    nsymb, nt, seed = 64, 8, 1000
    it = new_interp(seed)
    ex, ey = tx_through_reference(it, nsymb, nt, 1, 28.0, 2.0, 'unique', True)
    brfs = []
    for f in (fib(length=2e4, dgd=0.5, nplates=5), fib(length=3e4, dgd=0.8, nplates=8, disp=4.0)):
        brfs.append(it.call('fiber', [to_m(f), 'gp--'], 1)[0])
    after = snapshot(it)
    mat = np.array([[0.6, 0.8j], [0.8j, 0.6]]) * np.exp(0.3j)     # unitary, det != 1
    data = {'in_ex': ex, 'in_ey': ey, 'prop_FIELDX': after['FIELDX'], 'prop_FIELDY': after['FIELDY'], 'mat': mat}
    for k, b in enumerate(brfs):
        bb = from_m(b)
        for key in ('db0', 'theta', 'epsilon'):
            data['brf%d_%s' % (k, key)] = np.asarray(bb[key]).ravel()
        data['brf%d_lcorr' % k] = np.asarray(bb['lcorr']).ravel()
        data['brf%d_betat' % k] = np.asarray(bb['betat'])
        data['brf%d_db1' % k] = np.asarray(bb['db1'])
    variants = {'plain': None, 'mat_nogvd': MStruct({'mat': to_m(mat), 'gvd': 'no'}), 'noapply': MStruct({'apply': 'no'}),
                'apply_n': MStruct({'apply': 'n', 'mat': to_m(mat)})}
    for tag, opt in variants.items():
        G = it.globals['GSTATE'].copy()
        G['FIELDX'], G['FIELDY'] = np.array(after['FIELDX']), np.array(after['FIELDY'])
        G['DISP'] = np.array(after['DISP'])
        it.globals['GSTATE'] = G
        args = [MCell(brfs)] + ([opt] if opt is not None else [])
        uinv, u = it.call('inverse_pmd', args, 2)
        o = snapshot(it)
        data.update({tag + '_Uinv': np.asarray(uinv), tag + '_U': np.asarray(u), tag + '_FIELDX': o['FIELDX'],
                     tag + '_FIELDY': o['FIELDY'], tag + '_DISP': o['DISP']})
    data['meta'] = np.array(json.dumps({'name': name, 'nsymb': nsymb, 'nt': nt, 'seed': seed, 'rate': 28.0, 'pavg': 2.0}))
    os.makedirs(os.path.join(OUT, 'invpmd'), exist_ok=True)
    np.savez_compressed(os.path.join(OUT, 'invpmd', name + '.npz'), **data)
    print('%-28s N=%-6d variants=%s' % (name, nsymb * nt, list(variants)))


def rx_cases():
    """receiver_cohmix.m (+ myfilter.m, evaldelay.m) interpreted -> tests/golden/rx/*.npz: separate channels with two
    polarizations; three channels in one field with post-compensation, LO detuning and phase noise; one polarization with
    single photodiodes; back-to-back"""
    os.makedirs(os.path.join(OUT, 'rx'), exist_ok=True)
    g = np.random.Generator(np.random.PCG64(77))
    cases = {
        'rx_sep3_2pol': dict(nsymb=64, nt=16, nch=3, ftype='sepfields', two_pol=True, ich=2,
                             x={'oftype': 'gauss', 'obw': 1.9, 'eftype': 'bessel5', 'ebw': 0.65, 'lopower': 0.0}),
        'rx_unique3_post': dict(nsymb=64, nt=64, nch=3, ftype='unique', two_pol=True, ich=3,
                                x={'oftype': 'supergauss', 'oord': 2, 'obw': 1.6, 'eftype': 'butt2', 'ebw': 0.7, 'lopower': 3.0,
                                   'lodetuning': 3.2e9, 'dpost': -340.0, 'slopez': 1.5, 'lambda': 1550.0, 'lophasenoise': 'PN'}),
        'rx_unique3_mid': dict(nsymb=64, nt=64, nch=3, ftype='unique', two_pol=True, ich=2,
                               x={'oftype': 'butt4', 'obw': 1.8, 'eftype': 'rc2', 'ebw': 0.8}),
        'rx_scalar_normal': dict(nsymb=128, nt=8, nch=1, ftype='unique', two_pol=False, ich=1,
                                 x={'oftype': 'butt6', 'obw': 2.0, 'eftype': 'butt4', 'ebw': 0.75, 'pdtype': 'normal',
                                    'lopower': -2.0}),
        'rx_b2b': dict(nsymb=64, nt=16, nch=1, ftype='unique', two_pol=True, ich=1,
                       x={'oftype': 'gauss', 'obw': 1.9, 'eftype': 'bessel5', 'ebw': 0.65, 'b2b': 'b2b', 'dpost': 100.0,
                          'slopez': 0.0, 'lambda': 1550.0}),
    }
    for name, c in cases.items():
        it = new_interp(1000)
        n = c['nsymb'] * c['nt']
        ex, ey = tx_through_reference(it, c['nsymb'], c['nt'], c['nch'], 28.0, 2.0, c['ftype'], c['two_pol'])
        it.call('fiber', [to_m(fib(length=1.5e4)), 'g-s-' if c['two_pol'] else 'g-s-'], 0)
        pre = snapshot(it)
        xs = dict(c['x'])
        pn = None
        if xs.get('lophasenoise') == 'PN':
            pn = np.cumsum(0.02 * g.standard_normal(n))
            xs['lophasenoise'] = pn.reshape(-1, 1)
        xm = MStruct({k: (v if isinstance(v, str) else to_m(v)) for k, v in xs.items()})
        iric, xo = it.call('receiver_cohmix', [to_m(float(c['ich'])), xm], 2)
        post = snapshot(it)
        assert np.array_equal(post['FIELDX'], pre['FIELDX'])          # GSTATE is left unchanged
        G = it.globals['GSTATE']
        data = {'FIELDX': pre['FIELDX'], 'FIELDY': pre['FIELDY'], 'DELAY': pre['DELAY'], 'POWER': pre['POWER'],
                'FIELDX_TX': np.array(G['FIELDX_TX']),
                'FIELDY_TX': np.array(G['FIELDY_TX']) if np.size(G['FIELDY_TX']) else np.zeros((0, 0)),
                'Iric': np.asarray(iric), 'avgebx': np.asarray(xo['avgebx']).ravel(),
                'avgeby': np.asarray(xo['avgeby']).ravel() if 'avgeby' in xo else np.zeros(0),
                'post_delay': np.asarray(xo['post_delay'], dtype=np.float64).ravel()}
        if pn is not None:
            data['lophasenoise'] = pn
        meta = {k: v for k, v in c.items() if k != 'x'}
        meta.update(name=name, rate=28.0, pavg=2.0, seed=1000, x={k: v for k, v in c['x'].items()})
        data['meta'] = np.array(json.dumps(meta))
        np.savez_compressed(os.path.join(OUT, 'rx', name + '.npz'), **data)
        print('%-28s N=%-6d Iric %s  |Iric|=%.6e' % (name, n, np.shape(iric), np.linalg.norm(np.asarray(iric))))
    # myfilter.m / evaldelay.m over a frequency grid, every filter type
    it = new_interp(1)
    f = np.fft.fftshift(-8 + np.arange(512) / 32.0).reshape(1, -1)
    out = {'f': f.ravel()}
    for ft, bw, od in (('movavg', 0.9, 0), ('gauss', 0.95, 0), ('gauss_off', 0.95, 0.3), ('butt2', 0.7, 0), ('butt4', 0.7, 0),
                       ('butt6', 0.7, 0), ('ideal', 1.1, 0), ('bessel5', 0.65, 0), ('rc1', 0.8, 0), ('rc2', 0.8, 0),
                       ('supergauss', 0.9, 3)):
        out['H_' + ft] = np.asarray(it.call('myfilter', [ft, to_m(f), to_m(bw), to_m(float(od))], 1)[0]).ravel()
        out['D_' + ft] = np.asarray(it.call('evaldelay', [ft, to_m(bw)], 1)[0], dtype=np.float64).ravel()
        out['P_' + ft] = np.array([bw, od], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, 'rx', 'myfilter_all.npz'), **out)
    print('myfilter_all: %d filter types' % (len(out) // 3))


def run_c1_case(name='c1_cnlse_10plates_100km_2e16'):
    """BASELINE config C1 at its full size (Run_my_PDM_QPSK: 2^12 symbols x 16 samples = 2^16 samples, 28 GBaud, 100 km SMF,
    'gps-' CNLSE, 10 plates, x.dgd = 1.0 -- the value Run_my_PDM_QPSK.m:46 evaluates to) through the interpreted
    create_field.m + fiber.m.  Stored compactly under tests/golden/big/: the outputs, the plate draw and a checksum of
    the seeded inputs (the tests regenerate them), not the 2^16-sample inputs themselves."""
    import hashlib
    nsymb, nt, seed = 1 << 12, 16, 1000
    f = fib(length=1e5, dgd=1.0, nplates=10, manakov='no')
    it = new_interp(seed)
    ex, ey = tx_through_reference(it, nsymb, nt, 1, 28.0, 2.0, 'unique', True)
    pre = snapshot(it)
    res = it.call('fiber', [to_m(f), 'gps-'], 1)
    post = snapshot(it)
    brf = from_m(res[0])
    meta = {'name': name, 'nsymb': nsymb, 'nt': nt, 'nch': 1, 'ftype': 'unique', 'fiber': f, 'flag': 'gps-', 'seed': seed,
            'rate': 28.0, 'pavg': 2.0, 'two_pol': True, 'warnings': it.warnings,
            'sha256_in': hashlib.sha256(np.ascontiguousarray(ex).tobytes() + np.ascontiguousarray(ey).tobytes()).hexdigest(),
            'sha256_tx': hashlib.sha256(np.ascontiguousarray(pre['FIELDX']).tobytes()
                                        + np.ascontiguousarray(pre['FIELDY']).tobytes()).hexdigest()}
    data = {'out_FIELDX': post['FIELDX'], 'out_FIELDY': post['FIELDY'], 'out_DELAY': post['DELAY'], 'out_DISP': post['DISP'],
            'tx_power_sum': np.array([np.sum(np.abs(pre['FIELDX']) ** 2 + np.abs(pre['FIELDY']) ** 2)]),
            'meta': np.array(json.dumps(meta))}
    for k in ('db0', 'theta', 'epsilon'):
        data['brf_' + k] = np.asarray(brf[k]).ravel()
    data['brf_lcorr'] = np.asarray(brf['lcorr']).ravel()
    os.makedirs(os.path.join(OUT, 'big'), exist_ok=True)
    np.savez_compressed(os.path.join(OUT, 'big', name + '.npz'), **data)
    print('%-28s N=%-6d flag=gps-  |out|=%.6e' % (name, nsymb * nt, np.linalg.norm(post['FIELDX'])))


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[2] == 'c1':     # only the full-size C1 case
        run_c1_case()
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[2] == 'rx':        # only the receiver front-end
        rx_cases()
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[2] == 'invpmd':    # only inverse_pmd.m with its options
        inverse_pmd_cases()
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[2] == 'fixpower':  # only ampliflat(x,'fixpower',...) (ampliflat.m:65-72 + avg_power.m)
        fixpower_cases()
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[2] == 'small':  # only the small-field cases
        run_case('small_ex06_gsx_2e10', 32, 32, 1, 'unique', fib(length=1e5, dphimax=3e-3), 'g-sx', two_pol=False,
                 want_brf=False, rate=10.0, pavg=4.0)
        run_case('small_ex10_sep5_gsx_2e11', 64, 32, 5, 'sepfields', fib(length=1e5, dphimax=3e-3), 'g-sx', two_pol=False,
                 want_brf=False, rate=10.0, pavg=4.0)
        run_case('small_2pol_cnlse_2e9', 32, 16, 1, 'unique', fib(length=5e4, dgd=0.5, nplates=10, manakov='no'), 'gps-')
        run_case('small_sep3_manakov_2e8', 16, 16, 3, 'sepfields', fib(length=4e4, dgd=0.3, nplates=8, manakov='yes',
                                                                        slope=0.057), 'gps-')
        sys.exit(0)
    run_case('lin_gvd_2pol', 256, 16, 1, 'unique', fib(length=1e5), 'g---', want_brf=True)
    run_case('lin_pmd_20plates', 256, 16, 1, 'unique', fib(length=8e4, dgd=0.5, nplates=20), 'gp--')
    run_case('cnlse_10plates_100km', 256, 16, 1, 'unique', fib(length=1e5, dgd=1.0, nplates=10, manakov='no'), 'gps-',
             amp={'gain': 20.0, 'f': 5.0})
    run_case('manakov_100plates_80km', 256, 16, 1, 'unique', fib(length=8e4, dgd=0.1, nplates=100, manakov='yes'), 'gps-',
             amp={'gain': 16.0, 'f': 5.0})
    run_case('cnlse_nopmd', 256, 16, 1, 'unique', fib(length=5e4), 'g-s-')
    run_case('sep3_manakov', 256, 16, 3, 'sepfields', fib(length=5e4, dgd=0.3, nplates=20, manakov='yes', slope=0.057), 'gps-')
    run_case('wdm3_unique_manakov', 128, 64, 3, 'unique', fib(length=4e4, dgd=0.2, nplates=10, manakov='yes'), 'gps-', pavg=1.0)
    run_case('pmf_single', 256, 16, 1, 'unique', fib(length=3e4, dgd=0.7, db0=[1.1, -0.4, 2.0], theta=[0.3, -0.9, 1.2],
                                                     epsilon=[0.1, 0.5, -0.3]), 'gps-')
    run_case('scalar_gs', 256, 16, 1, 'unique', fib(length=5e4), 'g-s-', two_pol=False, want_brf=False)
    run_case('scalar_sep3_gsx', 256, 16, 3, 'sepfields', fib(length=3e4, slope=0.057), 'g-sx', two_pol=False, want_brf=False)
    run_case('scalar_sep3_xonly', 256, 16, 3, 'sepfields', fib(length=2e4, slope=0.057), 'g--x', two_pol=False, want_brf=False)
    run_case('scalar_spm_exact', 256, 16, 1, 'unique', fib(length=5e4), '--s-', two_pol=False, want_brf=False)
    # local-error adaptive step (x.ltol -> scalar_a_ssfm / adaptssfm, fiber.m:639-679,938-1010): 24 accepted and 4
    # rejected steps in the first case
    run_case('scalar_ltol_gs', 256, 16, 1, 'unique', fib(length=3e4, ltol=1e-6), 'g-s-', two_pol=False, want_brf=False)
    run_case('scalar_ltol_sep3_gsx', 256, 16, 3, 'sepfields', fib(length=2e4, ltol=2e-6, slope=0.057), 'g-sx', two_pol=False,
             want_brf=False)
    # x.dphiadapt (scalar_ssfm with tolflag 1, fiber.m:588-611): first step by the local-error method, dphimax
    # recalibrated from it, the rest by the phase criterion (52 steps in the first case)
    run_case('scalar_dphiadapt_gs', 256, 16, 1, 'unique', fib(length=5e4, ltol=1e-6, dphiadapt=True), 'g-s-', two_pol=False,
             want_brf=False)
    run_case('scalar_dphiadapt_sep3_gsx', 256, 16, 3, 'sepfields', fib(length=3e4, ltol=2e-6, dphiadapt=True, slope=0.057),
             'g-sx', two_pol=False, want_brf=False)
    # the sizes of the reference's own scalar-path scripts: ex06_ber.m:14-16 (32 x 32 = 2^10 samples, one channel) and
    # ex10_wdm.m:22-24 (64 x 32 = 2^11 samples, five separate channels with XPM), transmission fiber of ex10_wdm.m:44-52
    run_case('small_ex06_gsx_2e10', 32, 32, 1, 'unique', fib(length=1e5, dphimax=3e-3), 'g-sx', two_pol=False, want_brf=False,
             rate=10.0, pavg=4.0)
    run_case('small_ex10_sep5_gsx_2e11', 64, 32, 5, 'sepfields', fib(length=1e5, dphimax=3e-3), 'g-sx', two_pol=False,
             want_brf=False, rate=10.0, pavg=4.0)
    run_case('small_2pol_cnlse_2e9', 32, 16, 1, 'unique', fib(length=5e4, dgd=0.5, nplates=10, manakov='no'), 'gps-')
    run_case('small_sep3_manakov_2e8', 16, 16, 3, 'sepfields', fib(length=4e4, dgd=0.3, nplates=8, manakov='yes', slope=0.057),
             'gps-')
    # the summary block of fiber.m:392-456 (GSTATE.PRINT)
    run_print_case('manakov', 256, 16, 1, 'unique', fib(length=8e4, dgd=0.1, nplates=100, manakov='yes'), 'gps-')
    run_print_case('wdm3_pmf', 128, 64, 3, 'unique', fib(length=3e4, dgd=0.7, db0=[1.1, -0.4, 2.0], theta=[0.3, -0.9, 1.2],
                                                         epsilon=[0.1, 0.5, -0.3], manakov='no', slope=0.057), 'gps-', pavg=1.0)
    run_print_case('scalar_sep3_ltol', 256, 16, 3, 'sepfields', fib(length=2e4, ltol=2e-6, slope=0.057), 'g-sx', two_pol=False)
    fixpower_cases()
    inverse_pmd_cases()
    rx_cases()
    run_c1_case()
