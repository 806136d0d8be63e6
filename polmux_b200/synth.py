"""Seeded synthetic transmit fields for tests and benchmarks.

PDM-QPSK with NRZ raised-cosine-edge pulses built the way the reference's
electricsource does it (pulse of 2*NT samples overlapped symbol by symbol,
electricsource.m:219-231; 'cosroll' shape :256-269).  Symbols come from
numpy.random.Generator(PCG64(seed)) so that every consumer (oracle, CUDA path,
CPU baseline) sees the same bytes.
"""
from __future__ import annotations

import numpy as np


def cosroll_pulse(nt: int, roll: float = 0.2, duty: float = 1.0) -> np.ndarray:
    """electricsource.m:256-269 -> [2*nt] real pulse."""
    p = np.zeros(2 * nt)
    nl = int(round(0.5 * (1 - roll) * duty * nt))
    nr = int(duty * nt - nl - 1)
    p[nt:nt + nl] = 1.0
    hperiod = duty * nt - 2 * nl
    if hperiod != 0:
        ncos = np.arange(nl, nr + 1)
        p[ncos + nt] = 0.5 * (1 + np.cos(np.pi / hperiod * (ncos - nl + 0.5)))
    p[:nt] = p[nt:][::-1]
    return p


def qpsk_waveform(nsymb: int, nt: int, seed: int, roll: float = 0.2):
    """-> (waveform [nsymb*nt] complex, symbol indices [nsymb] in 0..3)."""
    g = np.random.Generator(np.random.PCG64(seed))
    idx = g.integers(0, 4, size=nsymb)
    sym = ((2 * (idx & 1) - 1) + 1j * (2 * (idx >> 1) - 1)) / np.sqrt(2.0)
    p = cosroll_pulse(nt, roll)
    # block m = sym[m+1]*pulse[:nt] + sym[m]*pulse[nt:]  (cyclic)
    blocks = np.roll(sym, -1)[:, None] * p[None, :nt] + sym[:, None] * p[None, nt:]
    return blocks.reshape(-1), idx


def pdm_qpsk(nsymb: int, nt: int, nch: int = 1, seed_x: int = 1, seed_y: int = 2, roll: float = 0.2):
    """-> (Ex, Ey) [nsymb*nt, nch] complex128 unit-scale fields, (symx, symy) [nsymb, nch]."""
    n = nsymb * nt
    ex = np.empty((n, nch), dtype=np.complex128)
    ey = np.empty((n, nch), dtype=np.complex128)
    sx = np.empty((nsymb, nch), dtype=np.int64)
    sy = np.empty((nsymb, nch), dtype=np.int64)
    for c in range(nch):
        ex[:, c], sx[:, c] = qpsk_waveform(nsymb, nt, seed_x + 100 * c, roll)
        ey[:, c], sy[:, c] = qpsk_waveform(nsymb, nt, seed_y + 100 * c, roll)
    return ex, ey, sx, sy


def wdm_lambdas(nch: int, lam: float = 1550.0, spac: float = 0.4) -> np.ndarray:
    """lasersource.m:163: lamt = lam + spac*(ch - (nch+1)/2)."""
    return lam + spac * (np.arange(1, nch + 1) - (nch + 1) / 2.0)


SMF = dict(**{'lambda': 1550.0}, alphadB=0.2, aeff=80.0, n2=2.7e-20, disp=17.0, slope=0.0, dzmax=2e4, dphimax=5e-3)
