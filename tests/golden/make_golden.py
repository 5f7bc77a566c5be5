#!/usr/bin/env python
"""Generates tests/golden/*.npz from oracle/_ref — the unmodified reference headers compiled against
oracle/shim (needs /root/reference; run in the build container, commit the outputs).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from huygens_b200 import design  # noqa: E402


def filterbank_case(dtype, fixed):
    N, L = 16, 700
    k_p, k_g = (0.001, 1.0) if fixed else (0.01, 0.05)
    fb = oracle.RefFilterbank(2, N, k_p, k_g, dtype=dtype, fixed=fixed)
    freqs = design.log_spaced(N, 40.0, 16000.0)
    fwd, back, gains = design.resonator_bank(freqs, 0.999)
    fwd[1::4] = [0.5, 0.25, -0.125]            # a few general (b1 != 0) tap sets
    for n in range(N):
        fb.coefficients(n, fwd[n], back[n])
    fb.mix(gains)
    fb.boost(np.ones(N))
    x = design.white_noise(L, dtype=dtype)
    s1, b1 = fb.process(x[:400], bands=True)
    fb.boost(3, 0.0)                           # a "note off" between blocks
    fb.mix(3, 0.0)
    s2, b2 = fb.process(x[400:], bands=True)
    return dict(N=N, k_p=k_p, k_g=k_g, fixed=fixed, fwd=fwd, back=back, gains=gains, x=x, split=400,
                out_sum=np.concatenate([s1, s2]), out_bands=np.concatenate([b1, b2], axis=1))


def oscbank_case(dtype):
    N, L = 8, 600
    osc = oracle.RefOscbank(N, dtype=dtype)
    freqs = 55.0 * 2.0 ** np.arange(N)
    for i, f in enumerate(freqs):
        osc.freqmod(i, f)
    osc.open()
    osc.deactivate([2])
    mix, ph = osc.process(L, phases=True)
    return dict(N=N, freqs=freqs, inactive=np.array([2]), mix=mix, phases_last=ph[-1], phases_end=osc.phases())


def rmsbank_case(dtype):
    N, L, width = 8, 500, 100
    zre = np.stack([design.white_noise(L, stream=2 * n, dtype=dtype) for n in range(N)])
    zim = np.stack([design.white_noise(L, stream=2 * n + 1, dtype=dtype) for n in range(N)])
    out = oracle.RefRMSbank(N, width, dtype=dtype).process(zre, zim)
    return dict(N=N, width=width, zre=zre, zim=zim, out=out)


def latchbank_case(dtype):
    N, L, width = 8, 1500, 50
    t = np.arange(L)
    env = np.stack([0.5 * (1 + np.sin(2 * np.pi * t / (300.0 + 37 * n) + n)) for n in range(N)])
    zre = (env * np.cos(0.3 * t)[None, :]).astype(dtype)
    zim = (env * np.sin(0.3 * t)[None, :]).astype(dtype)
    ore, oim, lv = oracle.RefLatchbank(N, 0.35, 0.2, width=width, dtype=dtype).process(zre, zim, level=True)
    return dict(N=N, width=width, thresh=0.35, ratio=0.2, zre=zre, zim=zim, out_re=ore, out_im=oim, level=lv)


def cascade_input(N, L, dtype):
    t = np.arange(L)
    z = np.stack([(0.3 + 0.7 * design.white_noise(L, stream=40 + n)) * np.exp(1j * (0.01 * (n + 1) * t)) for n in range(N)])
    z[:, 200:230] = 0.0                            # a stretch of exact zeros (sign-of-zero paths)
    return z.astype(np.complex64 if dtype == np.float32 else np.complex128)


def slidebank_case(dtype):
    N, L, order = 8, 600, 3
    radii = np.array([0.999 * np.exp(1j * 0.002 * n) * (1 - 0.01 * n) for n in range(N)])
    z = cascade_input(N, L, dtype)
    bank = oracle.RefSlidebank(N, order, radii, dtype=dtype)
    out = np.concatenate([bank.process(z[:, :257]), bank.process(z[:, 257:])], axis=1)
    return dict(N=N, order=order, radii=radii, z=z, split=257, out=out)


def stickbank_case():
    N, L = 8, 600
    z = cascade_input(N, L, np.float64)
    res = {}
    for order, rad in ((1, -0.9), (1, -0.999), (2, 0.3)):
        bank = oracle.RefStickbank(N, order, rad)
        a = bank.process(z[:, :300])
        bank.rmod(rad * 0.5)
        res[f"out_{order}_{rad}"] = np.concatenate([a, bank.process(z[:, 300:])], axis=1)
    return dict(N=N, z=z, split=300, cases=np.array([[1, -0.9], [1, -0.999], [2, 0.3]]), **res)


def main():
    assert oracle.ref_lib() is not None, "oracle/_ref is not built"
    for dtype, tag in ((np.float32, "f32"), (np.float64, "f64")):
        np.savez_compressed(os.path.join(HERE, f"filterbank_dyn_{tag}.npz"), **filterbank_case(dtype, False))
        np.savez_compressed(os.path.join(HERE, f"filterbank_fixed_{tag}.npz"), **filterbank_case(dtype, True))
        np.savez_compressed(os.path.join(HERE, f"oscbank_{tag}.npz"), **oscbank_case(dtype))
        np.savez_compressed(os.path.join(HERE, f"rmsbank_{tag}.npz"), **rmsbank_case(dtype))
        np.savez_compressed(os.path.join(HERE, f"latchbank_{tag}.npz"), **latchbank_case(dtype))
        np.savez_compressed(os.path.join(HERE, f"slidebank_{tag}.npz"), **slidebank_case(dtype))
    np.savez_compressed(os.path.join(HERE, "stickbank_f64.npz"), **stickbank_case())
    print("written:", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
