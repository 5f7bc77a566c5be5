"""Known-answer tests that pin the oracle (the reference ships no golden vectors).

(i)   biquad impulse response closed form
(ii)  unity steady-state gain at the centre frequency after the 1/|H| normalisation
(iii) Oscbank |phase| -> 1 and arg = n * 2 pi f / SR
(iv)  RMSbank of a unit-modulus input: sqrt(n/width) ramp, then 1
(v)   gain smoothing p[n] = pin (1 - s^(n+1)) from zero
plus the reference quirks: zero-initialised gains render silence (resynthesis.cpp never calls
boost/mix), coefficient truncation, FFilterbank constants, the RMSbank double-call sqrt.
"""
import numpy as np
import pytest

import oracle
from huygens_b200 import design

MODES = ["strict", "fma"]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_impulse_response_closed_form(oracle_lib, mode, dtype):
    r, f = 0.99, 1000.0
    th = 2 * design.PI * f / design.SR
    fb = oracle.Filterbank(2, 1, 0, 0, dtype=dtype, mode=mode)   # k=0 -> no smoothing: p = pin at once
    g = 0.5
    fb.coefficients(0, [g, 0, -g], [-2 * r * np.cos(th), r * r])
    fb.boost(0, 1.0)
    fb.mix(0, 1.0)
    x = np.zeros(400)
    x[0] = 1.0
    y = fb.process(x)
    n = np.arange(400, dtype=np.float64)
    h0 = g * r ** n * np.sin((n + 1) * th) / np.sin(th)
    h = h0.copy()
    h[2:] -= g * r ** (n[2:] - 2) * np.sin((n[2:] - 1) * th) / np.sin(th)
    tol = 2e-5 if dtype == np.float32 else 1e-12
    assert np.max(np.abs(y - h)) < tol


@pytest.mark.parametrize("mode", MODES)
def test_unity_gain_at_centre(oracle_lib, mode):
    N = 8
    freqs, fwd, back = design.resynthesis_bank(N)
    fb = oracle.Filterbank(2, N, 0, 0, dtype=np.float64, mode=mode)
    for n in range(N):
        fb.coefficients(n, fwd[n], back[n])
    fb.boost(np.ones(N))
    k = 3
    fb.mix(k, 1.0)          # listen to one band
    L = 60000
    t = np.arange(L)
    x = np.sin(2 * design.PI * freqs[k] * t / design.SR)
    y = fb.process(x)
    amp = np.sqrt(2 * np.mean(y[-9000:] ** 2))      # steady-state amplitude of the sampled sine
    assert abs(amp - 1.0) < 2e-3


def test_silence_without_boost_or_mix(oracle_lib):
    # tests/resynthesis.cpp:48-54 never calls boost()/mix()/open(): gains stay 0 (filterbank.h:65-69)
    fb = oracle.Filterbank(2, 4, dtype=np.float64)
    for n in range(4):
        fb.coefficients(n, [1, 0, -1], [-1.9, 0.95])
    assert np.all(fb.process(design.white_noise(256)) == 0)
    fb.open()               # gains_in = 1 but preamps still 0 -> still silent
    assert np.all(fb.process(design.white_noise(256)) == 0)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_smoothing_closed_form(oracle_lib, mode, dtype):
    k_p, k_g = 0.01, 0.02
    fb = oracle.Filterbank(2, 1, k_p, k_g, dtype=dtype, mode=mode)
    fb.boost(0, 0.7)
    fb.mix(0, 1.3)
    n = 500
    fb.process(np.zeros(n))
    p, g = fb.smoothed()
    sp, sg = design.relaxation(k_p), design.relaxation(k_g)
    tol = 1e-4 if dtype == np.float32 else 1e-12
    assert abs(p[0] - 0.7 * (1 - sp ** n)) < tol
    assert abs(g[0] - 1.3 * (1 - sg ** n)) < tol


def test_relaxation_values():
    # SURVEY.md §8 a1
    assert design.relaxation(0) == 0
    assert abs(design.relaxation(0.1) - 0.99251902835739503) < 1e-15
    assert abs(design.relaxation(1) - 0.99924937241633582) < 1e-15
    assert abs(design.relaxation(0.001) - 0.47193715634084676) < 1e-15
    assert abs(design.relaxation(50) - 0.99998498192386032) < 1e-15


def test_coefficient_truncation_and_order(oracle_lib):
    # filterbank.h:75-81: extra entries ignored, missing entries keep their value
    a = oracle.Filterbank(2, 1, 0, 0)
    b = oracle.Filterbank(2, 1, 0, 0)
    a.coefficients(0, [0.3, 0.2, 0.1, 99.0], [-0.5, 0.25, 99.0])
    b.coefficients(0, [0.3, 0.2, 0.1], [-0.5, 0.25])
    for f in (a, b):
        f.boost(0, 1)
        f.mix(0, 1)
    x = design.white_noise(300)
    assert np.array_equal(a.process(x), b.process(x))
    c = oracle.Filterbank(2, 1, 0, 0)
    c.coefficients(0, [0.3, 0.2, 0.1], [-0.5, 0.25])
    c.coefficients(0, [0.4], [-0.6])            # only b0 and a1 replaced
    d = oracle.Filterbank(2, 1, 0, 0)
    d.coefficients(0, [0.4, 0.2, 0.1], [-0.6, 0.25])
    for f in (c, d):
        f.boost(0, 1)
        f.mix(0, 1)
    assert np.array_equal(c.process(x), d.process(x))


def test_ffilterbank_constants_differ_only_in_float(oracle_lib):
    x = design.white_noise(2000)
    for dtype in (np.float32, np.float64):
        outs = []
        for fixed in (False, True):
            fb = oracle.Filterbank(2, 1, 0.001, 1, dtype=dtype, fixed=fixed)
            fb.coefficients(0, [1, 0, -1], [-1.8, 0.9])
            fb.boost(0, 1)
            fb.mix(0, 1)
            outs.append(fb.process(x))
        if dtype == np.float64:
            assert np.array_equal(outs[0], outs[1])
        else:
            assert np.allclose(outs[0], outs[1], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_oscbank_unit_modulus_and_angle(oracle_lib, mode, dtype):
    N = 4
    osc = oracle.Oscbank(N, dtype=dtype, mode=mode)
    freqs = [100.0, 440.0, 1000.0, 5000.0]
    for i, f in enumerate(freqs):
        osc.freqmod(i, f)
    osc.freqmod(7, 1.0)         # out of range: silently ignored (oscbank.h:51)
    # multichannel.h:19-24: nothing is active until open()
    osc.process(10)
    assert np.all(osc.phases() == 1)
    osc.open()
    assert osc.active_count() == N
    n = 1000
    mix, ph = osc.process(n, phases=True)
    p = osc.phases()
    tol = 3e-3 if dtype == np.float32 else 1e-9
    assert np.max(np.abs(np.abs(p) - 1)) < (1e-6 if dtype == np.float32 else 1e-14)
    for i, f in enumerate(freqs):
        expect = np.exp(1j * n * 2 * design.PI * f / design.SR)
        assert abs(p[i] - expect) < tol
    assert np.allclose(mix, ph.sum(axis=1), atol=1e-5 if dtype == np.float32 else 1e-13)
    osc.deactivate([1, 2])
    assert osc.active_count() == 2
    before = osc.phases()
    osc.process(5)
    after = osc.phases()
    assert after[1] == before[1] and after[2] == before[2] and after[0] != before[0]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_rmsbank_ramp_and_plateau(oracle_lib, mode, dtype):
    N, w, L = 3, 100, 350
    rms = oracle.RMSbank(N, w, dtype=dtype, mode=mode)
    ang = np.linspace(0, 50, L)
    zre = np.tile(np.cos(ang), (N, 1))
    zim = np.tile(np.sin(ang), (N, 1))
    out = rms.process(zre, zim)
    n = np.arange(1, L + 1)
    expect = np.sqrt(np.minimum(n, w) / w)
    tol = 1e-5 if dtype == np.float32 else 1e-12
    assert np.max(np.abs(out - expect[None, :])) < tol


def test_rmsbank_double_call_quirk(oracle_lib):
    # rmsbank.h:64: every call within a tick applies sqrt(out/width) again (tests/phasebank.cpp:87-88)
    N, w, L = 2, 50, 120
    z = np.abs(design.white_noise(N * L).reshape(N, L)) + 0.1
    once = oracle.RMSbank(N, w).process(z)
    twice = oracle.RMSbank(N, w).process(z, calls_per_tick=2)
    assert np.allclose(twice, np.sqrt(once / w), rtol=1e-14)


def test_strict_and_fma_modes_diverge_in_float(oracle_lib):
    # SURVEY.md §7 hard part 1: any differently rounded op decorrelates the fp32 trajectories
    from conftest import make_resonator, rel_err
    x = design.white_noise(48000, dtype=np.float32)
    a = make_resonator(oracle.Filterbank, 32, np.float32, mode="strict").process(x)
    b = make_resonator(oracle.Filterbank, 32, np.float32, mode="fma").process(x)
    assert 1e-7 < rel_err(a, b) < 1e-2
    c = make_resonator(oracle.Filterbank, 32, np.float64, mode="strict").process(x.astype(np.float64))
    d = make_resonator(oracle.Filterbank, 32, np.float64, mode="fma").process(x.astype(np.float64))
    assert rel_err(c, d) < 1e-12


@pytest.mark.parametrize("mode", ["strict", "fma"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_chained_partitions_equal_single_bank(oracle_lib, dtype, mode):
    """oracle.chained_sum (consecutive filter partitions, each continuing the running row of the one before —
    Filterbank::mix_from) performs the single bank's operations in the single bank's order: bit-identical,
    with and without the per-band distortion.  The full-size GPU tests lean on this to get the oracle's
    own 4096- and 65,536-term sums in seconds."""
    import oracle
    from conftest import make_resonator, oracle_chained
    N, L = 301, 6000
    x = (20.0 * design.white_noise(L)).astype(dtype)
    for dist in (0, 1):
        whole = make_resonator(oracle.Filterbank, N, dtype, mode=mode, k_p=0.01, k_g=0.05).process(x, distortion=dist)
        for parts in (1, 2, 7):
            got = oracle_chained(N, dtype, x, mode=mode, partitions=parts, distortion=dist, k_p=0.01, k_g=0.05)
            assert np.array_equal(got, whole)
