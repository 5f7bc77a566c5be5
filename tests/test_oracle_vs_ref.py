"""Pins oracle/huygens_oracle.hpp against the reference SOURCE: oracle/_ref is the unmodified
/root/reference/src/{filterbank,oscbank,rmsbank,multichannel}.h compiled against oracle/shim (a minimal
stand-in for the Eigen subset they use).  The restatement (strict mode) must agree BIT FOR BIT with it
on seeded inputs, including the internal state (per-band outputs, smoothed gains)."""
import numpy as np
import pytest

import oracle
from huygens_b200 import design

pytestmark = pytest.mark.skipif(oracle.ref_lib() is None, reason="oracle/_ref was not built (no /root/reference)")
DTYPES = [np.float32, np.float64]


def same(a, b):
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def test_relaxation_matches_reference_function():
    for k in (0, 0.001, 0.1, 1, 2.0 / 48000, 50):
        assert oracle.ref_lib().hgr_relaxation(k) == design.relaxation(k)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("N,fixed", [(1, False), (7, False), (64, False), (4, True), (16, True), (96, True)])
def test_filterbank_restatement_equals_reference(oracle_lib, N, fixed, dtype):
    rng = np.random.default_rng(N)
    k_p, k_g = (0.001, 1.0) if fixed else (0.01, 0.05)
    ref = oracle.RefFilterbank(2, N, k_p, k_g, dtype=dtype, fixed=fixed)
    mine = oracle.Filterbank(2, N, k_p, k_g, dtype=dtype, mode="strict", fixed=fixed)
    freqs = design.log_spaced(N, 27.5, 18000.0)
    fwd, back, gains = design.resonator_bank(freqs, 0.999)
    t = 0
    for blk in range(5):
        for f in (ref, mine):
            if blk == 0:
                for n in range(N):
                    # general taps (b1 != 0) on odd filters: exercises the GEMV accumulation order
                    f.coefficients(n, fwd[n] if n % 2 == 0 else [0.5, 0.25, -0.125], back[n])
                f.mix(gains)
                f.boost(np.ones(N))
            elif blk == 2:
                f.boost(0, 0.0)
                f.mix(N - 1, 0.0)
                f.coefficients(0, [0.3], [-0.2, 0.1, 7.0])      # partial / over-long vectors (filterbank.h:75-81)
            elif blk == 3:
                f.open()
        L = int(rng.integers(50, 700))
        x = design.white_noise(L, dtype=dtype, t0=t)
        t += L
        r_sum, r_b = ref.process(x, bands=True)
        m_sum, m_b = mine.process(x, bands=True)
        assert same(m_b, r_b)
        assert same(m_sum, r_sum)
        assert all(same(a, b) for a, b in zip(mine.smoothed(), ref.smoothed()))


@pytest.mark.parametrize("order", [0, 1])
def test_filterbank_low_orders_equal_reference(oracle_lib, order):
    N = 5
    ref = oracle.RefFilterbank(order, N, 0.01, 0.01)
    mine = oracle.Filterbank(order, N, 0.01, 0.01)
    for n in range(N):
        for f in (ref, mine):
            f.coefficients(n, [0.5 + 0.01 * n, 0.25, 0.1], [-0.3 - 0.01 * n, 0.2])
    for f in (ref, mine):
        f.boost(np.ones(N))
        f.open()
    x = design.white_noise(400)
    r_sum, r_b = ref.process(x, bands=True)
    m_sum, m_b = mine.process(x, bands=True)
    assert same(m_b, r_b) and same(m_sum, r_sum)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("N", oracle.REF_OSC_N)
def test_oscbank_restatement_equals_reference(oracle_lib, N, dtype):
    ref, mine = oracle.RefOscbank(N, dtype=dtype), oracle.Oscbank(N, dtype=dtype)
    freqs = design.log_spaced(N, 20.0, 20000.0)
    for o in (ref, mine):
        for i, f in enumerate(freqs):
            o.freqmod(i, f)
        o.freqmod(N + 2, 1.0)
        o.freqmod(-1, 1.0)
    assert same(ref.process(3), mine.process(3))            # nothing active yet (multichannel.h:19-24)
    for o in (ref, mine):
        o.open()
        o.deactivate([0, N - 1, N // 2, 10 ** 6])
        o.activate([N // 2])
    assert ref.active_count() == mine.active_count()
    r_mix, r_ph = ref.process(1500, phases=True)
    m_mix, m_ph = mine.process(1500, phases=True)
    assert same(m_ph.view(dtype), r_ph.view(dtype))
    assert same(m_mix.view(dtype), r_mix.view(dtype))
    for o in (ref, mine):
        o.close()
    assert ref.active_count() == mine.active_count() == 0
    assert same(ref.phases().view(dtype), mine.phases().view(dtype))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("N,width", [(1, 7), (8, 100), (40, 2400)])
@pytest.mark.parametrize("calls", [1, 2])
def test_rmsbank_restatement_equals_reference(oracle_lib, N, width, dtype, calls):
    L = 3000
    zre = np.stack([design.white_noise(L, stream=2 * n, dtype=dtype) for n in range(N)])
    zim = np.stack([design.white_noise(L, stream=2 * n + 1, dtype=dtype) for n in range(N)])
    r = oracle.RefRMSbank(N, width, dtype=dtype).process(zre, zim, calls_per_tick=calls)
    m = oracle.RMSbank(N, width, dtype=dtype).process(zre, zim, calls_per_tick=calls)
    assert same(m, r)


def _latch_signal(N, L, dtype):
    t = np.arange(L)
    env = np.stack([0.5 * (1 + np.sin(2 * np.pi * t / (300.0 + 37 * n) + n)) for n in range(N)])      # levels sweeping 0..1
    zre = (env * np.cos(0.3 * t)[None, :]).astype(dtype)
    zim = (env * np.sin(0.3 * t)[None, :]).astype(dtype)
    return zre, zim


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("N,width", [(1, 7), (8, 50), (40, 240)])
def test_latchbank_restatement_equals_reference(oracle_lib, N, width, dtype):
    """Latchbank hysteresis gate driven by its RMSbank (src/latchbank.h:81-100, tests/harmbank.cpp:84-99)
    and the real-input overload (:61-79)."""
    L = 2500
    zre, zim = _latch_signal(N, L, dtype)
    r = oracle.RefLatchbank(N, 0.35, 0.2, width=width, dtype=dtype).process(zre, zim, level=True)
    m = oracle.Latchbank(N, 0.35, 0.2, width=width, dtype=dtype).process(zre, zim, level=True)
    for a, b in zip(m, r):
        assert same(a, b)
    assert 0.05 < np.mean(r[0] != 0) < 0.95                 # the gate both opens and closes
    lv = np.abs(zre) + 0.01
    assert same(oracle.Latchbank(N, 0.35, 0.3, dtype=dtype).process_real(lv), oracle.RefLatchbank(N, 0.35, 0.3, dtype=dtype).process_real(lv))


def _casc_input(N, L, dtype, seed=3):
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((N, L)) + 1j * rng.standard_normal((N, L))
    z[:, L // 3:L // 3 + 17] = 0.0                            # exact zeros: the alpha*rhs products' zero signs
    z[:, L // 2] = -0.0
    return z.astype(np.complex64 if dtype == np.float32 else np.complex128)


def _same_c(a, b):
    return same(np.ascontiguousarray(a.real), np.ascontiguousarray(b.real)) and same(np.ascontiguousarray(a.imag), np.ascontiguousarray(b.imag))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("order", [0, 1, 2, 5])
@pytest.mark.parametrize("N", oracle.REF_CASC_N)
def test_slidebank_restatement_equals_reference(oracle_lib, N, order, dtype):
    """Slidebank's sparse-matrix cascade (src/slidebank.h:61-97,166-175), complex and real input, block splits."""
    L = 700
    rng = np.random.default_rng(N + order)
    radii = rng.uniform(0.5, 0.9995, N) * np.exp(1j * rng.uniform(-0.2, 0.2, N))
    z = _casc_input(N, L, dtype)
    r = oracle.RefSlidebank(N, order, radii, dtype=dtype)
    m = oracle.Slidebank(N, order, radii, dtype=dtype)
    for a, b in ((0, 1), (1, 300), (300, L)):
        assert _same_c(m.process(z[:, a:b]), r.process(z[:, a:b]))
    assert _same_c(m.process(z.real[:, :50]), r.process(z.real[:, :50]))      # the ArrayT overload (:109-115)


@pytest.mark.parametrize("order,rad", [(1, -0.9), (1, -0.999), (0, -0.5), (2, 0.3), (3, 0.2)])
@pytest.mark.parametrize("N", oracle.REF_CASC_N)
def test_stickbank_restatement_equals_reference(oracle_lib, N, order, rad):
    """Stickbank's all-pole smoother (src/stickbank.h:46-79,181-218) incl. rmod and coefficients(); the
    reference template only compiles for double."""
    L = 600
    z = _casc_input(N, L, np.float64, seed=11)
    r = oracle.RefStickbank(N, order, rad)
    m = oracle.Stickbank(N, order, rad)
    assert _same_c(m.process(z[:, :250]), r.process(z[:, :250]))
    r.rmod(rad * 0.7)
    m.rmod(rad * 0.7)
    assert _same_c(m.process(z[:, 250:]), r.process(z[:, 250:]))
    assert np.isfinite(r.process(z[:, :10])).all()
