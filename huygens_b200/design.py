"""Host-side designers and synthetic inputs (double precision, numpy).

Restates the recipes the reference's callers use to fill a Filterbank, and defines the
deterministic synthetic inputs of the offline render harness that replaces the PortAudio stream
(/root/reference/src/audio.h) for benchmarking.
"""
import numpy as np

SR = 48000                 # /root/reference/src/includes.h:32
PI = 3.14159265359         # /root/reference/src/includes.h:30 (12 significant digits, on purpose)


def relaxation(k):
    """/root/reference/src/includes.h:38-48."""
    if k == 0:
        return 0.0
    order = np.log2(np.finfo(np.float64).eps)
    return float(2.0 ** (order / (max(0.0, k) * SR)))


def transfer(frequency, r, f):
    """|H| helper of the callers: tests/resynthesis.cpp:23-27, tests/filterbank.cpp:152-156."""
    frequency = np.asarray(frequency, dtype=np.float64)
    f = np.asarray(f, dtype=np.float64)
    z = np.cos(2 * PI * f / SR) - 1j * np.sin(2 * PI * f / SR)
    return (1.0 - z * z) / (1.0 - 2 * r * np.cos(2 * PI * frequency / SR) * z + r * r * z * z)


def resonator_bank(freqs, r):
    """tests/filterbank.cpp:255-271 recipe: forward {1,0,-1}, back {-2 r cos(theta), r^2},
    mix gains 1/|H(f)|.  Returns (fwd[N,3], back[N,2], gains[N]) in float64."""
    freqs = np.asarray(freqs, dtype=np.float64)
    angle = freqs * (2 * PI / SR)
    fwd = np.tile(np.array([1.0, 0.0, -1.0]), (len(freqs), 1))
    back = np.stack([-2 * r * np.cos(angle), np.full_like(freqs, r * r)], axis=1)
    gains = 1.0 / np.abs(transfer(freqs, r, freqs))
    return fwd, back, gains


def resynthesis_bank(N, R=0.999):
    """tests/resynthesis.cpp:48-54 recipe with the gain folded into the forward taps.  The
    reference puts filter N-1 on Nyquist (f = 0.5 (i+1) SR / N); we use N+1 in the denominator."""
    i = np.arange(N, dtype=np.float64)
    freqs = 0.5 * (i + 1) * SR / (N + 1)
    cosine = np.cos(2 * PI * freqs / SR)
    gain = np.abs(transfer(freqs, R, freqs))
    fwd = np.stack([1.0 / gain, np.zeros(N), -1.0 / gain], axis=1)
    back = np.stack([-2 * R * cosine, np.full(N, R * R)], axis=1)
    return freqs, fwd, back


def log_spaced(N, f_lo=27.5, f_hi=20000.0):
    if N == 1:
        return np.array([f_lo])
    return f_lo * (f_hi / f_lo) ** (np.arange(N, dtype=np.float64) / (N - 1))


_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_STREAM = np.uint64(0xD1342543DE82EF95)


def _splitmix64(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def white_noise(L, seed=0x5EED, stream=0, dtype=np.float64, t0=0):
    """Counter-based uniform(-1, 1): x[t] = 2 * (splitmix64(seed_s + (t+1) * GOLDEN) >> 11) * 2^-53 - 1,
    seed_s = seed ^ (stream * 0xD1342543DE82EF95).  Replaces Noise<T> (src/noise.h:25-28, rand())."""
    with np.errstate(over="ignore"):
        s = np.uint64(seed) ^ (np.uint64(stream) * _STREAM)
        t = np.arange(t0 + 1, t0 + L + 1, dtype=np.uint64)
        z = _splitmix64(s + t * _GOLDEN)
    x = 2.0 * ((z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53) - 1.0
    return x.astype(dtype)


def chirp(L, f0=20.0, f1=20000.0, amp=0.5, dtype=np.float64):
    """Exponential sine sweep f0 -> f1 over L samples, computed in double then cast."""
    k = f1 / f0
    T = L / SR
    t = np.arange(L, dtype=np.float64) / L
    x = amp * np.sin(2 * np.pi * f0 * T / np.log(k) * (k ** t - 1.0))
    return x.astype(dtype)
