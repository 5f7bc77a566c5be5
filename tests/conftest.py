import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle
    oracle.build()
    return oracle.lib()


def make_resonator(cls, N, dtype, r=0.999, f_lo=27.5, f_hi=20000.0, k_p=0.001, k_g=1.0, **kw):
    """tests/filterbank.cpp:255-271 style bank on `cls` (oracle or CUDA Filterbank)."""
    import numpy as np
    from huygens_b200 import design
    freqs = design.log_spaced(N, f_lo, f_hi)
    fwd, back, gains = design.resonator_bank(freqs, r)
    fb = cls(2, N, k_p, k_g, dtype=dtype, **kw)
    for n in range(N):
        fb.coefficients(n, fwd[n], back[n])
    fb.mix(gains)
    fb.boost(np.ones(N))
    return fb


def rel_err(a, b):
    import numpy as np
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    peak = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (peak if peak > 0 else 1.0))


def snr_db(a, ref):
    import numpy as np
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    num = np.sum(ref * ref)
    den = np.sum((a - ref) ** 2)
    return float("inf") if den == 0 else float(10 * np.log10(num / den))


def oracle_chained(N, dtype, x, mode="strict", partitions=None, distortion=0, **kw):
    """The oracle's bank sum of an N-filter make_resonator bank over x, evaluated as consecutive filter
    partitions on all host threads (oracle.chained_sum: the single bank's operations in the single bank's
    order, bit for bit — tests/test_oracle_kat.py::test_chained_partitions_equal_single_bank)."""
    import numpy as np
    import oracle
    from huygens_b200 import design
    partitions = partitions or (os.cpu_count() or 1)
    r, f_lo, f_hi = kw.pop("r", 0.999), kw.pop("f_lo", 27.5), kw.pop("f_hi", 20000.0)
    k_p, k_g = kw.pop("k_p", 0.001), kw.pop("k_g", 1.0)
    freqs = design.log_spaced(N, f_lo, f_hi)
    fwd, back, gains = design.resonator_bank(freqs, r)
    banks = []
    for idx in np.array_split(np.arange(N), min(partitions, N)):
        fb = oracle.Filterbank(2, len(idx), k_p, k_g, dtype=dtype, mode=mode, **kw)
        for j, n in enumerate(idx):
            fb.coefficients(j, fwd[n], back[n])
        fb.mix(gains[idx])
        fb.boost(np.ones(len(idx)))
        banks.append(fb)
    return oracle.chained_sum(banks, x, distortion=distortion)
