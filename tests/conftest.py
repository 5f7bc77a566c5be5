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
