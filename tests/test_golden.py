"""Golden vectors produced by the reference's own headers (tests/golden/make_golden.py, oracle/_ref):
the oracle (CPU) and the CUDA path (GPU) must both reproduce them — per-band outputs, phases and RMS
values bit for bit, the bank sum within the stated tolerance for the GPU (summation order)."""
import glob
import os

import numpy as np
import pytest

import oracle
from conftest import rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TAGS = [("f32", np.float32), ("f64", np.float64)]


def same(a, b):
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def _run_fb(cls, g, dtype, **kw):
    N = int(g["N"])
    fb = cls(2, N, float(g["k_p"]), float(g["k_g"]), dtype=dtype, fixed=bool(g["fixed"]), **kw)
    for n in range(N):
        fb.coefficients(n, g["fwd"][n], g["back"][n])
    fb.mix(g["gains"])
    fb.boost(np.ones(N))
    k = int(g["split"])
    s1, b1 = fb.process(g["x"][:k], bands=True)
    fb.boost(3, 0.0)
    fb.mix(3, 0.0)
    s2, b2 = fb.process(g["x"][k:], bands=True)
    return np.concatenate([s1, s2]), np.concatenate([b1, b2], axis=1)


def test_golden_files_present():
    assert len(glob.glob(os.path.join(GOLD, "*.npz"))) == 13


@pytest.mark.parametrize("tag,dtype", TAGS)
@pytest.mark.parametrize("kind", ["dyn", "fixed"])
def test_oracle_reproduces_golden_filterbank(oracle_lib, kind, tag, dtype):
    g = np.load(os.path.join(GOLD, f"filterbank_{kind}_{tag}.npz"))
    s, b = _run_fb(oracle.Filterbank, g, dtype, mode="strict")
    assert same(b, g["out_bands"]) and same(s, g["out_sum"])


@pytest.mark.parametrize("tag,dtype", TAGS)
def test_oracle_reproduces_golden_osc_rms(oracle_lib, tag, dtype):
    g = np.load(os.path.join(GOLD, f"oscbank_{tag}.npz"))
    osc = oracle.Oscbank(int(g["N"]), dtype=dtype)
    for i, f in enumerate(g["freqs"]):
        osc.freqmod(i, f)
    osc.open()
    osc.deactivate(g["inactive"])
    mix = osc.process(len(g["mix"]))
    assert same(mix.view(dtype), g["mix"].view(dtype)) and same(osc.phases().view(dtype), g["phases_end"].view(dtype))
    g = np.load(os.path.join(GOLD, f"rmsbank_{tag}.npz"))
    out = oracle.RMSbank(int(g["N"]), int(g["width"]), dtype=dtype).process(g["zre"], g["zim"])
    assert same(out, g["out"])


@pytest.mark.parametrize("tag,dtype", TAGS)
def test_oracle_reproduces_golden_latchbank(oracle_lib, tag, dtype):
    g = np.load(os.path.join(GOLD, f"latchbank_{tag}.npz"))
    lt = oracle.Latchbank(int(g["N"]), float(g["thresh"]), float(g["ratio"]), width=int(g["width"]), dtype=dtype)
    ore, oim, lv = lt.process(g["zre"], g["zim"], level=True)
    assert same(ore, g["out_re"]) and same(oim, g["out_im"]) and same(lv, g["level"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag,dtype", TAGS)
def test_cuda_reproduces_golden_latchbank(tag, dtype):
    import huygens_b200 as hb
    g = np.load(os.path.join(GOLD, f"latchbank_{tag}.npz"))
    lt = hb.Latchbank(int(g["N"]), float(g["thresh"]), float(g["ratio"]), width=int(g["width"]), dtype=dtype)
    ore, oim, lv = lt.process(g["zre"], g["zim"], level=True)
    assert same(ore, g["out_re"]) and same(oim, g["out_im"]) and same(lv, g["level"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag,dtype", TAGS)
@pytest.mark.parametrize("kind", ["dyn", "fixed"])
def test_cuda_reproduces_golden_filterbank(kind, tag, dtype):
    import huygens_b200 as hb
    g = np.load(os.path.join(GOLD, f"filterbank_{kind}_{tag}.npz"))
    s, b = _run_fb(hb.Filterbank, g, dtype, arith="strict")
    assert same(b, g["out_bands"])
    assert rel_err(s, g["out_sum"]) <= (1e-5 if dtype == np.float32 else 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,dtype", TAGS)
def test_cuda_reproduces_golden_osc_rms(tag, dtype):
    import huygens_b200 as hb
    g = np.load(os.path.join(GOLD, f"oscbank_{tag}.npz"))
    osc = hb.Oscbank(int(g["N"]), dtype=dtype)
    for i, f in enumerate(g["freqs"]):
        osc.freqmod(i, f)
    osc.open()
    osc.deactivate(g["inactive"])
    mix = osc.process(len(g["mix"]))
    assert same(osc.phases().view(dtype), g["phases_end"].view(dtype))
    assert rel_err(mix.view(dtype), g["mix"].view(dtype)) <= (1e-5 if dtype == np.float32 else 1e-12)
    g = np.load(os.path.join(GOLD, f"rmsbank_{tag}.npz"))
    out = hb.RMSbank(int(g["N"]), int(g["width"]), dtype=dtype).process(g["zre"], g["zim"])
    assert same(out, g["out"])


def _run_slide(cls, g, dtype, **kw):
    bank = cls(int(g["N"]), int(g["order"]), g["radii"], dtype=dtype, **kw)
    sp = int(g["split"])
    return np.concatenate([bank.process(g["z"][:, :sp]), bank.process(g["z"][:, sp:])], axis=1)


def _run_stick(cls, g, order, rad, **kw):
    bank = cls(int(g["N"]), order, rad, dtype=np.float64, **kw)
    sp = int(g["split"])
    a = bank.process(g["z"][:, :sp])
    bank.rmod(rad * 0.5)
    return np.concatenate([a, bank.process(g["z"][:, sp:])], axis=1)


def _same_c(a, b):
    return same(np.ascontiguousarray(a.real), np.ascontiguousarray(b.real)) and same(np.ascontiguousarray(a.imag), np.ascontiguousarray(b.imag))


@pytest.mark.parametrize("tag,dtype", TAGS)
def test_oracle_reproduces_golden_slidebank(oracle_lib, tag, dtype):
    g = np.load(os.path.join(GOLD, f"slidebank_{tag}.npz"))
    assert _same_c(_run_slide(oracle.Slidebank, g, dtype), g["out"])


def test_oracle_reproduces_golden_stickbank(oracle_lib):
    g = np.load(os.path.join(GOLD, "stickbank_f64.npz"))
    for order, rad in g["cases"]:
        assert _same_c(_run_stick(oracle.Stickbank, g, int(order), float(rad)), g[f"out_{int(order)}_{float(rad)}"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag,dtype", TAGS)
def test_cuda_reproduces_golden_slidebank(tag, dtype):
    import huygens_b200 as hb
    g = np.load(os.path.join(GOLD, f"slidebank_{tag}.npz"))
    assert _same_c(_run_slide(hb.Slidebank, g, dtype), g["out"])


@pytest.mark.gpu
def test_cuda_reproduces_golden_stickbank():
    import huygens_b200 as hb
    g = np.load(os.path.join(GOLD, "stickbank_f64.npz"))
    for order, rad in g["cases"]:
        assert _same_c(_run_stick(hb.Stickbank, g, int(order), float(rad)), g[f"out_{int(order)}_{float(rad)}"])


def test_golden_vectors_regenerate_identically(oracle_lib):
    """The committed fixtures are what tests/golden/make_golden.py produces from oracle/_ref today (skipped where
    the reference build is not available, e.g. on the GPU box without /root/reference and without a prebuilt _ref)."""
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref not built")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    cases = {"filterbank_dyn_f64.npz": lambda: mg.filterbank_case(np.float64, False),
             "filterbank_fixed_f32.npz": lambda: mg.filterbank_case(np.float32, True),
             "oscbank_f32.npz": lambda: mg.oscbank_case(np.float32),
             "rmsbank_f64.npz": lambda: mg.rmsbank_case(np.float64),
             "latchbank_f32.npz": lambda: mg.latchbank_case(np.float32),
             "slidebank_f64.npz": lambda: mg.slidebank_case(np.float64),
             "stickbank_f64.npz": mg.stickbank_case}
    for name, make in cases.items():
        want, got = np.load(os.path.join(GOLD, name)), make()
        assert set(want.files) == set(got.keys()), name
        for k in want.files:
            a, b = np.asarray(want[k]), np.asarray(got[k])
            assert a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes(), (name, k)
