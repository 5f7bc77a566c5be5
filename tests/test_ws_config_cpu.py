"""Host logic of the warp-specialised kernel's launch geometry (k1_config + the Cfg table of filterbank_ws.inc, through
hg_fb_debug_ws_config): no GPU.  The ordered mixdown chains a stream's CTAs in arrival order, so a launch only works if
every CTA is resident at once: CTAs <= SMs << level, and 1 << level CTAs of that shared-memory size fit one SM."""
import ctypes as C
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from huygens_b200 import _lib as L

SMS = 148
SMEM_SM = 228 * 1024            # per SM; every resident CTA also reserves 1 KB
SMEM_CTA = 227 * 1024           # largest dynamic allocation of one CTA


def geometry(dtype, n, streams=1, nf=0, sms=SMS):
    out = (C.c_longlong * 8)()
    L.check(L.lib().hg_fb_debug_ws_config(dtype, n, streams, sms, nf, C.cast(out, C.c_void_p)))
    return dict(fits=out[0], level=out[1], nf=out[2], cps=out[3], ctas=out[4], tile=out[5], y_slots=out[6], smem=out[7])


def resident(g, sms=SMS):
    per_sm = 1 << g["level"]
    return g["ctas"] <= sms * per_sm and g["smem"] <= SMEM_CTA and (g["smem"] + 1024) * per_sm <= SMEM_SM


def test_baseline_banks_float():
    g = geometry(L.HG_F32, 4096)                     # BASELINE configs[1], the headline
    assert (g["fits"], g["level"], g["nf"], g["ctas"], g["tile"], g["y_slots"]) == (1, 0, 1, 128, 256, 3) and resident(g)
    g = geometry(L.HG_F32, 8192)                     # configs[2] over 8 GPUs
    assert (g["fits"], g["level"], g["nf"], g["ctas"], g["tile"], g["y_slots"]) == (1, 1, 1, 256, 128, 3) and resident(g)
    g = geometry(L.HG_F32, 16384)                    # ... over 4 GPUs
    assert (g["fits"], g["level"], g["nf"], g["ctas"], g["tile"], g["y_slots"]) == (1, 2, 1, 512, 64, 2) and resident(g)
    assert geometry(L.HG_F32, 32768)["fits"] == 0    # ... over 2 GPUs, and one GPU: the throughput-regime kernel
    assert geometry(L.HG_F32, 65536)["fits"] == 0


def test_level_boundaries():
    for n, level in [(1, 0), (SMS * 32, 0), (SMS * 32 + 1, 1), (SMS * 64, 1), (SMS * 64 + 1, 2), (SMS * 128, 2)]:
        g = geometry(L.HG_F32, n)
        assert g["fits"] == 1 and g["level"] == level and g["cps"] == -(-n // 32) and resident(g), (n, g)
    assert geometry(L.HG_F32, SMS * 128 + 1)["fits"] == 0
    for n, level in [(64, 0), (SMS * 32, 0), (SMS * 32 + 1, 1), (SMS * 64, 1)]:       # double: one or two CTAs per SM
        g = geometry(L.HG_F64, n)
        assert g["fits"] == 1 and g["level"] == level and resident(g), (n, g)
    assert geometry(L.HG_F64, SMS * 64 + 1)["fits"] == 0


def test_every_fitting_bank_is_one_wave():
    for dtype in (L.HG_F32, L.HG_F64):
        for nf in (0, 2):
            for n in list(range(1, 19500, 211)) + [4736, 4737, 9472, 9473, 18944, 18945]:
                g = geometry(dtype, n, nf=nf)
                if g["fits"]:
                    assert resident(g), (dtype, nf, n, g)
                    assert g["cps"] * 32 * g["nf"] >= n                     # every filter has a lane
                    assert g["tile"] % 32 == 0 and g["y_slots"] in (2, 3)


def test_streams_multiply_the_ctas():
    g = geometry(L.HG_F32, 64, streams=100)          # 2 CTAs per stream x 100 streams
    assert (g["fits"], g["level"], g["cps"], g["ctas"]) == (1, 1, 2, 200) and resident(g)
    assert geometry(L.HG_F32, 1024, streams=1024)["fits"] == 0           # BASELINE configs[3]: K1-ru
    g = geometry(L.HG_F32, 4096, sms=108)            # a GPU with fewer SMs: the same bank needs two CTAs per SM
    assert g["level"] == 1 and resident(g, sms=108)


def test_two_filters_per_lane_opt_in():
    g = geometry(L.HG_F32, 8192, nf=2)               # CTAs of 64 filters, one per SM
    assert (g["fits"], g["level"], g["nf"], g["ctas"], g["tile"]) == (1, 0, 2, 128, 128) and resident(g)
    g = geometry(L.HG_F32, 16384, nf=2)
    assert (g["fits"], g["level"], g["nf"], g["ctas"], g["tile"]) == (1, 1, 2, 256, 64) and resident(g)
    assert geometry(L.HG_F32, 18945, nf=2)["fits"] == 0
    g = geometry(L.HG_F64, 6000, nf=2)               # float only: double keeps one filter per lane
    assert g["nf"] == 1 and g["level"] == 1
    with pytest.raises(L.HuygensError):
        geometry(L.HG_F32, 0)
