"""Host logic of the throughput-regime kernel's launch geometry (fb_ru_config through hg_fb_debug_ru_config): no GPU.
The ordered mixdown chains a stream's CTAs, so they must all be resident at once (one wave); the term-tile kernels
give every S warp its own hand-over barriers, which bounds the S warps from below (tile slots) and above (registers)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from huygens_b200 import _lib as L

SMS = 148
SMEM_SM = 227 * 1024
REGS_SM = 65536


def geometry(dtype, n, streams=1, arith=0, bands=0, followers=0, sms=SMS):
    out = (C.c_longlong * 5)()
    L.check(L.lib().hg_fb_debug_ru_config(dtype, arith, n, streams, sms, bands, followers, C.cast(out, C.c_void_p)))
    return dict(W=out[0], NS=out[1], cps=out[2], threads=out[3], smem=out[4])


def ctas_per_sm(g):
    return min(SMEM_SM // (g["smem"] + 1024), REGS_SM // (g["threads"] * 80))


@pytest.mark.parametrize("followers", [0, 2])
def test_one_wave_for_every_float_bank_up_to_65536_filters(followers):
    for n in list(range(18945, 66000, 997)) + [32768, 65536]:
        g = geometry(L.HG_F32, n, followers=followers)
        assert 4 <= g["W"] <= 14 and 2 <= g["NS"] <= 8
        assert g["cps"] * g["W"] * 32 >= n                         # every filter has a lane
        per_sm = ctas_per_sm(g)
        assert per_sm >= 1
        assert g["cps"] <= per_sm * SMS, (n, g)                    # the whole chain is resident at once
        assert g["threads"] == (g["W"] + g["NS"] + 1) * 32 and g["threads"] <= 768


def test_batch_of_streams_packs_two_ctas_per_sm():
    # BASELINE configs[3]: 1024 streams x 1024 filters — 4 CTAs per stream, two of them per SM
    for followers in (0, 2):
        g = geometry(L.HG_F32, 1024, streams=1024, followers=followers)
        assert g["cps"] * g["W"] == 32 and ctas_per_sm(g) == 2 and g["NS"] >= 2


def test_y_tile_kernels_keep_the_s_chain():
    # fused arithmetic, per-band rows, ring followers: ceil(W / 2) S warps
    for kw in (dict(arith=1), dict(bands=1), dict(followers=1)):
        g = geometry(L.HG_F32, 65536, **kw)
        assert g["NS"] == (g["W"] + 1) // 2
        assert ctas_per_sm(g) >= 1


def test_double_banks_and_bad_arguments():
    g = geometry(L.HG_F64, 30000)
    assert g["cps"] <= ctas_per_sm(g) * SMS
    with pytest.raises(L.HuygensError):
        geometry(L.HG_F32, 0)
    with pytest.raises(L.HuygensError):
        geometry(7, 100)
