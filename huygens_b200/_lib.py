"""ctypes loader for libhuygens_cuda.so (the C ABI of include/huygens_cuda.h).

The library is built in-tree by `make -C huygens_b200/csrc` (see __graft_entry__.build()).
There is no fallback: if the shared object is missing, loading raises.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HG_LIB_PATH") or os.path.join(_HERE, "lib", "libhuygens_cuda.so")   # HG_LIB_PATH: experiment builds
_LIB = None

HG_OK, HG_ERR_INVALID, HG_ERR_UNSUPPORTED, HG_ERR_CUDA, HG_ERR_NOMEM, HG_ERR_NCCL = 0, 1, 2, 3, 4, 5
HG_SHARD_CHAIN, HG_SHARD_REDUCE = 0, 1
HG_F32, HG_F64 = 0, 1
HG_ARITH_STRICT, HG_ARITH_FMA = 0, 1
HG_MODE_AUTO, HG_MODE_EXACT, HG_MODE_SCAN = 0, 1, 2
HG_FB_FIXED = 1
HG_DIST_NONE, HG_DIST_SOFTCLIP = 0, 1
HG_SUM_TREE, HG_SUM_ORDERED = 0, 1
HG_EPI_NONE, HG_EPI_SOFTCLIP, HG_EPI_LIMITER = 0, 1, 2

# every symbol include/huygens_cuda.h declares: (name, restype, argtypes)
_vp, _i, _d, _sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
_pi, _pd = C.POINTER(C.c_int), C.POINTER(C.c_double)
SYMBOLS = [
    ("hg_last_error", C.c_char_p, []),
    ("hg_device_count", _i, []),
    ("hg_fma_peak", _i, [_i, _i, _pd, _pi, _pd]),
    ("hg_chain_latency", _i, [_i, _i, _pd]),
    ("hg_fb_create", _i, [C.POINTER(_vp), _i, _i, _i, _i, _i, _d, _d, _i, _i]),
    ("hg_fb_destroy", _i, [_vp]),
    ("hg_fb_set_coefficients", _i, [_vp, _i, _i, _vp, _i, _vp, _i]),
    ("hg_fb_set_boost", _i, [_vp, _i, _i, _vp]),
    ("hg_fb_set_mix", _i, [_vp, _i, _i, _vp]),
    ("hg_fb_open", _i, [_vp]),
    ("hg_fb_process", _i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _vp]),
    ("hg_fb_process_env", _i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _vp]),
    ("hg_fb_set_follower_mode", _i, [_vp, _i]),
    ("hg_fb_reset_followers", _i, [_vp]),
    ("hg_fb_process_events", _i, [_vp, _vp, _vp, _sz, _vp, _sz, _i, _vp]),
    ("hg_fb_process_tv", _i, [_vp, _vp, _vp, _sz, _vp, _sz, _i, _vp]),
    ("hg_fb_get_state", _i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    ("hg_fb_set_state", _i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    ("hg_fb_last_launch", _i, [_vp, _pi, _pi, _pi, _pi]),
    ("hg_fb_set_sum_mode", _i, [_vp, _i]),
    ("hg_fb_check", _i, [_vp]),
    ("hg_fb_set_epilogue", _i, [_vp, _i, _d, _d]),
    ("hg_fb_process_voice", _i, [_vp, _vp, _vp, _vp, _sz, _i, _vp]),
    ("hg_fb_debug_profile", _i, [_vp, _vp, _i]),
    ("hg_fb_debug_ru_config", _i, [_i, _i, _i, _i, _i, _i, _i, _vp]),
    ("hg_fb_debug_ws_config", _i, [_i, _i, _i, _i, _i, _vp]),
    ("hg_comm_get_unique_id", _i, [_vp]),
    ("hg_comm_init_rank", _i, [C.POINTER(_vp), _i, _i, _vp, _i]),
    ("hg_comm_init_all", _i, [C.POINTER(_vp), _i, _pi]),
    ("hg_comm_destroy", _i, [_vp]),
    ("hg_comm_group_start", _i, []),
    ("hg_comm_group_end", _i, []),
    ("hg_comm_rank", _i, [_vp]),
    ("hg_comm_size", _i, [_vp]),
    ("hg_fb_process_sharded", _i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _vp]),
    ("hg_osc_create", _i, [C.POINTER(_vp), _i, _i, _i, _i]),
    ("hg_osc_destroy", _i, [_vp]),
    ("hg_osc_freqmod", _i, [_vp, _i, _d]),
    ("hg_osc_set_frequencies", _i, [_vp, _i, _i, _vp]),
    ("hg_osc_activate", _i, [_vp, _vp, _i, _i]),
    ("hg_osc_open", _i, [_vp]),
    ("hg_osc_close", _i, [_vp]),
    ("hg_osc_active_count", _i, [_vp]),
    ("hg_osc_process", _i, [_vp, _vp, _vp, _vp, _sz, _vp]),
    ("hg_osc_process_env", _i, [_vp, _vp, _vp, _sz, _i, _vp, _sz, _vp]),
    ("hg_osc_get_phases", _i, [_vp, _vp]),
    ("hg_osc_set_phases", _i, [_vp, _vp]),
    ("hg_rms_create", _i, [C.POINTER(_vp), _i, _i, _i, _i, _i]),
    ("hg_rms_destroy", _i, [_vp]),
    ("hg_rms_process", _i, [_vp, _vp, _vp, _vp, _sz, _vp]),
    ("hg_rms_set_latch", _i, [_vp, _d, _d]),
    ("hg_rms_process_latch", _i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    ("hg_slide_create", _i, [C.POINTER(_vp), _i, _i, _i, _i, _vp, _i]),
    ("hg_stick_create", _i, [C.POINTER(_vp), _i, _i, _i, _i, _d, _i]),
    ("hg_stick_rmod", _i, [_vp, _d]),
    ("hg_casc_destroy", _i, [_vp]),
    ("hg_casc_process", _i, [_vp, _vp, _vp, _vp, _vp, _sz, _vp]),
]


class FbEvent(C.Structure):
    """hg_fb_event of include/huygens_cuda.h."""
    _fields_ = [("tick", C.c_size_t), ("kind", C.c_int), ("n0", C.c_int), ("count", C.c_int),
                ("a", C.c_void_p), ("na", C.c_int), ("b", C.c_void_p), ("nb", C.c_int)]


HG_EV_COEFFICIENTS, HG_EV_BOOST, HG_EV_MIX, HG_EV_OPEN = 0, 1, 2, 3


class HuygensError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"hg_status {status}: {message}")
        self.status = status


def build(verbose=False, jobs=8):
    """Compile libhuygens_cuda.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), f"-j{jobs}"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("libhuygens_cuda build failed:\n" + out.stdout[-4000:] + out.stderr[-4000:])
    if verbose:
        print(out.stdout)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(huygens_b200 has no CPU or PyTorch fallback)")
        _LIB = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(_LIB, name)
            fn.restype = res
            fn.argtypes = args
    return _LIB


def check(status):
    if status != HG_OK:
        raise HuygensError(status, lib().hg_last_error().decode(errors="replace"))
