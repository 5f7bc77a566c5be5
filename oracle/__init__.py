"""ctypes front-end of the parity oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under huygens_b200/ imports this package.

The classes mirror the reference's public surface (/root/reference/src/filterbank.h:16-188,
oscbank.h:16-97, rmsbank.h:16-85) plus the caller loop `out[t] = bank(in[t]); bank.tick()`
(tests/resynthesis.cpp:33-42) as `process`.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DTYPES = {np.dtype(np.float32): 0, np.dtype(np.float64): 1}


def build(verbose=False):
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if os.path.isdir("/root/reference/src") and os.path.exists(os.path.join(_HERE, "ref_capi.cpp")):
        out = subprocess.run(["make", "-C", _HERE, "ref"], capture_output=True, text=True)
        if out.returncode != 0:
            raise RuntimeError("oracle/_ref build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        for mode in ("strict", "fma"):
            f = lambda n: getattr(_LIB, f"{n}_{mode}")
            f("hgo_fb_create").restype = C.c_void_p
            f("hgo_fb_create").argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
            f("hgo_fb_destroy").argtypes = [C.c_void_p]
            f("hgo_fb_coefficients").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
            f("hgo_fb_boost").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
            f("hgo_fb_mix").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
            f("hgo_fb_open").argtypes = [C.c_void_p]
            f("hgo_fb_process").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int]
            f("hgo_fb_process_from").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int]
            f("hgo_fb_get_smoothed").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            f("hgo_epilogue").argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_double, C.c_double]
            f("hgo_osc_create").restype = C.c_void_p
            f("hgo_osc_create").argtypes = [C.c_int, C.c_int]
            f("hgo_osc_destroy").argtypes = [C.c_void_p]
            f("hgo_osc_freqmod").argtypes = [C.c_void_p, C.c_int, C.c_double]
            f("hgo_osc_open").argtypes = [C.c_void_p]
            f("hgo_osc_close").argtypes = [C.c_void_p]
            f("hgo_osc_activate").argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
            f("hgo_osc_active_count").argtypes = [C.c_void_p]
            f("hgo_osc_process").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
            f("hgo_osc_get_phases").argtypes = [C.c_void_p, C.c_void_p]
            f("hgo_rms_create").restype = C.c_void_p
            f("hgo_rms_create").argtypes = [C.c_int, C.c_int, C.c_int]
            f("hgo_rms_destroy").argtypes = [C.c_void_p]
            f("hgo_rms_process").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int]
            f("hgo_latch_create").restype = C.c_void_p
            f("hgo_latch_create").argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
            f("hgo_latch_destroy").argtypes = [C.c_void_p]
            f("hgo_latch_process").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
            f("hgo_latch_process_real").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
            f("hgo_slide_create").restype = C.c_void_p
            f("hgo_slide_create").argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
            f("hgo_stick_create").restype = C.c_void_p
            f("hgo_stick_create").argtypes = [C.c_int, C.c_int, C.c_int, C.c_double]
            f("hgo_casc_destroy").argtypes = [C.c_void_p]
            f("hgo_casc_process").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
            f("hgo_stick_rmod").argtypes = [C.c_void_p, C.c_double]
    return _LIB


_LIB_NATIVE = None


def lib_native():
    """liboracle_native.so: the strict restatement built with -march=native ON THIS HOST (make -C oracle native);
    only the Filterbank entry points are bound.  Used by bench.py's cpu_baseline for the second, fairer figure."""
    global _LIB_NATIVE
    if _LIB_NATIVE is None:
        out = subprocess.run(["make", "-B", "-C", _HERE, "native"], capture_output=True, text=True)   # always rebuilt: -march=native of another machine must never be loaded
        if out.returncode != 0:
            raise RuntimeError("oracle -march=native build failed:\n" + out.stdout + out.stderr)
        L = C.CDLL(os.path.join(_HERE, "liboracle_native.so"))
        ref = lib()
        for name in ("hgo_fb_create", "hgo_fb_destroy", "hgo_fb_coefficients", "hgo_fb_boost", "hgo_fb_mix", "hgo_fb_open",
                     "hgo_fb_process", "hgo_fb_process_from", "hgo_fb_get_smoothed"):
            f, g = getattr(L, name + "_strict"), getattr(ref, name + "_strict")
            f.restype, f.argtypes = g.restype, g.argtypes
        _LIB_NATIVE = L
    return _LIB_NATIVE


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _Base:
    def __init__(self, dtype, mode):
        self.dtype = np.dtype(dtype)
        assert self.dtype in DTYPES and mode in ("strict", "fma")
        self.mode = mode
        self._h = None

    native = False      # Filterbank(native=True): the -march=native build of the strict restatement (timing only)

    def _f(self, name):
        return getattr(lib_native() if self.native else lib(), f"{name}_{self.mode}")

    def _arr(self, v):
        return np.ascontiguousarray(np.asarray(v, dtype=np.float64).astype(self.dtype))


class Filterbank(_Base):
    """Filterbank<T>(order, N, k_p, k_g) — filterbank.h:36; fixed=True -> FFilterbank<T,N,order>(k_p,k_g) :203."""

    def __init__(self, order, N=1, k_p=0.1, k_g=1.0, dtype=np.float64, mode="strict", fixed=False, native=False):
        super().__init__(dtype, mode)
        self.native = bool(native)
        assert not (native and mode != "strict")
        self.N, self.order = N, order
        self._h = self._f("hgo_fb_create")(DTYPES[self.dtype], int(fixed), order, N, k_p, k_g)

    def __del__(self):
        if self._h:
            self._f("hgo_fb_destroy")(self._h)
            self._h = None

    def coefficients(self, n, forward, back):
        f, b = self._arr(forward), self._arr(back)
        self._f("hgo_fb_coefficients")(self._h, n, _ptr(f), len(f), _ptr(b), len(b))

    def boost(self, a, value=None):
        if value is None:
            v = self._arr(a)[: self.N]
            self._f("hgo_fb_boost")(self._h, 0, len(v), _ptr(v))
        else:
            v = self._arr([value])
            self._f("hgo_fb_boost")(self._h, int(a), 1, _ptr(v))

    def mix(self, a, value=None):
        if value is None:
            v = self._arr(a)[: self.N]
            self._f("hgo_fb_mix")(self._h, 0, len(v), _ptr(v))
        else:
            v = self._arr([value])
            self._f("hgo_fb_mix")(self._h, int(a), 1, _ptr(v))

    def open(self):
        self._f("hgo_fb_open")(self._h)

    def process(self, x, bands=False, distortion=0, want_sum=True):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        L = len(x)
        out = np.empty(L, self.dtype) if want_sum else None
        bd = np.empty((self.N, L), self.dtype) if bands else None
        self._f("hgo_fb_process")(self._h, _ptr(x), _ptr(out), _ptr(bd), L, distortion)
        return (out, bd) if bands else out

    def process_from(self, x, carry, distortion=0):
        """The bank sum continued from the running row `carry` (the sum over the filters that precede this
        bank's in a larger bank); carry = 0 gives process()."""
        x = np.ascontiguousarray(x, dtype=self.dtype)
        carry = np.ascontiguousarray(carry, dtype=self.dtype)
        out = np.empty(len(x), self.dtype)
        self._f("hgo_fb_process_from")(self._h, _ptr(x), _ptr(carry), _ptr(out), len(x), distortion)
        return out

    def smoothed(self):
        p = np.empty(self.N, self.dtype)
        g = np.empty(self.N, self.dtype)
        self._f("hgo_fb_get_smoothed")(self._h, _ptr(p), _ptr(g))
        return p, g


def epilogue(kind, x, s=None, dry=0.0, width=0.9):
    """The caller loops' post-processing per sample, in double like the reference: "softclip" ->
    softclip(dry*x + s, width) (tests/filterbank.cpp:205-207), "limiter" -> limiter(dry*x + s) (src/wave.h:150),
    "preclip" -> softclip(x, width).  x, s: arrays of one dtype; returns that dtype."""
    x = np.ascontiguousarray(x)
    s = None if s is None else np.ascontiguousarray(s, dtype=x.dtype)
    out = np.empty_like(x)
    k = {"softclip": 1, "limiter": 2, "preclip": 3}[kind]
    lib().hgo_epilogue_strict(k, DTYPES[np.dtype(x.dtype)], _ptr(x), _ptr(s), _ptr(out), len(x), float(dry), float(width))
    return out


def chained_sum(banks, x, block=4096, distortion=0):
    """Bank sum of the concatenation of `banks` (consecutive filter partitions of one bank), in the single
    bank's index order: partition p continues the running row of partition p-1 (Filterbank.process_from).
    One host thread per partition, pipelined over blocks of `block` samples — the same operations as one
    oracle bank of all the filters, bit for bit (tests/test_oracle_kat.py), at len(banks) times the speed."""
    import threading
    x = np.ascontiguousarray(x, dtype=banks[0].dtype)
    L = len(x)
    nb = (L + block - 1) // block
    rows = [np.zeros(L, banks[0].dtype)] + [np.empty(L, banks[0].dtype) for _ in banks]
    done = [[threading.Event() for _ in range(nb)] for _ in banks]
    errors = []

    def run(p):
        try:
            for j in range(nb):
                if p > 0:
                    done[p - 1][j].wait()
                a, b = j * block, min(L, (j + 1) * block)
                rows[p + 1][a:b] = banks[p].process_from(x[a:b], rows[p][a:b], distortion)
                done[p][j].set()
        except Exception as e:          # release the threads behind us
            errors.append(e)
            for j in range(nb):
                done[p][j].set()

    th = [threading.Thread(target=run, args=(p,)) for p in range(len(banks))]
    [t.start() for t in th]
    [t.join() for t in th]
    if errors:
        raise errors[0]
    return rows[-1]


class Oscbank(_Base):
    """Oscbank<T,N> — oscbank.h:16-97 (+ Multichannel activate/deactivate/open/close)."""

    def __init__(self, N, dtype=np.float64, mode="strict"):
        super().__init__(dtype, mode)
        self.N = N
        self._h = self._f("hgo_osc_create")(DTYPES[self.dtype], N)

    def __del__(self):
        if self._h:
            self._f("hgo_osc_destroy")(self._h)
            self._h = None

    def freqmod(self, index, target):
        self._f("hgo_osc_freqmod")(self._h, int(index), float(target))

    def open(self):
        self._f("hgo_osc_open")(self._h)

    def close(self):
        self._f("hgo_osc_close")(self._h)

    def activate(self, idx):
        a = np.ascontiguousarray(idx, dtype=np.int32)
        self._f("hgo_osc_activate")(self._h, _ptr(a), len(a), 1)

    def deactivate(self, idx):
        a = np.ascontiguousarray(idx, dtype=np.int32)
        self._f("hgo_osc_activate")(self._h, _ptr(a), len(a), 0)

    def active_count(self):
        return self._f("hgo_osc_active_count")(self._h)

    def process(self, L, phases=False, amps=None):
        """Per tick: mix[t] = mixdown(); [phases[t] = osc()]; [amix[t] = Mixer(Modbank(amps, osc()))]; tick()."""
        mix = np.empty(L, np.complex64 if self.dtype == np.float32 else np.complex128)
        ph = np.empty((L, self.N), mix.dtype) if phases else None
        am = None if amps is None else np.ascontiguousarray(amps, dtype=self.dtype)
        amix = None if amps is None else np.empty(L, self.dtype)
        self._f("hgo_osc_process")(self._h, _ptr(mix), _ptr(ph), _ptr(am), _ptr(amix), L)
        res = [mix]
        if phases:
            res.append(ph)
        if amps is not None:
            res.append(amix)
        return res[0] if len(res) == 1 else tuple(res)

    def phases(self):
        p = np.empty(self.N, np.complex64 if self.dtype == np.float32 else np.complex128)
        self._f("hgo_osc_get_phases")(self._h, _ptr(p))
        return p


class RMSbank(_Base):
    """RMSbank<T,N>(width = SR/20) — rmsbank.h:16-85."""

    def __init__(self, N, width=2400, dtype=np.float64, mode="strict"):
        super().__init__(dtype, mode)
        self.N, self.width = N, width
        self._h = self._f("hgo_rms_create")(DTYPES[self.dtype], N, width)

    def __del__(self):
        if self._h:
            self._f("hgo_rms_destroy")(self._h)
            self._h = None

    def process(self, zre, zim=None, calls_per_tick=1):
        """zre/zim: [N][L] channel-major.  Returns out[N][L]."""
        zre = np.ascontiguousarray(zre, dtype=self.dtype)
        zim = None if zim is None else np.ascontiguousarray(zim, dtype=self.dtype)
        N, L = zre.shape
        assert N == self.N
        out = np.empty((N, L), self.dtype)
        self._f("hgo_rms_process")(self._h, _ptr(zre), _ptr(zim), _ptr(out), L, calls_per_tick)
        return out


class _Cascade(_Base):
    """Shared block driver of Slidebank / Stickbank: complex channel-major input z[N][L] -> output [N][L]."""

    def __del__(self):
        if self._h:
            self._f("hgo_casc_destroy")(self._h)
            self._h = None

    def process(self, z):
        z = np.asarray(z)
        L = z.shape[1]
        assert z.shape[0] == self.N
        zre = np.ascontiguousarray(z.real, dtype=self.dtype)
        zim = np.ascontiguousarray(z.imag, dtype=self.dtype) if np.iscomplexobj(z) else None
        ore, oim = np.empty((self.N, L), self.dtype), np.empty((self.N, L), self.dtype)
        self._f("hgo_casc_process")(self._h, _ptr(zre), _ptr(zim), _ptr(ore), _ptr(oim), L)
        return ore + 1j * oim if self.dtype == np.float64 else (ore + 1j * oim).astype(np.complex64)


class Slidebank(_Cascade):
    """Slidebank<T,N>(order, radii) — slidebank.h:19-186: cascades of complex one-pole low-passes."""

    def __init__(self, N, order, radii, dtype=np.float64, mode="strict"):
        super().__init__(dtype, mode)
        self.N, self.order = N, max(1, order)
        r = np.asarray(radii, dtype=np.complex128)
        assert r.shape == (N,)
        rr, ri = np.ascontiguousarray(r.real), np.ascontiguousarray(r.imag)
        self._h = self._f("hgo_slide_create")(DTYPES[self.dtype], N, order, _ptr(rr), _ptr(ri))


class Stickbank(_Cascade):
    """Stickbank<T,N>(order, rad) — stickbank.h:15-221: N channels through one all-pole smoother."""

    def __init__(self, N, order, rad, dtype=np.float64, mode="strict"):
        super().__init__(dtype, mode)
        self.N, self.order = N, max(1, order)
        self._h = self._f("hgo_stick_create")(DTYPES[self.dtype], N, order, float(rad))

    def rmod(self, rad):
        self._f("hgo_stick_rmod")(self._h, float(rad))


class Latchbank(_Base):
    """Latchbank<T,N>(thresh, ratio) gating an RMSbank<T,N>(width) — latchbank.h:16-105.
    process(): per tick gated = latch(&rms, z); rms.tick().  process_real(): out = latch(level)."""

    def __init__(self, N, thresh, ratio=0.2, width=2400, dtype=np.float64, mode="strict"):
        super().__init__(dtype, mode)
        self.N = N
        self._h = self._f("hgo_latch_create")(DTYPES[self.dtype], N, width, float(thresh), float(ratio))

    def __del__(self):
        if self._h:
            self._f("hgo_latch_destroy")(self._h)
            self._h = None

    def process(self, zre, zim=None, level=False):
        zre = np.ascontiguousarray(zre, dtype=self.dtype)
        zim = None if zim is None else np.ascontiguousarray(zim, dtype=self.dtype)
        ore, oim = np.empty_like(zre), (None if zim is None else np.empty_like(zre))
        lv = np.empty_like(zre) if level else None
        self._f("hgo_latch_process")(self._h, _ptr(zre), _ptr(zim), _ptr(ore), _ptr(oim), _ptr(lv), zre.shape[1])
        res = [ore] if zim is None else [ore, oim]
        if level:
            res.append(lv)
        return res[0] if len(res) == 1 else tuple(res)

    def process_real(self, x):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        out = np.empty_like(x)
        self._f("hgo_latch_process_real")(self._h, _ptr(x), _ptr(out), x.shape[1])
        return out


# ---------------------------------------------------------------------------------------------
# oracle/_ref: the UNMODIFIED reference headers compiled against oracle/shim (see ref_capi.cpp).
# Used only to pin the restatement above (tests/test_oracle_vs_ref.py).
# ---------------------------------------------------------------------------------------------
_REF = None
REF_FIXED_N = (4, 16, 96)
REF_OSC_N = (1, 8, 96)
REF_RMS_N = (1, 8, 40)


def ref_lib():
    """Returns the ctypes handle of oracle/_ref/libhuygens_ref.so or None if it was never built
    (it is built where /root/reference exists and travels to the GPU box as a binary)."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libhuygens_ref.so")
        if not os.path.exists(path):
            if os.path.isdir("/root/reference/src"):
                build()
            if not os.path.exists(path):
                return None
        _REF = C.CDLL(path)
        vp, i, d, lg = C.c_void_p, C.c_int, C.c_double, C.c_long
        sig = {
            "hgr_relaxation": (d, [d]),
            "hgr_fb_create": (vp, [i, i, i, i, d, d]), "hgr_fb_destroy": (None, [vp]),
            "hgr_fb_coefficients": (None, [vp, i, vp, i, vp, i]), "hgr_fb_boost": (None, [vp, i, i, vp]),
            "hgr_fb_mix": (None, [vp, i, i, vp]), "hgr_fb_open": (None, [vp]),
            "hgr_fb_process": (None, [vp, vp, vp, vp, lg]), "hgr_fb_get_smoothed": (None, [vp, vp, vp]),
            "hgr_osc_create": (vp, [i, i]), "hgr_osc_destroy": (None, [vp]), "hgr_osc_freqmod": (None, [vp, i, d]),
            "hgr_osc_open": (None, [vp]), "hgr_osc_close": (None, [vp]), "hgr_osc_activate": (None, [vp, vp, i, i]),
            "hgr_osc_active_count": (i, [vp]), "hgr_osc_process": (None, [vp, vp, vp, lg]),
            "hgr_osc_get_phases": (None, [vp, vp]),
            "hgr_rms_create": (vp, [i, i, i]), "hgr_rms_destroy": (None, [vp]),
            "hgr_rms_process": (None, [vp, vp, vp, vp, lg, i]),
            "hgr_latch_create": (vp, [i, i, i, d, d]), "hgr_latch_destroy": (None, [vp]),
            "hgr_latch_process": (None, [vp, vp, vp, vp, vp, vp, lg]),
            "hgr_latch_process_real": (None, [vp, vp, vp, lg]),
            "hgr_slide_create": (vp, [i, i, i, vp, vp]), "hgr_stick_create": (vp, [i, i, i, d]),
            "hgr_casc_destroy": (None, [vp]), "hgr_casc_process": (None, [vp, vp, vp, vp, vp, lg]),
            "hgr_stick_rmod": (None, [vp, d]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(_REF, name)
            fn.restype = res
            fn.argtypes = args
    return _REF


class RefFilterbank:
    """soundmath::Filterbank<T> (fixed=False) or FFilterbank<T,N,2> (fixed=True, N in REF_FIXED_N) of
    the reference itself, driven through the callers' loop."""

    def __init__(self, order, N=1, k_p=0.1, k_g=1.0, dtype=np.float64, fixed=False):
        self.dtype, self.N = np.dtype(dtype), N
        self._h = ref_lib().hgr_fb_create(DTYPES[self.dtype], int(fixed), order, N, k_p, k_g)
        if not self._h:
            raise ValueError("this FFilterbank<T,N,order> is not instantiated in ref_capi.cpp")

    def __del__(self):
        if getattr(self, "_h", None):
            ref_lib().hgr_fb_destroy(self._h)
            self._h = None

    def _arr(self, v):
        return np.ascontiguousarray(np.asarray(v, dtype=np.float64).astype(self.dtype))

    def coefficients(self, n, forward, back):
        f, b = self._arr(forward), self._arr(back)
        ref_lib().hgr_fb_coefficients(self._h, n, _ptr(f), len(f), _ptr(b), len(b))

    def boost(self, a, value=None):
        v = self._arr(a)[: self.N] if value is None else self._arr([value])
        ref_lib().hgr_fb_boost(self._h, 0 if value is None else int(a), len(v), _ptr(v))

    def mix(self, a, value=None):
        v = self._arr(a)[: self.N] if value is None else self._arr([value])
        ref_lib().hgr_fb_mix(self._h, 0 if value is None else int(a), len(v), _ptr(v))

    def open(self):
        ref_lib().hgr_fb_open(self._h)

    def process(self, x, bands=False):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        out = np.empty(len(x), self.dtype)
        bd = np.empty((self.N, len(x)), self.dtype) if bands else None
        ref_lib().hgr_fb_process(self._h, _ptr(x), _ptr(out), _ptr(bd), len(x))
        return (out, bd) if bands else out

    def smoothed(self):
        p, g = np.empty(self.N, self.dtype), np.empty(self.N, self.dtype)
        ref_lib().hgr_fb_get_smoothed(self._h, _ptr(p), _ptr(g))
        return p, g


class RefOscbank:
    def __init__(self, N, dtype=np.float64):
        self.dtype, self.N = np.dtype(dtype), N
        self._h = ref_lib().hgr_osc_create(DTYPES[self.dtype], N)
        if not self._h:
            raise ValueError("this Oscbank<T,N> is not instantiated in ref_capi.cpp")

    def __del__(self):
        if getattr(self, "_h", None):
            ref_lib().hgr_osc_destroy(self._h)
            self._h = None

    def freqmod(self, i, f):
        ref_lib().hgr_osc_freqmod(self._h, int(i), float(f))

    def open(self):
        ref_lib().hgr_osc_open(self._h)

    def close(self):
        ref_lib().hgr_osc_close(self._h)

    def activate(self, idx):
        a = np.ascontiguousarray(idx, dtype=np.int32)
        ref_lib().hgr_osc_activate(self._h, _ptr(a), len(a), 1)

    def deactivate(self, idx):
        a = np.ascontiguousarray(idx, dtype=np.int32)
        ref_lib().hgr_osc_activate(self._h, _ptr(a), len(a), 0)

    def active_count(self):
        return ref_lib().hgr_osc_active_count(self._h)

    def process(self, L, phases=False):
        cdt = np.complex64 if self.dtype == np.float32 else np.complex128
        mix = np.empty(L, cdt)
        ph = np.empty((L, self.N), cdt) if phases else None
        ref_lib().hgr_osc_process(self._h, _ptr(mix), _ptr(ph), L)
        return (mix, ph) if phases else mix

    def phases(self):
        p = np.empty(self.N, np.complex64 if self.dtype == np.float32 else np.complex128)
        ref_lib().hgr_osc_get_phases(self._h, _ptr(p))
        return p


class RefRMSbank:
    def __init__(self, N, width=2400, dtype=np.float64):
        self.dtype, self.N = np.dtype(dtype), N
        self._h = ref_lib().hgr_rms_create(DTYPES[self.dtype], N, width)
        if not self._h:
            raise ValueError("this RMSbank<T,N> is not instantiated in ref_capi.cpp")

    def __del__(self):
        if getattr(self, "_h", None):
            ref_lib().hgr_rms_destroy(self._h)
            self._h = None

    def process(self, zre, zim=None, calls_per_tick=1):
        zre = np.ascontiguousarray(zre, dtype=self.dtype)
        zim = None if zim is None else np.ascontiguousarray(zim, dtype=self.dtype)
        out = np.empty_like(zre)
        ref_lib().hgr_rms_process(self._h, _ptr(zre), _ptr(zim), _ptr(out), zre.shape[1], calls_per_tick)
        return out


class RefLatchbank:
    """soundmath::Latchbank<T,N>(thresh, ratio) with its RMSbank<T,N>(width), N in REF_RMS_N."""

    def __init__(self, N, thresh, ratio=0.2, width=2400, dtype=np.float64):
        self.dtype, self.N = np.dtype(dtype), N
        self._h = ref_lib().hgr_latch_create(DTYPES[self.dtype], N, width, float(thresh), float(ratio))
        if not self._h:
            raise ValueError("this Latchbank<T,N> is not instantiated in ref_capi.cpp")

    def __del__(self):
        if getattr(self, "_h", None):
            ref_lib().hgr_latch_destroy(self._h)
            self._h = None

    def process(self, zre, zim=None, level=False):
        zre = np.ascontiguousarray(zre, dtype=self.dtype)
        zim = None if zim is None else np.ascontiguousarray(zim, dtype=self.dtype)
        ore, oim = np.empty_like(zre), (None if zim is None else np.empty_like(zre))
        lv = np.empty_like(zre) if level else None
        ref_lib().hgr_latch_process(self._h, _ptr(zre), _ptr(zim), _ptr(ore), _ptr(oim), _ptr(lv), zre.shape[1])
        res = [ore] if zim is None else [ore, oim]
        if level:
            res.append(lv)
        return res[0] if len(res) == 1 else tuple(res)

    def process_real(self, x):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        out = np.empty_like(x)
        ref_lib().hgr_latch_process_real(self._h, _ptr(x), _ptr(out), x.shape[1])
        return out


REF_CASC_N = (1, 8, 96)


class _RefCascade:
    def __del__(self):
        if getattr(self, "_h", None):
            ref_lib().hgr_casc_destroy(self._h)
            self._h = None

    def process(self, z):
        z = np.asarray(z)
        L = z.shape[1]
        zre = np.ascontiguousarray(z.real, dtype=self.dtype)
        zim = np.ascontiguousarray(z.imag, dtype=self.dtype) if np.iscomplexobj(z) else None
        ore, oim = np.empty((self.N, L), self.dtype), np.empty((self.N, L), self.dtype)
        ref_lib().hgr_casc_process(self._h, _ptr(zre), _ptr(zim), _ptr(ore), _ptr(oim), L)
        return ore + 1j * oim if self.dtype == np.float64 else (ore + 1j * oim).astype(np.complex64)


class RefSlidebank(_RefCascade):
    """soundmath::Slidebank<T,N>(order, radii), N in REF_CASC_N."""

    def __init__(self, N, order, radii, dtype=np.float64):
        self.dtype, self.N = np.dtype(dtype), N
        r = np.asarray(radii, dtype=np.complex128)
        rr, ri = np.ascontiguousarray(r.real), np.ascontiguousarray(r.imag)
        self._h = ref_lib().hgr_slide_create(DTYPES[self.dtype], N, order, _ptr(rr), _ptr(ri))
        if not self._h:
            raise ValueError("this Slidebank<T,N> is not instantiated in ref_capi.cpp")


class RefStickbank(_RefCascade):
    """soundmath::Stickbank<double,N>(order, rad), N in REF_CASC_N (the template only compiles for double)."""

    def __init__(self, N, order, rad, dtype=np.float64):
        self.dtype, self.N = np.dtype(dtype), N
        self._h = ref_lib().hgr_stick_create(DTYPES[self.dtype], N, order, float(rad))
        if not self._h:
            raise ValueError("this Stickbank<T,N> is not instantiated in ref_capi.cpp")

    def rmod(self, rad):
        ref_lib().hgr_stick_rmod(self._h, float(rad))
