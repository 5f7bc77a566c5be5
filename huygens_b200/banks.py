"""Python mirror of the reference's bank classes over the C ABI (ctypes).

Same member names and argument meaning as soundmath::Filterbank<T> (src/filterbank.h:16-188),
Oscbank<T,N> (src/oscbank.h) and RMSbank<T,N> (src/rmsbank.h), plus the block entry point
`process`.  Arrays may be numpy (host; copied inside the call) or torch CUDA tensors (device
resident; asynchronous on the current torch stream).  All arithmetic happens in
libhuygens_cuda.so; this file only marshals pointers.
"""
import ctypes as C

import numpy as np

from . import _lib as L

_NP2HG = {np.dtype(np.float32): L.HG_F32, np.dtype(np.float64): L.HG_F64}
_ARITH = {"strict": L.HG_ARITH_STRICT, "fma": L.HG_ARITH_FMA}
_MODE = {"auto": L.HG_MODE_AUTO, "exact": L.HG_MODE_EXACT, "scan": L.HG_MODE_SCAN}


def _is_torch(a):
    return type(a).__module__.startswith("torch")


def _torch_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(a):
    if a is None:
        return None
    if _is_torch(a):
        assert a.is_contiguous()
        return C.c_void_p(a.data_ptr())
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class _Handle:
    _destroy = None

    def __init__(self, dtype, arith):
        self.dtype = np.dtype(dtype)
        if self.dtype not in _NP2HG:
            raise ValueError("dtype must be float32 or float64")
        self.hg_dtype = _NP2HG[self.dtype]
        self.arith = _ARITH[arith]
        self._h = C.c_void_p()

    def close(self):
        if self._h:
            getattr(L.lib(), self._destroy)(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _arr(self, v):
        return np.ascontiguousarray(np.asarray(v, dtype=np.float64).astype(self.dtype))

    def _torch_dtype(self):
        import torch
        return torch.float32 if self.dtype == np.float32 else torch.float64


class Filterbank(_Handle):
    """Filterbank<T>(order, N, k_p, k_g) — src/filterbank.h:36; fixed=True gives
    FFilterbank<T,N,order>(k_p, k_g) smoothing semantics (:203,:291); streams>1 runs that many
    independent banks sharing coefficients (tests/filterbanks.cpp:185,195-208)."""
    _destroy = "hg_fb_destroy"

    def __init__(self, order, N=1, k_p=0.1, k_g=1.0, dtype=np.float64, arith="strict", streams=1,
                 fixed=False, device=0, sum_mode="ordered"):
        super().__init__(dtype, arith)
        self.order, self.N, self.streams, self.device = order, N, streams, device
        L.check(L.lib().hg_fb_create(C.byref(self._h), self.hg_dtype, self.arith, order, N, streams,
                                     k_p, k_g, L.HG_FB_FIXED if fixed else 0, device))
        if sum_mode != "ordered":
            self.set_sum_mode(sum_mode)

    def set_sum_mode(self, mode):
        """"ordered" (default): the mixdown is accumulated in filter-index order like the reference loop
        (src/filterbank.h:130) — bit-identical sums; "tree": per-CTA partial sums added in a fixed tree order."""
        L.check(L.lib().hg_fb_set_sum_mode(self._h, {"ordered": L.HG_SUM_ORDERED, "tree": L.HG_SUM_TREE}[mode]))

    def set_epilogue(self, kind=None, dry=0.0, width=0.9):
        """Fused post-processing of the bank sum: kind "softclip" -> softclip(dry*in + sum, width)
        (tests/filterbank.cpp:205-207), "limiter" -> 2/PI*atan(dry*in + sum) (src/wave.h:150), None -> off."""
        k = {None: L.HG_EPI_NONE, "softclip": L.HG_EPI_SOFTCLIP, "limiter": L.HG_EPI_LIMITER}[kind]
        L.check(L.lib().hg_fb_set_epilogue(self._h, k, float(dry), float(width)))

    def process_voice(self, x, dry_in=None, distortion=0):
        """out[t] = epilogue(dry * dry_in[t] + bank(x[t], distortion)) — tests/filterbank.cpp:200-215 in block form
        (set_epilogue first).  Host arrays."""
        x = np.ascontiguousarray(x, dtype=self.dtype)
        d = None if dry_in is None else np.ascontiguousarray(dry_in, dtype=self.dtype)
        out = np.empty_like(x)
        L.check(L.lib().hg_fb_process_voice(self._h, _ptr(x), _ptr(d), _ptr(out), x.shape[-1], distortion, None))
        return out

    def check(self):
        """Synchronous health check of the ordered mixdown (raises if its watchdog fired)."""
        L.check(L.lib().hg_fb_check(self._h))

    # -- reference setters ------------------------------------------------------------------
    def coefficients(self, n, forward, back):
        """coefficients(n, forward, back) (:73-82).  n may be an int (one filter; forward/back 1-D)
        or forward/back may be 2-D [count][k] starting at filter n."""
        f, b = self._arr(forward), self._arr(back)
        if f.ndim == 1:
            f, b = f[None, :], b[None, :]
        L.check(L.lib().hg_fb_set_coefficients(self._h, int(n), f.shape[0], _ptr(f), f.shape[1], _ptr(b), b.shape[1]))

    def boost(self, a, value=None):
        """boost(n, value) / boost(values) (:85-96)."""
        if value is None:
            v = self._arr(a)[: self.N]
            L.check(L.lib().hg_fb_set_boost(self._h, 0, len(v), _ptr(v)))
        else:
            L.check(L.lib().hg_fb_set_boost(self._h, int(a), 1, _ptr(self._arr([value]))))

    def mix(self, a, value=None):
        """mix(n, value) / mix(values) (:99-110)."""
        if value is None:
            v = self._arr(a)[: self.N]
            L.check(L.lib().hg_fb_set_mix(self._h, 0, len(v), _ptr(v)))
        else:
            L.check(L.lib().hg_fb_set_mix(self._h, int(a), 1, _ptr(self._arr([value]))))

    def open(self):
        L.check(L.lib().hg_fb_open(self._h))

    # -- block entry point --------------------------------------------------------------------
    def process(self, x, out=None, bands=False, mode="exact", distortion=0):
        """for t: out[t] = bank(x[t]); bank.tick().  x: [L] or [streams][L].
        Returns out (and bands [streams][N][L] when bands=True).  mode "exact" (default) is bit-identical to the
        reference loop for any block split; "scan" / "auto" may use the tolerance-bounded time-parallel scan."""
        if _is_torch(x):
            import torch
            assert x.is_cuda and x.dtype == self._torch_dtype() and x.is_contiguous()
            Ls = x.shape[-1]
            if out is None:
                out = torch.empty_like(x)
            bd = torch.empty((self.streams, self.N, Ls), dtype=x.dtype, device=x.device) if bands else None
            stream = _torch_stream()
        else:
            x = np.ascontiguousarray(x, dtype=self.dtype)
            Ls = x.shape[-1]
            if out is None:
                out = np.empty_like(x)
            bd = np.empty((self.streams, self.N, Ls), self.dtype) if bands else None
            stream = None
        assert x.size if not _is_torch(x) else x.numel()
        n_el = x.numel() if _is_torch(x) else x.size
        if n_el != self.streams * Ls:
            raise ValueError(f"input must hold streams*L = {self.streams}*{Ls} samples")
        L.check(L.lib().hg_fb_process(self._h, _ptr(x), _ptr(out), _ptr(bd), Ls, _MODE[mode], distortion, stream))
        if bands:
            if x.ndim == 1:
                bd = bd[0]
            return out, bd
        return out

    def process_sharded(self, comm, x, out=None, mode="chain", root=-1):
        """Collective render of a filter-partitioned bank (hg_fb_process_sharded): this handle holds rank
        comm.rank's consecutive slice of the filters, every rank passes the same input.  mode "chain" keeps the
        reference's index-order sum across the GPUs (bit-identical to one bank; result on the last rank unless `root`
        says otherwise), "reduce" adds order-free partial sums with ncclReduce (result on rank 0).  Returns `out` on
        the root rank, None elsewhere."""
        m = {"chain": L.HG_SHARD_CHAIN, "reduce": L.HG_SHARD_REDUCE}[mode]
        eff_root = root if root >= 0 else (comm.size - 1 if m == L.HG_SHARD_CHAIN else 0)
        if _is_torch(x):
            import torch
            Ls, stream = x.shape[-1], _torch_stream()
            if out is None and comm.rank == eff_root:
                out = torch.empty_like(x)
        else:
            x = np.ascontiguousarray(x, dtype=self.dtype)
            Ls, stream = x.shape[-1], None
            if out is None and comm.rank == eff_root:
                out = np.empty_like(x)
        L.check(L.lib().hg_fb_process_sharded(self._h, comm._h, _ptr(x), _ptr(out) if comm.rank == eff_root else None, Ls, m, root, stream))
        return out if comm.rank == eff_root else None

    def process_env(self, x, width=2400, hop=64, out=None, env=None, followers=None):
        """Block render + one RMSbank(width) envelope follower per (stream, filter), fused on the device.
        Returns (out_sum, env[streams][L // hop][N]); env rows are block samples hop-1, 2*hop-1, ...
        followers = "replica" (default: no ring traffic, ~2x faster) or "ring"; either way every tick must then be rendered
        through process_env), chosen at the first call."""
        if followers is not None:
            L.check(L.lib().hg_fb_set_follower_mode(self._h, {"ring": 1, "replica": 2}[followers]))
        if _is_torch(x):
            import torch
            Ls = x.shape[-1]
            out = torch.empty_like(x) if out is None else out
            env = torch.empty((self.streams, Ls // hop, self.N), dtype=x.dtype, device=x.device) if env is None else env
            stream = _torch_stream()
        else:
            x = np.ascontiguousarray(x, dtype=self.dtype)
            Ls = x.shape[-1]
            out = np.empty_like(x) if out is None else out
            env = np.empty((self.streams, Ls // hop, self.N), self.dtype) if env is None else env
            stream = None
        L.check(L.lib().hg_fb_process_env(self._h, _ptr(x), _ptr(out), _ptr(env), Ls, width, hop, stream))
        return out, env

    def process_events(self, x, events, mode="exact"):
        """Block render with setter calls at exact ticks (hg_fb_process_events): events = [(tick, kind, ...)] with
        ("coefficients", n, forward, back) | ("boost", n0, values) | ("mix", n0, values) | ("open",), ticks ascending.
        Equals calling the setter between ticks `tick - 1` and `tick` of the caller loop."""
        x = np.ascontiguousarray(x, dtype=self.dtype)
        out = np.empty_like(x)
        keep, evs = [], (L.FbEvent * len(events))()
        for i, (tick, kind, *args) in enumerate(events):
            ev = evs[i]
            ev.tick = int(tick)
            if kind == "coefficients":
                f, b = self._arr(args[1]), self._arr(args[2])
                if f.ndim == 1:
                    f, b = f[None, :], b[None, :]
                keep += [f, b]
                ev.kind, ev.n0, ev.count = L.HG_EV_COEFFICIENTS, int(args[0]), f.shape[0]
                ev.a, ev.na, ev.b, ev.nb = f.ctypes.data, f.shape[1], b.ctypes.data, b.shape[1]
            elif kind in ("boost", "mix"):
                v = np.atleast_1d(self._arr(args[1]))
                keep.append(v)
                ev.kind, ev.n0, ev.count, ev.a = (L.HG_EV_BOOST if kind == "boost" else L.HG_EV_MIX), int(args[0]), len(v), v.ctypes.data
            elif kind == "open":
                ev.kind = L.HG_EV_OPEN
            else:
                raise ValueError(f"unknown event kind {kind!r}")
        L.check(L.lib().hg_fb_process_events(self._h, _ptr(x), _ptr(out), x.shape[-1], C.cast(evs, C.c_void_p), len(events), _MODE[mode], None))
        return out

    def process_tv(self, x, rows, hop=1):
        """Block render with a parameter table: rows [n_rows][7][N] = b0 b1 b2 a1 a2 boost mix, row r replacing the
        bank's parameters before block tick r*hop (hg_fb_process_tv; hop = 1 is audio-rate modulation)."""
        x = np.ascontiguousarray(x, dtype=self.dtype)
        rows = np.ascontiguousarray(rows, dtype=self.dtype)
        if rows.ndim != 3 or rows.shape[1:] != (7, self.N):
            raise ValueError("rows must be [n_rows][7][N]")
        out = np.empty_like(x)
        L.check(L.lib().hg_fb_process_tv(self._h, _ptr(x), _ptr(out), x.shape[-1], _ptr(rows), rows.shape[0], int(hop), None))
        return out

    def process_bands(self, x, mode="exact"):
        """Per-band outputs only (no sum)."""
        x = np.ascontiguousarray(x, dtype=self.dtype)
        Ls = x.shape[-1]
        bd = np.empty((self.streams, self.N, Ls), self.dtype)
        L.check(L.lib().hg_fb_process(self._h, _ptr(x), None, _ptr(bd), Ls, _MODE[mode], 0, None))
        return bd[0] if x.ndim == 1 else bd

    # -- per-sample API of the reference (1-sample blocks; not a timed path) -------------------
    def __call__(self, sample):
        """T operator()(T sample) (:125-131); memoised until tick() like the reference."""
        if getattr(self, "_computed", None) is None:
            self._computed = self.process(np.full(self.streams, sample, self.dtype).reshape(self.streams, 1), mode="exact")
        return self._computed.reshape(-1)[0] if self.streams == 1 else self._computed.reshape(-1).copy()

    def tick(self):
        """tick() (:142-148).  The block kernel already advanced the ring; this clears the memo."""
        self._computed = None

    # -- state --------------------------------------------------------------------------------
    def get_state(self):
        sn = (self.streams, self.N)
        y1, y2, p, g = (np.empty(sn, self.dtype) for _ in range(4))
        xh = np.empty((self.streams, 2), self.dtype)
        L.check(L.lib().hg_fb_get_state(self._h, _ptr(y1), _ptr(y2), _ptr(p), _ptr(g), _ptr(xh)))
        return dict(y1=y1, y2=y2, p=p, g=g, xh=xh)

    def set_state(self, st):
        a = {k: np.ascontiguousarray(v, dtype=self.dtype) for k, v in st.items()}
        L.check(L.lib().hg_fb_set_state(self._h, _ptr(a.get("y1")), _ptr(a.get("y2")), _ptr(a.get("p")),
                                        _ptr(a.get("g")), _ptr(a.get("xh"))))

    def last_launch(self):
        m, n, f, g = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        L.check(L.lib().hg_fb_last_launch(self._h, C.byref(m), C.byref(n), C.byref(f), C.byref(g)))
        return dict(mode={1: "exact", 2: "scan"}.get(m.value, "none"), launches=n.value, filters_per_thread=f.value, ctas=g.value)


class Comm:
    """One rank of a filter-partitioned group of GPUs (hg_comm): `Comm.unique_id()` on rank 0, shared with the
    other ranks by any means (e.g. a torch.distributed broadcast), then `Comm(n_ranks, rank, uid, device)` on
    every rank; or `Comm.init_all(n)` in one process that drives several GPUs."""

    def __init__(self, n_ranks, rank, uid, device=0, _handle=None):
        self._h = C.c_void_p() if _handle is None else _handle
        if _handle is None:
            buf = (C.c_ubyte * 128).from_buffer_copy(bytes(uid))
            L.check(L.lib().hg_comm_init_rank(C.byref(self._h), int(n_ranks), int(rank), C.cast(buf, C.c_void_p), int(device)))
        self.rank, self.size = L.lib().hg_comm_rank(self._h), L.lib().hg_comm_size(self._h)

    @staticmethod
    def unique_id():
        buf = (C.c_ubyte * 128)()
        L.check(L.lib().hg_comm_get_unique_id(C.cast(buf, C.c_void_p)))
        return bytes(buf)

    @staticmethod
    def init_all(n_devices, devices=None):
        hs = (C.c_void_p * n_devices)()
        dv = None if devices is None else (C.c_int * n_devices)(*devices)
        L.check(L.lib().hg_comm_init_all(hs, n_devices, dv))
        return [Comm(n_devices, i, None, _handle=C.c_void_p(hs[i])) for i in range(n_devices)]

    def close(self):
        if self._h:
            L.lib().hg_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Oscbank(_Handle):
    """Oscbank<T,N> — src/oscbank.h:16-97 with Multichannel activate/deactivate/open/close."""
    _destroy = "hg_osc_destroy"

    def __init__(self, N, dtype=np.float64, arith="strict", device=0):
        super().__init__(dtype, arith)
        self.N = N
        L.check(L.lib().hg_osc_create(C.byref(self._h), self.hg_dtype, self.arith, N, device))

    def freqmod(self, index, target):
        L.check(L.lib().hg_osc_freqmod(self._h, int(index), float(target)))

    def set_frequencies(self, targets, n0=0):
        t = np.ascontiguousarray(targets, dtype=np.float64)
        L.check(L.lib().hg_osc_set_frequencies(self._h, n0, len(t), _ptr(t)))

    def activate(self, indices):
        a = np.ascontiguousarray(indices, dtype=np.int32)
        L.check(L.lib().hg_osc_activate(self._h, _ptr(a), len(a), 1))

    def deactivate(self, indices):
        a = np.ascontiguousarray(indices, dtype=np.int32)
        L.check(L.lib().hg_osc_activate(self._h, _ptr(a), len(a), 0))

    def open(self):
        L.check(L.lib().hg_osc_open(self._h))

    def close_all(self):
        L.check(L.lib().hg_osc_close(self._h))

    def active_count(self):
        return L.lib().hg_osc_active_count(self._h)

    def process(self, Ls, amps=None, want_mix=True, device_out=False, hop=None):
        """Ls ticks: mix[t] = mixdown() before tick t; amix[t] = sum_n amps[n] * Re(phase_n).
        With hop, amps is [rows][N] and row r drives ticks [r*hop, (r+1)*hop) (hg_osc_process_env)."""
        cdt = np.complex64 if self.dtype == np.float32 else np.complex128
        if device_out:
            import torch
            mix = torch.empty((Ls, 2), dtype=self._torch_dtype(), device="cuda") if want_mix else None
            amix = torch.empty(Ls, dtype=self._torch_dtype(), device="cuda") if amps is not None else None
            stream = _torch_stream()
        else:
            mix = np.empty(Ls, cdt) if want_mix else None
            amix = np.empty(Ls, self.dtype) if amps is not None else None
            stream = None
        am = None
        if amps is not None:
            am = amps if _is_torch(amps) else np.ascontiguousarray(amps, dtype=self.dtype)
        if hop is not None:
            L.check(L.lib().hg_osc_process_env(self._h, _ptr(mix), _ptr(am), int(am.shape[0]), int(hop), _ptr(amix), Ls, stream))
        else:
            L.check(L.lib().hg_osc_process(self._h, _ptr(mix), _ptr(am), _ptr(amix), Ls, stream))
        res = [r for r in (mix, amix) if r is not None]
        return res[0] if len(res) == 1 else tuple(res)

    def phases(self):
        p = np.empty(self.N, np.complex64 if self.dtype == np.float32 else np.complex128)
        L.check(L.lib().hg_osc_get_phases(self._h, _ptr(p)))
        return p

    def set_phases(self, p):
        p = np.ascontiguousarray(p, dtype=np.complex64 if self.dtype == np.float32 else np.complex128)
        L.check(L.lib().hg_osc_set_phases(self._h, _ptr(p)))


class RMSbank(_Handle):
    """RMSbank<T,N>(width = SR/20) — src/rmsbank.h:16-85."""
    _destroy = "hg_rms_destroy"

    def __init__(self, N, width=2400, dtype=np.float64, arith="strict", device=0):
        super().__init__(dtype, arith)
        self.N, self.width = N, width
        L.check(L.lib().hg_rms_create(C.byref(self._h), self.hg_dtype, self.arith, N, width, device))

    def process(self, zre, zim=None):
        """zre/zim: [N][L] channel-major; returns out[N][L] = sqrt(S/width) per tick."""
        if _is_torch(zre):
            import torch
            out = torch.empty_like(zre)
            stream = _torch_stream()
        else:
            zre = np.ascontiguousarray(zre, dtype=self.dtype)
            zim = None if zim is None else np.ascontiguousarray(zim, dtype=self.dtype)
            out = np.empty_like(zre)
            stream = None
        assert zre.shape[0] == self.N
        L.check(L.lib().hg_rms_process(self._h, _ptr(zre), _ptr(zim), _ptr(out), zre.shape[1], stream))
        return out


class Latchbank:
    """Latchbank<T,N>(thresh, ratio) gating its RMSbank<T,N>(width) — src/latchbank.h:16-105, used as
    `latchbank(&rmsbank, signal)` (tests/harmbank.cpp:84-99)."""

    def __init__(self, N, thresh, ratio=0.2, width=2400, dtype=np.float64, arith="strict", device=0):
        self.rms = RMSbank(N, width, dtype=dtype, arith=arith, device=device)
        self.N, self.dtype = N, self.rms.dtype
        L.check(L.lib().hg_rms_set_latch(self.rms._h, float(thresh), float(ratio)))

    def process(self, zre, zim=None, level=False):
        """zre/zim [N][L]; returns gated re (, gated im) (, level)."""
        zre = np.ascontiguousarray(zre, dtype=self.dtype)
        zim = None if zim is None else np.ascontiguousarray(zim, dtype=self.dtype)
        ore = np.empty_like(zre)
        oim = None if zim is None else np.empty_like(zre)
        lv = np.empty_like(zre) if level else None
        L.check(L.lib().hg_rms_process_latch(self.rms._h, _ptr(zre), _ptr(zim), _ptr(ore), _ptr(oim), _ptr(lv), zre.shape[1], None))
        res = [ore] if zim is None else [ore, oim]
        if level:
            res.append(lv)
        return res[0] if len(res) == 1 else tuple(res)


class _Cascade(_Handle):
    _destroy = "hg_casc_destroy"

    def process(self, z):
        """z [N][L] (complex or real) -> complex [N][L]: L ticks of `out = bank(&z); bank.tick()`."""
        z = np.asarray(z)
        Ls = z.shape[1]
        zre = np.ascontiguousarray(z.real, dtype=self.dtype)
        zim = np.ascontiguousarray(z.imag, dtype=self.dtype) if np.iscomplexobj(z) else None
        ore, oim = np.empty((self.N, Ls), self.dtype), np.empty((self.N, Ls), self.dtype)
        L.check(L.lib().hg_casc_process(self._h, _ptr(zre), _ptr(zim), _ptr(ore), _ptr(oim), Ls, None))
        out = np.empty((self.N, Ls), np.complex64 if self.dtype == np.float32 else np.complex128)
        out.real, out.imag = ore, oim
        return out


class Slidebank(_Cascade):
    """Slidebank<T,N>(order, radii) — src/slidebank.h:19-186."""

    def __init__(self, N, order, radii, dtype=np.float64, arith="strict", device=0):
        super().__init__(dtype, arith)
        self.N, self.order = N, max(1, order)
        r = np.ascontiguousarray(np.asarray(radii, dtype=np.complex128))
        if r.shape != (N,):
            raise ValueError("radii must hold one complex radius per channel")
        L.check(L.lib().hg_slide_create(C.byref(self._h), self.hg_dtype, self.arith, N, order, _ptr(r.view(np.float64)), device))


class Stickbank(_Cascade):
    """Stickbank<T,N>(order, rad) — src/stickbank.h:15-221."""

    def __init__(self, N, order, rad, dtype=np.float64, arith="strict", device=0):
        super().__init__(dtype, arith)
        self.N, self.order = N, max(1, order)
        L.check(L.lib().hg_stick_create(C.byref(self._h), self.hg_dtype, self.arith, N, order, float(rad), device))

    def rmod(self, rad):
        L.check(L.lib().hg_stick_rmod(self._h, float(rad)))


def fma_peak(dtype=np.float32, device=0):
    """Measured FMA issue rate (lane-FMA slots/s), SM count and SM clock (MHz) of `device`."""
    s, n, m = C.c_double(), C.c_int(), C.c_double()
    L.check(L.lib().hg_fma_peak(device, _NP2HG[np.dtype(dtype)], C.byref(s), C.byref(n), C.byref(m)))
    return dict(slots_per_s=s.value, sm_count=n.value, sm_mhz=m.value)


def chain_latency(dtype=np.float32, device=0):
    """Measured cycles per sample of the bit-exact recurrence chain (FMUL -> FADD -> FADD) of one lane."""
    c = C.c_double()
    L.check(L.lib().hg_chain_latency(device, _NP2HG[np.dtype(dtype)], C.byref(c)))
    return c.value
