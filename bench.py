#!/usr/bin/env python
"""bench.py — filter-samples/s of the biquad-bank hot path on B200 (BASELINE.json metric).

A step is one pass of the hot path over one block of synthetic input: a Filterbank<float> of 4096
log-spaced resonant biquads (tests/filterbank.cpp:255-271 recipe) rendering 60 s of a 48 kHz
chirp to its bank sum (BASELINE.json configs[1]).  With N > 1 GPUs the bank is filter-partitioned
(each rank owns 4096 more filters: weak scaling) and the partial sums are reduced to rank 0 with
one NCCL reduce per block — the path's only exchange step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode exact|scan]

`value`   device-resident throughput (CUDA events on the launching stream, L2 flushed between steps)
`e2e`     same metric through the public API with HOST buffers (pinned), H2D + D2H inside the timing
`roofline` FP32 FMA-pipe roofline: algorithmic 11 issue slots per filter-sample over
          the FFMA rate measured live by hg_fma_peak
`cpu_baseline` the oracle's strict tick loop on one host core over a bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 48000
SLOTS_PER_FILTER_SAMPLE = 11        # 11 mul + 7 add, every add fusable: SURVEY.md §8(d)
WORKLOADS = {
    # BASELINE configs[0]: Filterbank<double>, 64 resonant band-passes (resynthesis recipe), 10 s white noise
    # (f_i = 0.5 (i+1) SR / (N+1), R = 0.999, gains 1/|H| in the taps {g,0,-g}, boost(1) + open(), k_p = 0.1, k_g = 1)
    "c1": dict(N=64, seconds=10, recipe="resynthesis", r=0.999, k_p=0.1, k_g=1.0, input="white", dtype="f64", f_lo=0.0, f_hi=0.0),
    # name: (filters per GPU, seconds, f_lo, f_hi, r, k_p, k_g, input)
    "c2": dict(N=4096, seconds=60, f_lo=27.5, f_hi=20000.0, r=0.999, k_p=0.001, k_g=1.0, input="chirp"),
    # BASELINE configs[2] (65,536 filters; a parity/scaling case, not the headline): 10 s blocks
    "c3": dict(N=65536, seconds=10, f_lo=27.5, f_hi=20000.0, r=0.999, k_p=0.001, k_g=1.0, input="chirp"),
    # the per-GPU shards of configs[2] at 8 and 4 GPUs
    "c3_8192": dict(N=8192, seconds=10, f_lo=27.5, f_hi=20000.0, r=0.999, k_p=0.001, k_g=1.0, input="chirp"),
    "c3_16384": dict(N=16384, seconds=10, f_lo=27.5, f_hi=20000.0, r=0.999, k_p=0.001, k_g=1.0, input="chirp"),
    # BASELINE configs[3] without the envelope followers: 1024 independent streams x 1024 filters, 1 s blocks
    "c4": dict(N=1024, streams=1024, seconds=1, f_lo=27.5, f_hi=20000.0, r=0.999, k_p=0.001, k_g=1.0, input="white"),
    # BASELINE configs[4]: 65,536 sinusoid partials driven by Filterbank/RMSbank analysis of the input, 60 s
    # (run_c5: analysis bank + fused followers -> envelope rows -> Oscbank amplitudes; --strong splits the partials)
    "c5": dict(N=65536, seconds=60, f_lo=27.5, f_hi=20000.0, r=0.999, k_p=0.001, k_g=1.0, input="chirp"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="exact", choices=["exact", "scan", "auto"])
    ap.add_argument("--arith", default="strict", choices=["strict", "fma"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--seconds", type=float, default=None, help="override block length (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scan", action="store_true", help="skip the scan-mode side measurement")
    ap.add_argument("--strong", action="store_true", help="strong scaling: the workload's filters are split over the GPUs (BASELINE configs[2]) instead of N filters per GPU")
    ap.add_argument("--followers", default="replica", choices=["ring", "replica"], help="with --env: how y[t-width] is obtained (HBM ring, or lagged replica filters)")
    ap.add_argument("--sum", default="ordered", choices=["ordered", "tree"], help="mixdown: ordered = the reference's index order and rounding (bit-identical, default); tree = per-CTA partial sums (tolerance mode)")
    ap.add_argument("--exchange", default="chain", choices=["chain", "reduce"], help="N > 1, filter-partitioned: chain = the running bank-sum row crosses the GPUs in filter order (bit-identical to one bank); reduce = ncclReduce of order-free partial sums")
    ap.add_argument("--parity-samples", type=int, default=48000, help="prefix of the block checked against the oracle's own sequential sum (0: skip)")
    ap.add_argument("--no-side", action="store_true", help="N = 1: skip the side measurements of the other BASELINE configurations")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the configs[2] strong-scaling block")
    ap.add_argument("--env", action="store_true", help="fuse per-band RMSbank envelope followers (width 2400, hop 64) behind the filters")
    return ap.parse_args()


def bank_params(w, rank, world):
    from huygens_b200 import design
    from huygens_b200.sharding import filter_range
    if w.get("streams", 1) > 1:                            # stream-partitioned: the same bank on every GPU, no exchange
        rank, world = 0, 1
    lo, hi = filter_range(w["N"] * world, rank, world)     # rank's slice of an N*world bank (N filters per GPU)
    if w.get("recipe") == "resynthesis":                   # tests/resynthesis.cpp:48-54: gains in the taps, mix = open()
        _, fwd, back = design.resynthesis_bank(w["N"] * world, w["r"])
        return fwd[lo:hi], back[lo:hi], np.ones(hi - lo)
    freqs = design.log_spaced(w["N"] * world, w["f_lo"], w["f_hi"])[lo:hi]
    return design.resonator_bank(freqs, w["r"])


def make_bank(cls, w, rank, world, dtype, **kw):
    fwd, back, gains = bank_params(w, rank, world)
    fb = cls(2, w["N"], w["k_p"], w["k_g"], dtype=dtype, streams=w.get("streams", 1), **kw)
    fb.coefficients(0, fwd, back)
    fb.mix(gains)
    fb.boost(np.ones(w["N"]))
    return fb


def make_comm(rank, world, local):
    """hg_comm over the ranks of this torchrun job: rank 0's ncclUniqueId travels by torch.distributed"""
    import torch
    import torch.distributed as dist
    import huygens_b200 as hb
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(hb.Comm.unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, src=0)
    return hb.Comm(world, rank, bytes(uid.cpu().numpy().tobytes()), device=local)


def oracle_prefix(w, world, dtype, x, threads=None):
    """the oracle's own sequential bank sum of the WHOLE bank (all ranks' filters) over a prefix of the input:
    consecutive filter partitions chained on the host threads (oracle.chained_sum — the single bank's operations in
    the single bank's order)"""
    import oracle
    oracle.build()
    threads = threads or min(os.cpu_count() or 1, 32)
    ntot = w["N"] * world
    fwd, back, gains = bank_params(dict(w, N=ntot), 0, 1)
    banks = []
    for idx in np.array_split(np.arange(ntot), min(threads, ntot)):
        fb = oracle.Filterbank(2, len(idx), w["k_p"], w["k_g"], dtype=dtype, mode="strict")
        for j, n in enumerate(idx):
            fb.coefficients(j, fwd[n], back[n])
        fb.mix(gains[idx])
        fb.boost(np.ones(len(idx)))
        banks.append(fb)
    return oracle.chained_sum(banks, x)


def parity_block(got, ref, what):
    a, b = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    peak = float(np.max(np.abs(b))) or 1.0
    err = float(np.max(np.abs(a - b)))
    den = float(np.sum((a - b) ** 2))
    return {"vs": what, "samples": int(len(b)), "bit_identical": bool(np.array_equal(np.asarray(got), np.asarray(ref))),
            "max_abs_err_rel_peak": err / peak, "snr_db": None if den == 0 else float(10 * np.log10(np.sum(b * b) / den))}


def synth_input(w, L, dtype):
    from huygens_b200 import design
    S = w.get("streams", 1)
    if w["input"] == "chirp":
        x = design.chirp(L, dtype=dtype)
        return x if S == 1 else np.tile(x, (S, 1))
    if S == 1:
        return design.white_noise(L, dtype=dtype)
    return np.stack([design.white_noise(L, stream=s, dtype=dtype) for s in range(S)])


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [c for c in sm if mx and c > 0.5 * mx] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline(w, dtype, n_samples, threads, native=False):
    """Oracle strict tick loop over the head of the same input; `threads` bank partitions in parallel
    (the reference itself is single-threaded; partitioning filters is what a host could do).
    native: the same source built with -march=native on this host (BASELINE.md §2's second figure)."""
    import oracle
    oracle.build()
    N = w["N"]
    x = synth_input(w, int(w["seconds"] * SR), dtype)[:n_samples]
    fwd, back, gains = bank_params(w, 0, 1)
    parts = np.array_split(np.arange(N), threads)
    banks = []
    for idx in parts:
        fb = oracle.Filterbank(2, len(idx), w["k_p"], w["k_g"], dtype=dtype, mode="strict", native=native)
        for j, n in enumerate(idx):
            fb.coefficients(j, fwd[n], back[n])
        fb.mix(gains[idx])
        fb.boost(np.ones(len(idx)))
        banks.append(fb)
    outs = [None] * threads

    def run(i):
        outs[i] = banks[i].process(x)

    t0 = time.perf_counter()
    if threads == 1:
        run(0)
    else:
        th = [threading.Thread(target=run, args=(i,)) for i in range(threads)]
        [t.start() for t in th]
        [t.join() for t in th]
    dt = time.perf_counter() - t0
    return N * n_samples / dt, dt


def run_reference(args, w, rank, world):
    """--impl reference: the reference's CPU tick loop on all host cores.  The timed arm is the oracle
    restatement ("port"): the reference's own headers only compile against oracle/shim (Eigen is absent,
    DESIGN.md §2), whose eager temporaries make oracle/_ref much slower than the real Eigen build would be —
    timing it would flatter the GPU arm.  Its throughput is reported next to the port for transparency."""
    if rank != 0:
        return
    dtype = np.float64 if w.get("dtype") == "f64" else np.float32
    threads = os.cpu_count() or 1
    if w.get("streams", 1) == 1 and world > 1:
        # the whole job's bank: the GPU arm at N GPUs renders N x filters_per_gpu filters (weak scaling), so does this arm
        w = dict(w, N=w["N"] * world)
    n_samples = 2 * SR                       # bounded sample per step: 2 s of the block
    while w["N"] * n_samples > 8e9 and n_samples > 4800:
        n_samples //= 2
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(w, dtype, n_samples // 4, threads)
    times = []
    for _ in range(args.steps):
        _, dt = cpu_baseline(w, dtype, n_samples, threads)
        times.append(dt)
    T = float(np.mean(times))
    val = w["N"] * n_samples / T
    sample = (f"{w['N']} filters (the whole {world}-GPU job's bank) x first {n_samples} samples of the {w['seconds']} s {w['input']} per step, "
              f"{threads} threads over filter partitions")
    ref_shim = None
    try:
        import oracle
        if oracle.ref_lib() is not None:
            n_ref = 2400
            fwd, back, gains = bank_params(w, 0, 1)
            fb = oracle.RefFilterbank(2, w["N"], w["k_p"], w["k_g"], dtype=dtype)
            for n in range(w["N"]):
                fb.coefficients(n, fwd[n], back[n])
            fb.mix(gains)
            fb.boost(np.ones(w["N"]))
            x = synth_input(w, int(w["seconds"] * SR), dtype)[:n_ref]
            t0 = time.perf_counter()
            fb.process(x)
            dt = time.perf_counter() - t0
            ref_shim = {"value": w["N"] * n_ref / dt, "unit": "filter-samples/s", "cores": 1, "kind": "reference headers compiled against oracle/shim (not Eigen)",
                        "sample": f"{w['N']} filters x first {n_ref} samples, {dt:.1f} s"}
    except Exception:
        ref_shim = None
    line = {
        "impl": "reference", "metric": "filter-samples/s (biquad bank)", "value": val, "unit": "filter-samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": T * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": w.get("dtype", "f32"), "data": "synthetic",
        "config": dict(workload_config(dict(w, N=w["N"] // (world if w.get("streams", 1) == 1 else 1)), args, world), filters_timed=w["N"], samples_timed_per_step=n_samples),
        "cpu_baseline": {"value": val, "unit": "filter-samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "filter-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if ref_shim:
        line["reference_headers_via_shim"] = ref_shim
    print(json.dumps(line), flush=True)


def run_c5(args, w, rank, world, local):
    """BASELINE configs[4]: Oscbank additive synthesis driven by analysis envelopes.  One step = the whole block:
    Filterbank(N) with one RMSbank(2400) follower per band (hop 64, replica followers) -> envelope rows ->
    Oscbank(N) with one amplitude row per hop -> mixed output; ranks hold N filters + N partials each and
    reduce the mix once per block.  Unit: partial-samples/s (a partial-sample includes its analysis band)."""
    import torch
    import torch.distributed as dist
    import huygens_b200 as hb
    from huygens_b200 import design
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, L, hop = w["N"], int(w["seconds"] * SR), 64
    fb = make_bank(hb.Filterbank, w, rank, world, np.float32, arith=args.arith, device=local)
    from huygens_b200.sharding import filter_range
    lo, hi = filter_range(N * world, rank, world)
    freqs = design.log_spaced(N * world, w["f_lo"], w["f_hi"])[lo:hi]
    osc = hb.Oscbank(N, dtype=np.float32, arith=args.arith, device=local)
    osc.set_frequencies(freqs)
    osc.open()
    x = torch.from_numpy(synth_input(w, L, np.float32)).cuda()
    out = torch.empty_like(x)
    env = torch.empty((1, L // hop, N), dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    peak = hb.fma_peak(np.float32, device=local)
    mix = [None]

    def step():
        fb.process_env(x, width=2400, hop=hop, out=out, env=env, followers="replica")
        mix[0] = osc.process((L // hop) * hop, amps=env[0], hop=hop, want_mix=False, device_out=True)
        if world > 1:
            dist.reduce(mix[0], dst=0, op=dist.ReduceOp.SUM)

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        ev[i][0].record()
        fb.process_env(x, width=2400, hop=hop, out=out, env=env, followers="replica")
        ev[i][1].record()
        mix[0] = osc.process((L // hop) * hop, amps=env[0], hop=hop, want_mix=False, device_out=True)
        if world > 1:
            dist.reduce(mix[0], dst=0, op=dist.ReduceOp.SUM)
        ev[i][2].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_an = sum(a.elapsed_time(b) for a, b, _ in ev) * 1e-3
    t_all = sum(a.elapsed_time(c) for a, _, c in ev) * 1e-3
    if world > 1:
        tt = torch.tensor([t_all, t_an], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_all, t_an = tt.tolist()
    # end to end: host input in, mixed audio out (the envelope rows never leave the device)
    xh = x.cpu().pin_memory().numpy()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fb.process_env(xh, width=2400, hop=hop, env=env, followers="replica")
        m = osc.process((L // hop) * hop, amps=env[0], hop=hop, want_mix=False, device_out=True)
        m_host = m.cpu()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    units = float(N) * world * L * args.steps
    slots = SLOTS_PER_FILTER_SAMPLE + 3 + 11            # band filter + follower + phasor tick (9 slots + 2 divisions)
    per_gpu = float(N) * L * args.steps / t_all
    line = {
        "metric": "partial-samples/s (analysis band + follower + Oscbank partial)", "value": units / t_all, "unit": "partial-samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_all / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong" if (args.strong and world > 1) else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{N} analysis biquads + RMSbank(2400) followers -> {N} Oscbank partials per GPU, {w['seconds']} s 48 kHz chirp, hop 64",
                   "partials_total": N * world, "samples": L, "partition": "filters/partials" if world > 1 else "none",
                   "arith": args.arith, "l2": "flushed between steps (256 MiB write)"},
        "e2e": {"value": units / t_e2e, "unit": "partial-samples/s", "h2d_bytes_per_step": int(L * 4), "d2h_bytes_per_step": int(m_host.numel() * 4)},
        "gpu_launches": None, "clocks": clocks,
        "roofline": {"bound": "fp32_fma", "achieved": per_gpu * slots / 1e12, "peak": peak["slots_per_s"] / 1e12, "unit": "Tslot/s",
                     "frac": per_gpu * slots / peak["slots_per_s"], "traffic": None, "slots_per_unit": slots,
                     "analysis_ms": t_an / args.steps * 1e3, "synthesis_ms": (t_all - t_an) / args.steps * 1e3,
                     "peak_source": "hg_fma_peak FFMA microbenchmark, measured in this run"},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def strong_c3(args, rank, world, local, comm, flush):
    """BASELINE configs[2]: Filterbank<float> of 65,536 filters partitioned over the GPUs, one block of 10 s.
    Times the chain exchange (bit-identical) and the ncclReduce exchange, and — on rank 0 alone — the whole bank on
    one GPU, all in this run; parity of both exchanges against the oracle's sequential sum on a prefix."""
    import torch
    import torch.distributed as dist
    import huygens_b200 as hb
    w = dict(WORKLOADS["c3"])
    ntot, L = w["N"], int(w["seconds"] * SR)
    if ntot % world:
        return {"skipped": f"{ntot} filters do not split over {world} GPUs"}
    w["N"] = ntot // world
    x_np = synth_input(w, L, np.float32)
    x = torch.from_numpy(x_np).cuda()
    root = world - 1
    out = torch.empty(L, dtype=torch.float32, device="cuda")
    res = {"workload": f"Filterbank<float> {ntot} filters over {world} GPUs ({w['N']} per GPU), {w['seconds']} s 48 kHz chirp, summed output",
           "filters_total": ntot, "samples": L}
    n_par = min(L, 24000)
    ref = oracle_prefix(w, world, np.float32, x_np[:n_par]) if rank in (0, root) else None
    for mode in ("chain", "reduce"):
        fb = make_bank(hb.Filterbank, w, rank, world, np.float32, arith=args.arith, device=local)
        dst = world - 1 if mode == "chain" else 0
        for _ in range(3):
            fb.process_sharded(comm, x, out=out if rank == dst else None, mode=mode)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i in range(args.steps):
            flush.fill_(i & 0xFF)
            dist.barrier()
            ev[i][0].record()
            fb.process_sharded(comm, x, out=out if rank == dst else None, mode=mode)
            ev[i][1].record()
        torch.cuda.synchronize()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) * 1e-3 / args.steps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # parity: a fresh bank renders the prefix (the timed bank's state has moved on)
        fbp = make_bank(hb.Filterbank, w, rank, world, np.float32, arith=args.arith, device=local)
        fbp.process_sharded(comm, x[:n_par].contiguous(), out=out[:n_par] if rank == dst else None, mode=mode)
        torch.cuda.synchronize()
        box = [parity_block(out[:n_par].cpu().numpy(), ref, f"oracle sequential sum of all {ntot} filters, first {n_par} samples") if rank == dst else None]
        dist.broadcast_object_list(box, src=dst)
        info = fb.last_launch()
        res[mode] = {"ms_per_block": t.item() * 1e3, "value": float(ntot) * L / t.item(), "unit": "filter-samples/s",
                     "kernel": {0: "K1-ws", -2: "K1-ru"}.get(info["filters_per_thread"], "K1 lanes"), "ctas_per_gpu": info["ctas"], "parity": box[0]}
        del fb, fbp
    # the same bank on ONE GPU (rank 0 alone, the others wait)
    one = [None]
    if rank == 0:
        w1 = dict(WORKLOADS["c3"])
        fb1 = make_bank(hb.Filterbank, w1, 0, 1, np.float32, arith=args.arith, device=local)
        for _ in range(3):
            fb1.process(x, out=out)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i in range(args.steps):
            flush.fill_(i & 0xFF)
            ev[i][0].record()
            fb1.process(x, out=out)
            ev[i][1].record()
        torch.cuda.synchronize()
        one[0] = sum(a.elapsed_time(b) for a, b in ev) * 1e-3 / args.steps
        del fb1
    dist.barrier()
    dist.broadcast_object_list(one, src=0)
    res["one_gpu_ms_per_block"] = one[0] * 1e3
    for mode in ("chain", "reduce"):
        res[mode]["speedup_vs_one_gpu"] = one[0] * 1e3 / res[mode]["ms_per_block"]
        res[mode]["efficiency"] = res[mode]["speedup_vs_one_gpu"] / world
    return res


def side_measurements(args, local, flush, peak):
    """The other BASELINE configurations on this GPU, timed in the same run so that the driver's record carries them
    (each a few short launches): configs[2] on one GPU, configs[3] with and without the fused followers, the Oscbank
    and RMSbank kernels alone.  Device-resident inputs, CUDA events, L2 flushed between launches."""
    import torch
    import huygens_b200 as hb
    from huygens_b200 import design

    def timed(fn, reps=3):
        fn()
        fn()
        torch.cuda.synchronize()
        ts = []
        for i in range(reps):
            flush.fill_(i & 0xFF)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e-3)
        return float(np.mean(ts))

    res = {}
    slots = peak["slots_per_s"]
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm = 6650.0
    for name in ("c3", "c4"):
        w = dict(WORKLOADS[name])
        L, S = int(w["seconds"] * SR), w.get("streams", 1)
        fb = make_bank(hb.Filterbank, w, 0, 1, np.float32, arith=args.arith, device=local)
        x = torch.from_numpy(synth_input(w, L, np.float32)).cuda()
        out = torch.empty_like(x)
        t = timed(lambda: fb.process(x, out=out))
        fs = float(w["N"]) * S * L / t
        res[name] = {"workload": workload_config(w, args, 1)["workload"], "ms_per_block": t * 1e3, "value": fs, "unit": "filter-samples/s",
                     "roofline_frac": fs * SLOTS_PER_FILTER_SAMPLE / slots, "kernel": {0: "K1-ws", -2: "K1-ru"}.get(fb.last_launch()["filters_per_thread"], "lanes")}
        if name == "c4":
            env = torch.empty((S, L // 64, w["N"]), dtype=torch.float32, device="cuda")
            fbe = make_bank(hb.Filterbank, w, 0, 1, np.float32, arith=args.arith, device=local)
            t = timed(lambda: fbe.process_env(x, width=2400, hop=64, out=out, env=env))
            fs = float(w["N"]) * S * L / t
            res["c4_followers"] = {"workload": "the same + one RMSbank(2400) follower per (stream, band), envelopes every 64 samples (replica followers)",
                                   "ms_per_block": t * 1e3, "value": fs, "unit": "filter-samples/s", "roofline_frac": fs * (SLOTS_PER_FILTER_SAMPLE + 3) / slots,
                                   "slots_per_unit": SLOTS_PER_FILTER_SAMPLE + 3}
            del fbe, env
        del fb, x, out
    N, Lo = 65536, 480000
    osc = hb.Oscbank(N, dtype=np.float32, arith=args.arith, device=local)
    osc.set_frequencies(design.log_spaced(N, 20.0, 20000.0))
    osc.open()
    amps = torch.full((N,), 1.0 / N, dtype=torch.float32, device="cuda")
    t = timed(lambda: osc.process(Lo, amps=amps, want_mix=True, device_out=True))
    ps = float(N) * Lo / t
    res["oscbank"] = {"workload": f"Oscbank<float> {N} partials x {Lo} ticks, mixdown() + amplitude-weighted mix per tick (index-order sums)",
                      "ms_per_block": t * 1e3, "value": ps, "unit": "partial-samples/s", "roofline_frac": ps * 11 / slots,
                      "slots_per_unit": "9 FP32 slots + 2 IEEE divisions counted as 1 each (SURVEY.md 8d)",
                      "issued_instructions_per_partial_tick": {"value": 50.6, "source": "profiles/r2_ncu_osc.txt: smsp__inst_executed.sum x 32 / (partials x ticks) of one ncu --set full launch; ~22 of them are the tick itself (SASS), the rest the index-order mixdowns and their hand-overs"}}
    del osc
    C_, Lr = 65536, 96000
    z = torch.rand((C_, Lr), dtype=torch.float32, device="cuda")
    rms = hb.RMSbank(C_, 2400, dtype=np.float32, device=local)
    t = timed(lambda: rms.process(z))
    res["rmsbank"] = {"workload": f"RMSbank<float> {C_} channels x {Lr} ticks, width 2400", "ms_per_block": t * 1e3,
                      "value": float(C_) * Lr / t, "unit": "channel-samples/s", "hbm_GBps_algorithmic": C_ * Lr * 8 / t / 1e9,
                      "roofline_frac_hbm": C_ * Lr * 8 / t / 1e9 / hbm}
    return res


def workload_config(w, args, world):
    tname = "double" if w.get("dtype") == "f64" else "float"
    sname = f"{w['streams']} streams x " if w.get("streams", 1) > 1 else ""
    bank = (f"resonant biquads/GPU of the resynthesis recipe (f_i = 0.5 (i+1) SR/(N+1), R={w['r']}, gains in the taps, boost(1) + open(), "
            if w.get("recipe") == "resynthesis" else
            f"log-spaced resonant biquads/GPU ({w['f_lo']} Hz-{w['f_hi']} Hz, r={w['r']}, fwd {{1,0,-1}}, ")
    return {"workload": f"Filterbank<{tname}> {sname}{w['N']} {bank}k_p={w['k_p']}, k_g={w['k_g']}), {w['seconds']} s 48 kHz {w['input']}, summed output",
            "filters_per_gpu": w["N"], "filters_total": w["N"] * (1 if w.get("streams", 1) > 1 else world),
            "streams_total": w.get("streams", 1) * (world if w.get("streams", 1) > 1 else 1), "samples": int(w["seconds"] * SR),
            "partition": ("streams" if w.get("streams", 1) > 1 else "filters") if world > 1 else "none", "mode": args.mode, "arith": args.arith,
            "sum": args.sum, "l2": "flushed between steps (256 MiB write)"}


def main():
    args = parse()
    w = dict(WORKLOADS[args.workload])
    if args.seconds:
        w["seconds"] = args.seconds
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.strong and world > 1 and w.get("streams", 1) == 1:
        if w["N"] % world:
            raise SystemExit(f"--strong: {w['N']} filters do not split over {world} GPUs")
        w["N"] //= world
    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return
    if args.workload == "c5":
        run_c5(args, w, rank, world, local)
        return

    import torch
    import torch.distributed as dist
    import huygens_b200 as hb

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dtype, tdtype = (np.float64, torch.float64) if w.get("dtype") == "f64" else (np.float32, torch.float32)
    L = int(w["seconds"] * SR)
    N = w["N"]
    fb = make_bank(hb.Filterbank, w, rank, world, dtype, arith=args.arith, device=local, sum_mode=args.sum)
    S = w.get("streams", 1)
    x_host = torch.from_numpy(synth_input(w, L, dtype)).pin_memory()
    out_host = torch.empty(x_host.shape, dtype=tdtype).pin_memory()
    x = x_host.cuda(non_blocking=True)
    out = torch.empty(x_host.shape, dtype=tdtype, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    peak = hb.fma_peak(dtype, device=local)
    chain_cycles = hb.chain_latency(dtype, device=local)

    env_buf = torch.empty((S, L // 64, N), dtype=tdtype, device="cuda") if args.env else None
    exchange = world > 1 and S == 1            # filter-partitioned: the running row crosses the GPUs; stream-partitioned: no exchange
    comm = make_comm(rank, world, local) if exchange else None
    shard_mode = "reduce" if args.exchange == "reduce" else "chain"
    root = world - 1 if shard_mode == "chain" else 0      # where the finished bank sum lands

    def render(xi, oi):
        if args.env:
            fb.process_env(xi, width=2400, hop=64, out=oi, env=env_buf, followers=args.followers)
        elif exchange:
            # one kernel per GPU; in chain mode the last CTA of rank r hands the running row to rank r+1 through
            # peer memory (hg_fb_process_sharded), in reduce mode ncclReduce adds the ranks' partial sums
            fb.process_sharded(comm, xi, out=oi, mode=shard_mode)
        else:
            fb.process(xi, out=oi, mode=args.mode)

    def step():
        render(x, out)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    launches_per_step = fb.last_launch()["launches"]
    info = fb.last_launch()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.fill_(i & 0xFF)                    # L2 flush, outside the timed events
        if world > 1:
            dist.barrier()                       # the ranks of a chain start a block together
        ev[i][0].record()
        render(x, out)
        ev[i][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_total = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
    t_kernel = t_total
    if world > 1:
        tt = torch.tensor([t_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_total = t_kernel = tt.item()

    # parity: a fresh bank (zero state, like the oracle's) renders a prefix of the block through the same path; the
    # result is compared with the oracle's own sequential sum of the WHOLE bank (all ranks' filters)
    parity = None
    n_par = min(L, args.parity_samples)
    if n_par > 0 and not args.env and args.mode == "exact" and S == 1:
        fbp = make_bank(hb.Filterbank, w, rank, world, dtype, arith=args.arith, device=local, sum_mode=args.sum)
        xp = x[:n_par].contiguous()
        if exchange:
            fbp.process_sharded(comm, xp, out=out[:n_par] if rank == root else None, mode=shard_mode)
        else:
            fbp.process(xp, out=out[:n_par], mode=args.mode)
        torch.cuda.synchronize()
        if rank == root:
            ref = oracle_prefix(w, world, dtype, x_host.numpy()[:n_par])
            parity = parity_block(out[:n_par].cpu().numpy(), ref,
                                  f"oracle (strict, sequential index-order sum of all {N * world} filters), first {n_par} samples")
        del fbp
    if world > 1:
        box = [parity]
        dist.broadcast_object_list(box, src=root)
        parity = box[0]

    # end to end through the public API with host buffers (H2D + D2H inside the call)
    xh = x_host.numpy()
    oh = out_host.numpy() if (not exchange or rank == root) else None
    env_h = torch.empty(env_buf.shape, dtype=tdtype).pin_memory().numpy() if args.env else None

    def render_host():
        if args.env:
            fb.process_env(xh, width=2400, hop=64, out=oh, env=env_h, followers=args.followers)   # envelopes come back to the host too
        elif exchange:
            fb.process_sharded(comm, xh, out=oh, mode=shard_mode)      # every rank uploads the block; the root downloads the sum
        else:
            fb.process(xh, out=oh, mode=args.mode)

    for _ in range(2):
        render_host()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        render_host()                                # synchronises before returning
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = tt.item()

    # N > 1: BASELINE configs[2] — 65,536 filters split over the GPUs (strong scaling), chain and reduce exchange,
    # against the same bank on one GPU measured in this run, with parity against the oracle on a prefix
    strong = strong_c3(args, rank, world, local, comm, flush) if (exchange and args.workload == "c2" and not args.no_strong) else None

    # side measurement (N = 1 only): the same block through the time-parallel scan, whose result is
    # tolerance-bounded instead of bit-identical
    scan_info = None
    if world == 1 and args.mode == "exact" and not args.no_scan and S == 1:
        out_exact = out.clone()
        for _ in range(3):
            fb.process(x, out=out, mode="scan")
        sev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i in range(args.steps):
            flush.fill_(i & 0xFF)
            sev[i][0].record()
            fb.process(x, out=out, mode="scan")
            sev[i][1].record()
        torch.cuda.synchronize()
        t_scan = sum(a.elapsed_time(b) for a, b in sev) * 1e-3 / args.steps
        fb.process(x, out=out_exact, mode="exact")
        fb.process(x, out=out, mode="scan")
        diff = (out.double() - out_exact.double()).abs().max().item() / out_exact.double().abs().max().item()
        scan_info = {"value": float(N) * L / t_scan, "unit": "filter-samples/s", "ms_per_step": t_scan * 1e3,
                     "roofline_frac": float(N) * L / t_scan * SLOTS_PER_FILTER_SAMPLE / peak["slots_per_s"],
                     "max_abs_diff_vs_exact_rel_peak": diff,
                     "note": "not bit-identical; measured error vs the oracle in profiles/scan_error_r2.json and DESIGN.md §5"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    units = float(N) * S * world * L * args.steps
    value = units / t_total
    per_gpu_kernel = float(N) * S * L * args.steps / t_kernel
    achieved = per_gpu_kernel * SLOTS_PER_FILTER_SAMPLE
    nominal = peak["sm_count"] * (64 if w.get("dtype") == "f64" else 128) * 1.965e9
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        if args.mode == "exact" and args.workload == "c2":
            traffic = {"dram_bytes_per_launch": tr["dram_bytes_read"] + tr["dram_bytes_write"], "samples_per_launch": tr["samples_per_launch"],
                       "algorithmic_bytes_per_launch": 8 * tr["samples_per_launch"], "source": tr["source"], "note": tr["note"],
                       "recurrence_warp": tr.get("r_warp")}
    except Exception:
        pass
    line = {
        "metric": "filter-samples/s (biquad bank)", "value": value, "unit": "filter-samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong" if (args.strong and world > 1 and S == 1) else "weak", "vs_baseline": None,
        "dtype": w.get("dtype", "f32"), "data": "synthetic",
        "config": workload_config(w, args, world),
        "e2e": {"value": units / t_e2e, "unit": "filter-samples/s", "h2d_bytes_per_step": int(S * L * x_host.element_size()), "d2h_bytes_per_step": int(S * L * x_host.element_size() + (env_h.nbytes if args.env else 0))},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "fp64_fma" if w.get("dtype") == "f64" else "fp32_fma", "achieved": achieved / 1e12, "peak": peak["slots_per_s"] / 1e12, "unit": "Tslot/s",
                     "frac": achieved / peak["slots_per_s"], "traffic": traffic,
                     "peak_source": "hg_fma_peak FFMA / DFMA microbenchmark, measured in this run (MEASURED_PEAKS.json has no FP32 / FP64 figure)",
                     "frac_of_nominal": achieved / nominal, "slots_per_unit": SLOTS_PER_FILTER_SAMPLE,
                     "kernel_ms": t_kernel / args.steps * 1e3,
                     # bit-exact ceiling: every lane must run the FMUL->FADD->FADD chain once per sample, serially
                     "chain_cycles_per_sample_measured": chain_cycles,
                     "bit_exact_floor_ms": L * chain_cycles / (peak["sm_mhz"] * 1e3) if peak["sm_mhz"] else None, "mode": info["mode"], "filters_per_thread": info["filters_per_thread"],
                     "ctas": info["ctas"]},
    }
    if args.env and args.followers == "ring":
        # the fused envelope followers add 8 B of ring traffic per filter-sample (read q[t-width], write q[t]):
        # that kernel is HBM-bound, so its roofline is the measured copy bandwidth
        try:
            hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            src = "MEASURED_PEAKS.json hbm_gbs (measured)"
        except Exception:
            hbm, src = 6650.0, "B200_PROFILING.md fallback"
        gbs = per_gpu_kernel * 8 / 1e9
        line["roofline"] = {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "traffic": None,
                            "peak_source": src, "bytes_per_unit": 8, "kernel_ms": t_kernel / args.steps * 1e3,
                            "fma_frac": achieved / peak["slots_per_s"], "mode": info["mode"], "ctas": info["ctas"],
                            "note": "Filterbank + per-band RMSbank(2400) followers, ring of y^2 in HBM; envelopes every 64 samples"}
    elif args.env:
        # replica followers: no ring; algorithmic work per filter-sample = the filter (11 slots) + the follower
        # (y*y, q - q_old, + S: 3 slots).  The lagged replica that regenerates q_old is overhead, not counted.
        slots = SLOTS_PER_FILTER_SAMPLE + 3
        line["roofline"].update({"achieved": per_gpu_kernel * slots / 1e12, "frac": per_gpu_kernel * slots / peak["slots_per_s"],
                                 "frac_of_nominal": per_gpu_kernel * slots / nominal, "slots_per_unit": slots,
                                 "note": "Filterbank + per-band RMSbank(2400) followers; y[t-2400] regenerated by lagged replica filters (no ring); envelopes every 64 samples"})
    if parity:
        line["parity"] = parity
    if world == 1 and args.workload == "c2" and not args.no_side and args.mode == "exact" and not args.env:
        try:
            line["side"] = side_measurements(args, local, flush, peak)
        except Exception as e:
            line["side"] = {"failed": str(e).splitlines()[0][:300]}
    if strong:
        line["strong_c3"] = strong
    if exchange:
        line["config"]["exchange"] = ("running bank-sum row handed GPU to GPU through peer memory in filter order (hg_fb_process_sharded, HG_SHARD_CHAIN); result on the last rank"
                                      if shard_mode == "chain" else "ncclReduce of the ranks' partial sums (HG_SHARD_REDUCE)")
    if scan_info:
        line["scan"] = scan_info
    if not args.no_cpu_baseline:
        n_s = 3 * SR if N * 3 * SR <= 4e9 else max(2400, int(4e9 / N) // 2400 * 2400)
        v, dt = cpu_baseline(w, dtype, n_s, 1)
        line["cpu_baseline"] = {"value": v, "unit": "filter-samples/s", "cores": 1, "kind": "port",
                                "sample": f"{N} filters x first {n_s} samples of the same input, strict oracle built like tests/build.sh:8 (g++ -O3, no -march), {dt:.1f} s"}
        try:
            v2, dt2 = cpu_baseline(w, dtype, n_s, 1, native=True)
            line["cpu_baseline"]["march_native"] = {"value": v2, "unit": "filter-samples/s", "cores": 1,
                                                    "sample": f"same loop and sample, g++ -O3 -march=native -ffp-contract=off on this host, {dt2:.1f} s"}
        except Exception as e:          # the second figure is informative only
            line["cpu_baseline"]["march_native"] = {"unavailable": str(e).splitlines()[0][:200]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
