#!/usr/bin/env python
"""Side benchmarks of the Oscbank / RMSbank kernels (BASELINE configs[4]: 65,536 partials).
    python tools/bench_banks.py [--partials 65536] [--seconds 10]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import huygens_b200 as hb  # noqa: E402
from huygens_b200 import design  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--partials", type=int, default=65536)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--channels", type=int, default=1024)
    ap.add_argument("--chain", action="store_true", help="also time the analysis -> resynthesis chain (BASELINE configs[4])")
    args = ap.parse_args()
    L = int(args.seconds * 48000)
    peak = hb.fma_peak(np.float32)["slots_per_s"]
    res = {}
    for dtype in (np.float32, np.float64):
        N = args.partials
        osc = hb.Oscbank(N, dtype=dtype)
        osc.set_frequencies(design.log_spaced(N, 20.0, 20000.0))
        osc.open()
        amps = torch.full((N,), 1.0 / N, dtype=torch.float32 if dtype == np.float32 else torch.float64, device="cuda")
        t = timed(lambda: osc.process(L, amps=amps, want_mix=True, device_out=True))
        # 9 FP32 slots + 2 IEEE divisions per partial-sample (SURVEY.md §8 d: divisions counted as 1 slot each)
        res[f"oscbank_{np.dtype(dtype).name}"] = {"partials": N, "samples": L, "s": t, "partial_samples_per_s": N * L / t,
                                                  "roofline_frac_11slots": N * L / t * 11 / peak}
        C = args.channels
        Lr = min(L, 480000)
        z = torch.rand((C, Lr), dtype=amps.dtype, device="cuda")
        rms = hb.RMSbank(C, 2400, dtype=dtype)
        t = timed(lambda: rms.process(z))
        res[f"rmsbank_{np.dtype(dtype).name}"] = {"channels": C, "samples": Lr, "s": t, "channel_samples_per_s": C * Lr / t,
                                                  "hbm_GBps_algorithmic": C * Lr * 2 * z.element_size() / t / 1e9}
    if args.chain:
        # BASELINE configs[4]: N partials driven by Filterbank/RMSbank analysis of the same signal —
        # Filterbank(N) + fused per-band followers (hop 64) -> envelope rows -> Oscbank(N) amplitudes
        N, hop = args.partials, 64
        freqs = design.log_spaced(N, 27.5, 20000.0)
        fwd, back, gains = design.resonator_bank(freqs, 0.999)
        fb = hb.Filterbank(2, N, k_p=0.001, k_g=1.0, dtype=np.float32)
        fb.coefficients(0, fwd, back)
        fb.boost(np.ones(N))
        fb.mix(gains / N)
        osc = hb.Oscbank(N, dtype=np.float32)
        osc.set_frequencies(freqs)
        osc.open()
        x = torch.from_numpy(design.chirp(L, dtype=np.float32)).cuda()
        out = torch.empty_like(x)
        env = torch.empty((1, L // hop, N), dtype=torch.float32, device="cuda")
        t_an = timed(lambda: fb.process_env(x, width=2400, hop=hop, out=out, env=env, followers="replica"))
        rows = env[0]
        t_sy = timed(lambda: osc.process(rows.shape[0] * hop, amps=rows, hop=hop, want_mix=False, device_out=True))
        res["resynthesis_chain_float32"] = {"partials": N, "samples": L, "hop": hop, "analysis_s": t_an, "synthesis_s": t_sy,
                                            "partial_samples_per_s": N * L / (t_an + t_sy),
                                            "followers": "replica"}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
