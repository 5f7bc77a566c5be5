#!/usr/bin/env python
"""Measured error of the CUDA paths against the oracle on the bench workload (BASELINE configs[1]):
exact mode (must be bit-exact per filter) and scan mode (tolerance-bounded), for the full-range bank
and for a 500 Hz..20 kHz bank.  Test/measurement tool: uses the oracle as the checker.

    python tools/scan_error.py [--seconds 60] [--out profiles/scan_error_r2.json]
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
import huygens_b200 as hb  # noqa: E402
from huygens_b200 import design  # noqa: E402
from conftest import rel_err, snr_db  # noqa: E402


def oracle_sum(fwd, back, gains, x, dtype, k_p, k_g, threads):
    """the oracle's own sequential index-order sum of the whole bank (oracle.chained_sum: consecutive filter
    partitions on the host threads, each continuing the running row of the one before — bit-identical to one
    oracle.Filterbank of all the filters)"""
    N = len(gains)
    banks = []
    for idx in np.array_split(np.arange(N), min(threads, N)):
        fb = oracle.Filterbank(2, len(idx), k_p, k_g, dtype=dtype, mode="strict")
        for j, n in enumerate(idx):
            fb.coefficients(j, fwd[n], back[n])
        fb.mix(gains[idx])
        fb.boost(np.ones(len(idx)))
        banks.append(fb)
    return oracle.chained_sum(banks, x)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--filters", type=int, default=4096)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    oracle.build()
    L = int(args.seconds * 48000)
    N = args.filters
    threads = os.cpu_count() or 1
    res = []
    for dtype in (np.float32, np.float64):
        x = design.chirp(L, dtype=dtype)
        for f_lo in (27.5, 500.0):
            freqs = design.log_spaced(N, f_lo, 20000.0)
            fwd, back, gains = design.resonator_bank(freqs, 0.999)
            t0 = time.time()
            ref = oracle_sum(fwd, back, gains, x, dtype, 0.001, 1.0, threads)
            t_ref = time.time() - t0
            row = {"dtype": np.dtype(dtype).name, "filters": N, "samples": L, "f_lo": f_lo, "r": 0.999,
                   "oracle": "strict, the reference's sequential index-order sum in T (chained over %d host threads)" % threads, "oracle_s": round(t_ref, 1),
                   "peak": float(np.max(np.abs(ref)))}
            for mode, summ in (("exact", "ordered"), ("exact", "tree"), ("scan", "ordered")):
                fb = hb.Filterbank(2, N, 0.001, 1.0, dtype=dtype, sum_mode=summ)
                fb.coefficients(0, fwd, back)
                fb.mix(gains)
                fb.boost(np.ones(N))
                d = fb.process(x, mode=mode)
                key = mode if summ == "ordered" else f"{mode}_{summ}_sum"
                row[key] = {"max_abs_err_rel_peak": rel_err(d, ref), "snr_db": snr_db(d, ref), "bit_identical": bool(np.array_equal(d, ref))}
            res.append(row)
            print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
