"""Copies what tools/collect_profiles.sh left in gpurun_out/ into profiles/ under this round's names and writes the
derived summaries (ncu text summaries, launch-list shares).  Run on the CPU box after the gpurun call.

    python tools/install_profiles.py r2
"""
import collections
import csv
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    plain = {"gpu_tests.log": f"{tag}_gpu_tests.log", "bench_n1.json": f"{tag}_bench_n1.json", "bench_n1_reference.json": f"{tag}_bench_n1_reference.json",
             "scan_error.json": f"scan_error_{tag}.json", "hop_bench.txt": f"{tag}_hop_bench.txt", "issue_bench.txt": f"{tag}_issue_bench.txt",
             "launches_bench.csv": f"{tag}_launches_bench.csv", "bench_banks.json": f"{tag}_bench_banks.json",
             "bench_banks_65536ch.json": f"{tag}_bench_banks_65536ch.json"}
    for w in ("c1", "c3", "c3_8192", "c3_16384", "c4", "c5", "c4_env_replica", "c4_env_ring"):
        plain[f"bench_{w}.json"] = f"{tag}_bench_{w}.json"
    for w in ("c2", "c3", "c3_8192", "c3_16384", "c4"):
        plain[f"bench_{w}_treesum.json"] = f"{tag}_bench_{w}_treesum.json"
    for src, dst in plain.items():
        a = os.path.join(O, src)
        if os.path.exists(a) and os.path.getsize(a) > 0:
            shutil.copyfile(a, os.path.join(P, dst))
        else:
            print("missing:", src)
    for rep, dst in (("full_ws", "ncu_ws"), ("full_ru_c3", "ncu_ru_c3"), ("full_ru_env_replica", "ncu_ru_env_replica"), ("full_scan", "ncu_scan"),
                     ("full_osc", "ncu_osc"), ("full_rms", "ncu_rms")):
        a = os.path.join(O, rep + ".ncu-rep")
        if os.path.exists(a):
            subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), a, os.path.join(P, f"{tag}_{dst}.txt")], stdout=subprocess.DEVNULL)
        else:
            print("missing:", rep)
    # launch list: GPU time by kernel
    a = os.path.join(O, "launches_bench.csv")
    if os.path.exists(a):
        rows = [r for r in csv.reader(l for l in open(a) if not l.startswith("=="))]
        hdr = rows[0]
        ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        tot, cnt = collections.Counter(), collections.Counter()
        for r in rows[1:]:
            if len(r) <= iv:
                continue
            v = float(r[iv].replace(",", ""))
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
            name = r[ik].split("(")[0][:70]
            tot[name] += v
            cnt[name] += 1
        allms = sum(tot.values())
        with open(os.path.join(P, f"{tag}_launches_bench_summary.txt"), "w") as f:
            f.write(f"ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side` (profiles/{tag}_launches_bench.csv): GPU time by kernel\n")
            f.write("(per-launch times are cold-cache and serialised under the profiler: compare shares, not absolutes)\n\n")
            for k, v in tot.most_common():
                f.write(f"{v:10.3f} ms  {100 * v / allms:5.1f} %  x{cnt[k]:<3d}  {k}\n")


if __name__ == "__main__":
    main()
