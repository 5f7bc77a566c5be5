#!/bin/bash
# Runs on the GPU box (gpurun): the evidence behind profiles/ — GPU test suite, bench lines, ncu launch list,
# one ncu --set full capture per hot kernel.  Everything lands in gpurun_out/; tools/ncu_summary.py turns the
# .ncu-rep files into the text summaries that are committed.
set -u
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/gpu_tests.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_n1_reference.json 2> $O/bench_ref.err
for w in c1 c3 c3_8192 c3_16384 c4; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-scan --no-side > $O/bench_$w.json 2>/dev/null; done
for w in c2 c3 c3_8192 c3_16384 c4; do timeout 300 python bench.py --workload $w --sum tree --steps 5 --warmup 3 --no-cpu-baseline --no-scan --no-side > $O/bench_${w}_treesum.json 2>/dev/null; done
timeout 600 python tools/scan_error.py --out $O/scan_error.json > $O/scan_error.log 2>&1
timeout 600 tools/hop_bench > $O/hop_bench.txt 2>&1 || true
timeout 600 tools/issue_bench > $O/issue_bench.txt 2>&1 || true
timeout 600 python bench.py --workload c4 --env --steps 5 --warmup 3 --no-cpu-baseline --no-scan > $O/bench_c4_env_replica.json 2>/dev/null
timeout 600 python bench.py --workload c4 --env --followers ring --steps 5 --warmup 3 --no-cpu-baseline --no-scan > $O/bench_c4_env_ring.json 2>/dev/null
timeout 600 python bench.py --workload c5 --steps 3 --warmup 3 > $O/bench_c5.json 2>/dev/null
timeout 600 python tools/bench_banks.py --seconds 10 --chain > $O/bench_banks.json 2>/dev/null
timeout 600 python tools/bench_banks.py --seconds 2 --channels 65536 > $O/bench_banks_65536ch.json 2>/dev/null
# launch list of the headline command (per-launch times are cold-cache and serialised: compare shares, not absolutes)
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side > $O/ncu_launches.log 2>&1
# one full capture per hot kernel
timeout 600 $NCU --set full -k regex:fb_exact_ws -s 3 -c 1 -o $O/full_ws -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-scan --no-side --parity-samples 0 > $O/ncu_ws.log 2>&1
timeout 600 $NCU --set full -k regex:fb_exact_ru -s 3 -c 1 -o $O/full_ru_c3 -f python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline --no-scan --no-side --parity-samples 0 > $O/ncu_ru.log 2>&1
timeout 600 $NCU --set full -k regex:fb_exact_ru -s 3 -c 1 -o $O/full_ru_env_replica -f python bench.py --workload c4 --env --seconds 0.25 --steps 1 --warmup 3 --no-cpu-baseline --no-scan --no-side > $O/ncu_env.log 2>&1
timeout 600 $NCU --set full -k regex:scan_pass -s 2 -c 2 -o $O/full_scan -f python bench.py --mode scan --steps 1 --warmup 3 --no-cpu-baseline --no-side > $O/ncu_scan.log 2>&1
timeout 600 $NCU --set full -k regex:osc_kernel -s 1 -c 1 -o $O/full_osc -f python tools/bench_banks.py --seconds 2 > $O/ncu_osc.log 2>&1
timeout 600 $NCU --set full -k regex:rms_ws -s 1 -c 1 -o $O/full_rms -f python tools/bench_banks.py --seconds 2 > $O/ncu_rms.log 2>&1
ls -la $O | tail -30
cat $O/gpu_tests.log
