#!/bin/bash
# Subset of tools/collect_profiles.sh for a change that only touches the K1-ru kernel: GPU tests, the default bench
# line (its side block carries the configs[2..4] figures), the K1-ru workloads, and the two K1-ru ncu captures.
set -u
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/gpu_tests.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
for w in c3 c4; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-scan --no-side > $O/bench_$w.json 2>/dev/null; done
for w in c3 c4; do timeout 300 python bench.py --workload $w --sum tree --steps 5 --warmup 3 --no-cpu-baseline --no-scan --no-side > $O/bench_${w}_treesum.json 2>/dev/null; done
timeout 300 python bench.py --workload c4 --env --steps 5 --warmup 3 --no-cpu-baseline --no-scan > $O/bench_c4_env_replica.json 2>/dev/null
timeout 300 python bench.py --workload c4 --env --followers ring --steps 5 --warmup 3 --no-cpu-baseline --no-scan > $O/bench_c4_env_ring.json 2>/dev/null
timeout 300 python bench.py --workload c5 --steps 3 --warmup 3 > $O/bench_c5.json 2>/dev/null
timeout 300 $NCU --set full -k regex:fb_exact_ru -s 3 -c 1 -o $O/full_ru_c3 -f python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline --no-scan --no-side --parity-samples 0 > $O/ncu_ru.log 2>&1
timeout 300 $NCU --set full -k regex:fb_exact_ru -s 3 -c 1 -o $O/full_ru_env_replica -f python bench.py --workload c4 --env --seconds 0.25 --steps 1 --warmup 3 --no-cpu-baseline --no-scan --no-side > $O/ncu_env.log 2>&1
ls -la $O | tail -20
cat $O/gpu_tests.log
