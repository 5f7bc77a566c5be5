cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "filterbank" 2>&1 | tail -4
for a in "4096 10 ordered" "8192 10 ordered" "16384 10 ordered" "65536 10 ordered"; do timeout 120 python tools/chain_prof.py $a; done > gpurun_out/prof3.log 2>&1; cat gpurun_out/prof3.log
for w in c2 c3 c3_8192 c3_16384 c4; do timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-scan --workload $w 2>&1 | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(\"$w\", l[\"roofline\"][\"kernel_ms\"], l[\"roofline\"][\"frac\"])"; done
