cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "filterbank" 2>&1 | tail -4
for a in "4096 10 ordered" "4096 10 tree" "65536 10 ordered" "65536 10 tree"; do timeout 120 python tools/chain_prof.py $a; done > gpurun_out/prof4.log 2>&1; cat gpurun_out/prof4.log
for w in c2 c3 c4; do for sm in ordered tree; do timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-scan --workload $w --sum $sm 2>&1 | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(\"$w $sm\", l[\"roofline\"][\"kernel_ms\"], l[\"roofline\"][\"frac\"])"; done; done
