cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "filterbank" 2>&1 | tail -3
for w in c2 c3_8192 c3_16384; do for sm in ordered tree; do timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-scan --workload $w --sum $sm 2>&1 | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(\"$w $sm\", l[\"roofline\"][\"kernel_ms\"], l[\"roofline\"][\"frac\"])"; done; done
