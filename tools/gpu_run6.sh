cd $GRAFT_REPO_ROOT
run() { timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-scan --workload $1 --sum $2 2>&1 | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(\"$3 $1 $2\", l[\"roofline\"][\"kernel_ms\"], l[\"roofline\"][\"frac\"])"; }
for v in "" NO_TICKET NO_LEAN BOTH; do
  if [ -n "$v" ]; then export HG_LIB_PATH=$GRAFT_REPO_ROOT/huygens_b200/lib/exp/lib_$v.so; else unset HG_LIB_PATH; fi
  run c2 ordered "[$v]"; run c2 tree "[$v]"
done
unset HG_LIB_PATH
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "filterbank" 2>&1 | tail -3
for w in c3 c4; do run $w ordered "[]"; run $w tree "[]"; done
