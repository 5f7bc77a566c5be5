set -x
cd $GRAFT_REPO_ROOT
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-scan"
for w in c2 c3 c3_8192 c3_16384 c4; do
  for sm in ordered tree; do
    $B --workload $w --sum $sm > gpurun_out/b2_${w}_${sm}.log 2>&1
    python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/b2_${w}_${sm}.log").read().strip().splitlines()[-1])
    print("$w $sm kernel_ms", l["roofline"]["kernel_ms"], "frac", l["roofline"]["frac"])
except Exception as e:
    print("$w $sm FAILED", e)
PY
  done
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fb_exact_ws -c 1 -o gpurun_out/ws_c2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-scan > gpurun_out/ncu_ws.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fb_exact_ru -c 1 -o gpurun_out/ru_c3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-scan --workload c3 > gpurun_out/ncu_ru.log 2>&1
ls -la gpurun_out
