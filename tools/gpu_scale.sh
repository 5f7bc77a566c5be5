# usage: bash tools/gpu_scale.sh N   (run under gpurun --gpus N)
cd $GRAFT_REPO_ROOT
N=$1
nvidia-smi -L | head -8
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_sharded_gpu.py tests/test_harness_gpu.py -x -q -m gpu -k "sharded" 2>&1 | tail -4; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_n$N.log 2>&1
tail -1 gpurun_out/scale_n$N.log > gpurun_out/r2_bench_n$N.json
python - <<PY
import json
l=json.load(open("gpurun_out/r2_bench_n$N.json"))
print("N=$N value",l["value"],"ms",l["ms_per_step"],"e2e",l["e2e"]["value"],"parity",l["parity"]["bit_identical"])
s=l["strong_c3"]
print("strong one_gpu_ms",s["one_gpu_ms_per_block"])
for m in ("chain","reduce"): print(m, s[m]["ms_per_block"], "eff", s[m]["efficiency"], "kernel", s[m]["kernel"], "bit", s[m]["parity"]["bit_identical"], "err", s[m]["parity"]["max_abs_err_rel_peak"])
PY
