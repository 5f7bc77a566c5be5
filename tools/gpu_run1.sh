set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_golden.py -x -q -m gpu -k "filterbank or full_size or golden" > gpurun_out/t1.log 2>&1; echo "rc=$?" >> gpurun_out/t1.log
tail -15 gpurun_out/t1.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_c2.log 2>&1; tail -2 gpurun_out/b_c2.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-scan --workload c3 > gpurun_out/b_c3.log 2>&1; tail -2 gpurun_out/b_c3.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-scan --workload c3_8192 > gpurun_out/b_c3_8192.log 2>&1; tail -2 gpurun_out/b_c3_8192.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-scan --workload c3_16384 > gpurun_out/b_c3_16384.log 2>&1; tail -2 gpurun_out/b_c3_16384.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-scan --workload c4 > gpurun_out/b_c4.log 2>&1; tail -2 gpurun_out/b_c4.log
