#!/bin/bash
# round-2 GPU call B: new warp kernels (tests + sweep), stacked training, re-stated parity tolerances, bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
timeout 300 python tools/bench_warp.py > gpurun_out/r2b_warp_sweep.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench rc=$?" >> gpurun_out/r2b_bench.err
PTK_WARP_VAR=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_var1.json 2>> gpurun_out/r2b_bench.err
grep -E "passed|failed" gpurun_out/r2b_pytest.log | tail -3
cat gpurun_out/r2b_warp_sweep.txt
