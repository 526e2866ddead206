#!/bin/bash
# round-2 GPU call F: warp forward PX=2 at C=64, device data path (tests + e2e), full suite incl. bench-size parity
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
timeout 300 python tools/bench_warp.py > gpurun_out/r2f_warp_sweep.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers gpurun_out/r2f_layers.txt > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
grep -E "passed|failed" gpurun_out/r2f_pytest.log | tail -3
grep -E "var=2|level" gpurun_out/r2f_warp_sweep.txt
tail -3 gpurun_out/r2f_bench.err
