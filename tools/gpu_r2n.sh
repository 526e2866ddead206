#!/bin/bash
# round-2 GPU call N: micro-benchmarks (norm family occupancy variants, conv tile shapes), warp tests after the clean-up
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2n_pytest_kernels.log 2>&1
echo "rc=$?" >> gpurun_out/r2n_pytest_kernels.log
timeout 300 python tools/bench_gn.py > gpurun_out/r2n_bench_gn.txt 2>&1
timeout 300 python tools/bench_conv.py > gpurun_out/r2n_bench_conv.txt 2>&1
timeout 200 python tools/bench_warp.py > gpurun_out/r2n_bench_warp.txt 2>&1
tail -2 gpurun_out/r2n_pytest_kernels.log; cat gpurun_out/r2n_bench_gn.txt gpurun_out/r2n_bench_conv.txt; head -5 gpurun_out/r2n_bench_warp.txt
