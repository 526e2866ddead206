#!/bin/bash
# round-2 GPU call AJ: NN-loss kernels at two resident CTAs per SM (register cap 72-80, some spills) vs one
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
PTK_NNLOSS_OCC=2 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "nnloss or nn_loss" > gpurun_out/r2aj_pytest.log 2>&1
timeout 300 python tools/bench_nnloss.py > gpurun_out/r2aj_bench_nnloss.txt 2>&1
tail -2 gpurun_out/r2aj_pytest.log; cat gpurun_out/r2aj_bench_nnloss.txt
