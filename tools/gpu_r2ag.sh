#!/bin/bash
# round-2 GPU call AG: tiny-extent tensor-core cases + step tests after the tolerance note
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_step_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2ag_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2ag_pytest.log
grep -E "passed|failed" gpurun_out/r2ag_pytest.log | tail -1; grep -E "^FAILED" gpurun_out/r2ag_pytest.log | head
