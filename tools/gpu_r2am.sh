#!/bin/bash
# round-2 GPU call AM: full -m gpu suite + smoke at the final HEAD
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2am_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2am_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2am_smoke.log 2>&1
grep -E "passed|failed" gpurun_out/r2am_pytest.log | tail -1; grep -E "^FAILED" gpurun_out/r2am_pytest.log | head; tail -1 gpurun_out/r2am_smoke.log
