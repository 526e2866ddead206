#!/bin/bash
# round-2 GPU call J: coalescing epilogue of the persistent conv kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_step_gpu.py -m gpu -q -x -p no:cacheprovider -k "not other_baseline" > gpurun_out/r2j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
PTK_STEM=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers gpurun_out/r2j_layers.txt > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
grep -E "passed|failed" gpurun_out/r2j_pytest.log | tail -3
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2j_bench.json").read().strip().splitlines()[-1]); print(round(d["value"],1), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["kernel_ms_per_step"].items()})
PY
head -30 gpurun_out/r2j_layers.txt
