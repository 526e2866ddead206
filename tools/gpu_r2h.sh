#!/bin/bash
# round-2 GPU call H: persistent double-buffered conv kernel (tests, bench with / without), pose-data fix
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_pose_data_gpu.py tests/test_step_gpu.py -m gpu -q -x -p no:cacheprovider -k "not other_baseline" > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers gpurun_out/r2h_layers.txt > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
PTK_TC_PERSIST=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers gpurun_out/r2h_layers_nops.txt > gpurun_out/r2h_bench_nops.json 2>> gpurun_out/r2h_bench.err
grep -E "passed|failed" gpurun_out/r2h_pytest.log | tail -3
python - <<'PY'
import json
for f in ("gpurun_out/r2h_bench.json","gpurun_out/r2h_bench_nops.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["kernel_ms_per_step"]["conv_forward"])
    except Exception as e: print(f, "ERR", e)
PY
