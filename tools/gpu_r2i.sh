#!/bin/bash
# round-2 GPU call I: persistence threshold sweep, stem through the generic persistent kernel, full suite, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
for cfg in "1 1" "5 1" "1 0"; do set -- $cfg; PTK_TC_PERSIST=$1 PTK_STEM=$2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers gpurun_out/r2i_layers_p$1_s$2.txt > gpurun_out/r2i_bench_p$1_s$2.json 2>> gpurun_out/r2i_bench.err; done
grep -E "passed|failed" gpurun_out/r2i_pytest.log | tail -3
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2i_bench_p*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],2), round(d["kernel_ms_per_step"]["conv_forward"],2))
    except Exception as e: print(f, "ERR", e)
PY
