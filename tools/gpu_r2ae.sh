#!/bin/bash
# round-2 GPU call AE: tensor-map cache + tensor-core path for the 1x1 / 2x2 bottleneck extents: full suite, bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2ae_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2ae_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --diag > gpurun_out/r2ae_bench.json 2> gpurun_out/r2ae_bench.err
grep -E "passed|failed" gpurun_out/r2ae_pytest.log | tail -1; grep -E "^FAILED" gpurun_out/r2ae_pytest.log | head
python - <<'PY'
import json
for l in open('gpurun_out/r2ae_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'dev', round(d['e2e_device_data_path']['value'],1), d['clocks'])
PY
grep -i "host\|enqueue" gpurun_out/r2ae_bench.err | head -5
