#!/bin/bash
# round-2 GPU call O: recalibrated conv tile model, GN occupancy by size, fused mask pyramid: tests + bench with per-layer table
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_conv_tc_gpu.py tests/test_step_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2o_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2o_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers gpurun_out/r2o_layers.txt > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
tail -3 gpurun_out/r2o_pytest.log
python - <<'PY'
import json
for l in open('gpurun_out/r2o_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['conv_roofline']['frac'], d['gn_roofline']['frac'], d['kernel_ms_per_step'])
PY
