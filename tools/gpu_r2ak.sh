#!/bin/bash
# round-2 GPU call AK: NN-loss kernels at two CTAs per SM: kernel + step + bench-size tests, bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_step_gpu.py tests/test_bench_sizes_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2ak_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2ak_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2ak_bench.json 2> gpurun_out/r2ak_bench.err
timeout 200 python tools/bench_nnloss.py > gpurun_out/r2ak_bench_nnloss.txt 2>&1
grep -E "passed|failed" gpurun_out/r2ak_pytest.log | tail -1; grep -E "^FAILED" gpurun_out/r2ak_pytest.log | head
python - <<'PY'
import json
for l in open('gpurun_out/r2ak_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'dev', round(d['e2e_device_data_path']['value'],1), d['clocks'], d['kernel_ms_per_step']['nnloss'])
PY
cat gpurun_out/r2ak_bench_nnloss.txt
