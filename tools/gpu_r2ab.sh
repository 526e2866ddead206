#!/bin/bash
# round-2 GPU call AB: vectorised ptk_gather_nhwc: kernel + step tests, bench, per-launch DRAM list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_step_gpu.py tests/test_modules_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2ab_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2ab_pytest.log
export PTK_TC_TUNE_FILE=$PWD/gpurun_out/r2ab_tune.txt
rm -f $PTK_TC_TUNE_FILE
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2ab_bench.json 2> gpurun_out/r2ab_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r2ab_dram.csv python bench.py --ncu-step --warmup 3 > /dev/null 2>&1
grep -E "passed|failed" gpurun_out/r2ab_pytest.log | tail -1; grep -E "^FAILED" gpurun_out/r2ab_pytest.log | head
python - <<'PY'
import json
for l in open('gpurun_out/r2ab_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'dev', round(d['e2e_device_data_path']['value'],1), d['clocks'])
PY
grep gather_nhwc gpurun_out/r2ab_dram.csv | grep time_duration | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '
