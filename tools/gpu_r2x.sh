#!/bin/bash
# round-2 GPU call X: phase-fastest tile order alone (tests), bench, and a per-launch DRAM-traffic list of one step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PTK_TC_TUNE_FILE=$PWD/gpurun_out/r2x_tune.txt
rm -f $PTK_TC_TUNE_FILE
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_step_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r2x_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2x_pytest.log
rm -f $PTK_TC_TUNE_FILE
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers gpurun_out/r2x_layers.txt > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r2x_dram.csv python bench.py --ncu-step --warmup 3 > gpurun_out/r2x_ncu0.log 2>&1
grep -E "passed|failed" gpurun_out/r2x_pytest.log | tail -1
python - <<'PY'
import json
for l in open('gpurun_out/r2x_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'conv', round(d['conv_roofline']['frac'],3), d['clocks'])
PY
