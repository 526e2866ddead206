#!/bin/bash
# round-2 GPU call AA: one-pass input row assembly (ptk_gather_nhwc): full suite + A/B-free bench + launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2aa_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2aa_pytest.log
export PTK_TC_TUNE_FILE=$PWD/gpurun_out/r2aa_tune.txt
rm -f $PTK_TC_TUNE_FILE
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r2aa_dram.csv python bench.py --ncu-step --warmup 3 > /dev/null 2>&1
grep -E "passed|failed" gpurun_out/r2aa_pytest.log | tail -1; grep -E "^FAILED" gpurun_out/r2aa_pytest.log | head
python - <<'PY'
import json
for l in open('gpurun_out/r2aa_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'dev', round(d['e2e_device_data_path']['value'],1), d['clocks'])
PY
grep -c gather_nhwc gpurun_out/r2aa_dram.csv
