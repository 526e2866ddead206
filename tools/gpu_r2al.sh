#!/bin/bash
# round-2 GPU call AL: evidence refresh after the NN-loss change (bench line, launch list, DRAM list, nnloss captures)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PTK_TC_TUNE_FILE=$PWD/gpurun_out/r2al_tune.txt
rm -f $PTK_TC_TUNE_FILE
timeout 900 python bench.py --steps 20 --warmup 5 --layers gpurun_out/r2al_layers.txt > gpurun_out/r2al_bench.json 2> gpurun_out/r2al_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2al_launches.csv \
  python bench.py --ncu-step --warmup 3 > gpurun_out/r2al_ncu0.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r2al_dram.csv python bench.py --ncu-step --warmup 3 > /dev/null 2>&1
for k in nnloss_forward nnloss_backward; do
  timeout 200 ncu --set full --clock-control none --profile-from-start off -k regex:$k -c 1 --csv --page raw \
    --log-file gpurun_out/r2al_full_$k.csv python bench.py --ncu-step --warmup 3 > /dev/null 2>&1
done
python - <<'PY'
import json
for l in open('gpurun_out/r2al_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'dev', round(d['e2e_device_data_path']['value'],1), d['clocks'])
PY
ls gpurun_out | grep -c r2al
