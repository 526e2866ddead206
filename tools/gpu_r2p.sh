#!/bin/bash
# round-2 GPU call P: conv / wgrad tile sweep (with wgrad), bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/bench_conv.py > gpurun_out/r2p_bench_conv.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers gpurun_out/r2p_layers.txt > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
cat gpurun_out/r2p_bench_conv.txt
python - <<'PY'
import json
for l in open('gpurun_out/r2p_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['conv_roofline']['frac'], d['kernel_ms_per_step'])
PY
