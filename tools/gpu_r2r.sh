#!/bin/bash
# round-2 GPU call R: wider autotune candidate space: conv / step tests, A/B bench on the same box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_conv_tc_gpu.py tests/test_step_gpu.py tests/test_bench_sizes_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2r_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2r_pytest.log
PTK_TC_AUTOTUNE=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2r_bench_model.json 2> gpurun_out/r2r_bench_model.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers gpurun_out/r2r_layers_tuned.txt > gpurun_out/r2r_bench_tuned.json 2> gpurun_out/r2r_bench_tuned.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2r_bench_tuned2.json 2> gpurun_out/r2r_bench_tuned2.err
tail -3 gpurun_out/r2r_pytest.log
python - <<'PY'
import json
for f in ("model","tuned","tuned2"):
    for l in open('gpurun_out/r2r_bench_%s.json'%f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'conv', round(d['conv_roofline']['frac'],3), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()})
PY
