#!/bin/bash
# round-2 GPU call V: row-halo stages + phase-fastest tile order: conv / step tests, A/B bench on the same box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_conv_tc_gpu.py tests/test_step_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2v_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2v_pytest.log
PTK_TC_HALO=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers gpurun_out/r2v_layers_nohalo.txt > gpurun_out/r2v_bench_nohalo.json 2> gpurun_out/r2v_bench_nohalo.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers gpurun_out/r2v_layers_halo.txt > gpurun_out/r2v_bench_halo.json 2> gpurun_out/r2v_bench_halo.err
PTK_TC_HALO=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2v_bench_nohalo2.json 2> gpurun_out/r2v_bench_nohalo2.err
PTK_TC_TUNE_FILE=$PWD/gpurun_out/r2v_tune.txt timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2v_bench_halo2.json 2> gpurun_out/r2v_bench_halo2.err
grep -E "passed|failed" gpurun_out/r2v_pytest.log | tail -2; grep -E "^FAILED" gpurun_out/r2v_pytest.log | head
python - <<'PY'
import json
for f in ("nohalo","halo","nohalo2","halo2"):
    try:
        for l in open('gpurun_out/r2v_bench_%s.json'%f):
            if l.startswith('{'):
                d=json.loads(l); print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'conv', round(d['conv_roofline']['frac'],3), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if 'conv' in k})
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/r2v_bench_halo.err
