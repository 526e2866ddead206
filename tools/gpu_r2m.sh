#!/bin/bash
# round-2 GPU call M: per-warp geometry phase + record-staged backward tiles (tests, micro-bench), VGG-prefix tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "warp" > gpurun_out/r2m_pytest_warp.log 2>&1
echo "rc=$?" >> gpurun_out/r2m_pytest_warp.log
PTK_WARP_BWD=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "warp" > gpurun_out/r2m_pytest_warp_bwd1.log 2>&1
timeout 300 python tools/bench_warp.py > gpurun_out/r2m_bench_warp.txt 2>&1
timeout 600 python -m pytest tests/test_modules_gpu.py tests/test_step_gpu.py -m gpu -q -p no:cacheprovider \
  -k "feature_extractor or vgg_prefix or deeper" > gpurun_out/r2m_pytest_vgg.log 2>&1
echo "rc=$?" >> gpurun_out/r2m_pytest_vgg.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
tail -3 gpurun_out/r2m_pytest_warp.log; tail -2 gpurun_out/r2m_pytest_warp_bwd1.log; cat gpurun_out/r2m_bench_warp.txt; tail -4 gpurun_out/r2m_pytest_vgg.log
python - <<'PY'
import json
for l in open('gpurun_out/r2m_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['warp_backward_roofline']['frac'], d['kernel_ms_per_step'])
PY
