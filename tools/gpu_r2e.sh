#!/bin/bash
# round-2 GPU call E: warp forward v4 (balanced tiles, L2 prefetch), interleaved encoder backward / finer buckets
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "not other_baseline and not main_py and not bench_sizes" > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
for pf in 1 0; do echo "== PTK_WARP_PF=$pf"; PTK_WARP_PF=$pf timeout 300 python tools/bench_warp.py; done > gpurun_out/r2e_warp_sweep.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:'warp_forward_tiles' -c 1 \
  --csv --page raw --log-file gpurun_out/r2e_warp_raw.csv python bench.py --ncu-step --warmup 3 > gpurun_out/r2e_ncu.log 2>&1
grep -E "passed|failed" gpurun_out/r2e_pytest.log | tail -3
cat gpurun_out/r2e_warp_sweep.txt
