#!/bin/bash
# round-2 GPU call C: pipelined warp forward (sweep), GN kernels v2, PDL on the conv kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "not bench_sizes and not other_baseline and not main_py" > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
for pf in 0 2 4 8; do echo "== PTK_WARP_PF=$pf"; PTK_WARP_PF=$pf timeout 300 python tools/bench_warp.py; done > gpurun_out/r2c_warp_sweep.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
PTK_PDL=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_nopdl.json 2>> gpurun_out/r2c_bench.err
timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:'warp_forward_levels|warp_backward_levels' -c 2 \
  --csv --page raw --log-file gpurun_out/r2c_warp_raw.csv python bench.py --ncu-step --warmup 3 > gpurun_out/r2c_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'warp_forward_levels' -c 1 \
  -o gpurun_out/r2c_warp python bench.py --ncu-step --warmup 3 >> gpurun_out/r2c_ncu.log 2>&1
grep -E "passed|failed" gpurun_out/r2c_pytest.log | tail -3
grep -E "==|forward  var=[23] TH=(4|8) |backward TH=8" gpurun_out/r2c_warp_sweep.txt
