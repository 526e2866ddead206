#!/bin/bash
# round-2 GPU call S: evidence pass on the final kernels: full suite, smoke, bench (with the CPU reference arm), launch list, ncu --set full extracts
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2s_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2s_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --layers gpurun_out/r2s_layers.txt > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2s_bench_reference.json 2> gpurun_out/r2s_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2s_launches.csv \
  python bench.py --ncu-step --warmup 3 > gpurun_out/r2s_ncu0.log 2>&1
for k in warp_forward_tiles warp_backward_tiles mask_pyramid fill4; do
  timeout 200 ncu --set full --clock-control none --profile-from-start off -k regex:$k -c 1 --csv --page raw \
    --log-file gpurun_out/r2s_full_$k.csv python bench.py --ncu-step --warmup 3 > /dev/null 2>&1
done
for k in gn_bwd_reduce conv_tc_persist conv_tc_kernel wgrad_tc; do
  timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:$k -c 6 --csv --page raw \
    --log-file gpurun_out/r2s_full_$k.csv python bench.py --ncu-step --warmup 3 > /dev/null 2>&1
done
grep -E "passed|failed" gpurun_out/r2s_pytest.log | tail -2; tail -1 gpurun_out/r2s_smoke.log; cat gpurun_out/r2s_bench_reference.json | cut -c1-400
ls gpurun_out | grep -c r2s
