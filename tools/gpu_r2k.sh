#!/bin/bash
# round-2 GPU call K: full suite + bench + launch list + ncu extracts of the memory-bound families (few launches each)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --layers gpurun_out/r2k_layers.txt > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2k_launches.csv \
  python bench.py --ncu-step --warmup 3 > gpurun_out/r2k_ncu0.log 2>&1
for k in warp_forward_tiles warp_backward_levels adam_kernel nnloss_forward nnloss_backward splitk_reduce pose_masks pose_heatmaps mask_pyramid; do
  timeout 200 ncu --set full --clock-control none --profile-from-start off -k regex:$k -c 1 --csv --page raw \
    --log-file gpurun_out/r2k_full_$k.csv python bench.py --ncu-step --warmup 3 > /dev/null 2>&1
done
for k in gn_apply gn_bwd_reduce gn_bwd_apply conv_tc_persist conv_tc_kernel wgrad_tc; do
  timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:$k -c 6 --csv --page raw \
    --log-file gpurun_out/r2k_full_$k.csv python bench.py --ncu-step --warmup 3 > /dev/null 2>&1
done
grep -E "passed|failed" gpurun_out/r2k_pytest.log | tail -3
ls -la gpurun_out | grep r2k | wc -l
