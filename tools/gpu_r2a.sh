#!/bin/bash
# round-2 GPU call A: full -m gpu suite (new parity tests), bench + diag, ncu evidence for the memory-bound families
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --diag --layers gpurun_out/r2a_layers.txt > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" >> gpurun_out/r2a_bench.err
timeout 900 ncu --set full --clock-control none --profile-from-start off \
  -k regex:'gn_|adam_kernel|nnloss_|stem_conv|warp_backward|warp_forward|mask_pyramid|sum_parts|transpose_weight|nchw_to_nhwc|head_shift' -c 220 \
  --csv --page raw --log-file gpurun_out/r2a_mem_raw.csv python bench.py --ncu-step --warmup 3 > gpurun_out/r2a_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/r2a_ncu.log
ls -la gpurun_out | tail -8
tail -5 gpurun_out/r2a_pytest.log
