#!/bin/bash
# round-2 GPU call G (2 GPUs): 2-rank NCCL test, pose-data tests, bench at N=1 and N=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2g_smi.txt 2>&1
timeout 900 python -m pytest tests/test_ddp_gpu.py tests/test_pose_data_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2g_bench_n2.json 2>> gpurun_out/r2g_bench.err
grep -E "passed|failed|rel-L2" gpurun_out/r2g_pytest.log | tail -5
tail -2 gpurun_out/r2g_bench.err
