#!/bin/bash
# round-2 GPU call T (8 GPUs): weak-scaling bench N = 1, 2, 4, 8 (torchrun, one rank per GPU) + the 2-rank NCCL tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2t_gpus.txt
timeout 600 python -m pytest tests/test_ddp_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2t_pytest_ddp.log 2>&1
echo "rc=$?" >> gpurun_out/r2t_pytest_ddp.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2t_bench_n1.json 2> gpurun_out/r2t_bench_n1.err
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
    bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2t_bench_n$n.json 2> gpurun_out/r2t_bench_n$n.err
done
tail -2 gpurun_out/r2t_pytest_ddp.log
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    for l in open('gpurun_out/r2t_bench_n%d.json'%n):
        if l.startswith('{'):
            d=json.loads(l)
            if n==1: base=d['value']
            print(n, round(d['value'],1), round(d['ms_per_step'],3), 'eff', round(d['value']/(n*base),3), 'e2e', round(d['e2e']['value'],1), d.get('e2e_device_data_path',{}).get('value'))
PY
