#!/bin/bash
# round-2 GPU call AF (2 GPUs): does capping NCCL's channel count (= SMs its kernels occupy) help the overlapped all-reduces?
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for ch in default 2 4 8; do
  if [ "$ch" = "default" ]; then unset NCCL_MAX_NCHANNELS; else export NCCL_MAX_NCHANNELS=$ch; fi
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2af_bench_ch$ch.json 2> gpurun_out/r2af_bench_ch$ch.err
done
python - <<'PY'
import json
for ch in ("default","2","4","8"):
    for l in open('gpurun_out/r2af_bench_ch%s.json'%ch):
        if l.startswith('{'):
            d=json.loads(l); print("NCCL_MAX_NCHANNELS", ch, round(d['value'],1), round(d['ms_per_step'],3))
PY
