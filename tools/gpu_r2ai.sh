#!/bin/bash
# round-2 GPU call AI: how much of the content loss is exposed in the step?
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 400 python tools/bench_loss_exposure.py > gpurun_out/r2ai_exposure.txt 2> gpurun_out/r2ai_exposure.err
cat gpurun_out/r2ai_exposure.txt; tail -2 gpurun_out/r2ai_exposure.err
