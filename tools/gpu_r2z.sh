#!/bin/bash
# round-2 GPU call Z: compute-sanitizer passes over the final kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2z_memcheck.log \
  python -m pytest tests/test_kernels_gpu.py tests/test_conv_tc_gpu.py tests/test_pose_data_gpu.py -m gpu -q -x -p no:cacheprovider \
  > gpurun_out/r2z_memcheck_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2z_memcheck_pytest.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/r2z_racecheck.log \
  python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "warp or gn or nnloss or pyramid or adam" \
  > gpurun_out/r2z_racecheck_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2z_racecheck_pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2z_memcheck_step.log \
  python -m pytest tests/test_step_gpu.py tests/test_modules_gpu.py -m gpu -q -x -p no:cacheprovider -k "(train_step_nn_loss and auto) or deeper_content_layer or vgg_prefix or stacked_training" \
  > gpurun_out/r2z_memcheck_step_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2z_memcheck_step_pytest.log
tail -2 gpurun_out/r2z_memcheck_pytest.log; tail -2 gpurun_out/r2z_memcheck.log
tail -2 gpurun_out/r2z_racecheck_pytest.log; tail -2 gpurun_out/r2z_racecheck.log; tail -2 gpurun_out/r2z_memcheck_step_pytest.log; tail -2 gpurun_out/r2z_memcheck_step.log
