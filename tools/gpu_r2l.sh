#!/bin/bash
# round-2 GPU call L: VGG-prefix tests + compute-sanitizer passes over the kernel-level tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_modules_gpu.py tests/test_step_gpu.py -m gpu -q -p no:cacheprovider \
  -k "feature_extractor or vgg_prefix or deeper" > gpurun_out/r2l_pytest_vgg.log 2>&1
echo "rc=$?" >> gpurun_out/r2l_pytest_vgg.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2l_memcheck.log \
  python -m pytest tests/test_kernels_gpu.py tests/test_conv_tc_gpu.py tests/test_pose_data_gpu.py -m gpu -q -x -p no:cacheprovider \
  > gpurun_out/r2l_memcheck_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2l_memcheck_pytest.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/r2l_racecheck.log \
  python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "warp or gn or nnloss or pyramid or adam" \
  > gpurun_out/r2l_racecheck_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2l_racecheck_pytest.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2l_memcheck_step.log \
  python -m pytest tests/test_step_gpu.py -m gpu -q -x -p no:cacheprovider -k "train_step_nn_loss and auto" \
  > gpurun_out/r2l_memcheck_step_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2l_memcheck_step_pytest.log
tail -3 gpurun_out/r2l_pytest_vgg.log; tail -2 gpurun_out/r2l_memcheck_pytest.log; tail -3 gpurun_out/r2l_memcheck.log
tail -2 gpurun_out/r2l_racecheck_pytest.log; tail -3 gpurun_out/r2l_racecheck.log; tail -2 gpurun_out/r2l_memcheck_step_pytest.log; tail -3 gpurun_out/r2l_memcheck_step.log
