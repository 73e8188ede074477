#!/usr/bin/env bash
# Runs the GPU test files one process at a time, most basic first, and stops at the first file that hits a
# CUDA fault (a sticky context error would fail everything after it anyway, and faults are rationed).
# usage: scripts/gpu_tests.sh [file ...]   (logs under gpurun_out/)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
files=("$@")
[ ${#files[@]} -eq 0 ] && files=(tests/test_gpu_kernels.py tests/test_gpu_preprocess.py tests/test_gpu_disagreement.py tests/test_gpu_model.py tests/test_gpu_episode.py tests/test_gpu_textenc.py)
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu_info.txt 2>&1
rc_all=0
for f in "${files[@]}"; do
  log=gpurun_out/$(basename "$f" .py).log
  timeout ${EC_TEST_TIMEOUT:-900} python -m pytest "$f" -m gpu -q -s --tb=short -p no:cacheprovider > "$log" 2>&1
  rc=$?
  echo "== $f rc=$rc"; tail -n 40 "$log"
  [ $rc -ne 0 ] && rc_all=1
  if grep -qiE "illegal memory access|unspecified launch failure|illegal instruction|misaligned address|device-side assert|EC_ERR_TIMEOUT" "$log"; then
    echo "!! CUDA fault or pipeline timeout in $f -- stopping"; exit 3
  fi
done
exit $rc_all
