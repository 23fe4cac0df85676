#!/bin/bash
mkdir -p gpurun_out
echo "=== clip tests"; timeout 400 python -m pytest tests/test_model_gpu.py -q -m gpu -k "clip_forward" -s --no-header -p no:cacheprovider 2>&1 | grep -vE "^\s*$" | tail -25 | cut -c1-300
echo "=== full gpu test-suite"; timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
