#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2s5_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2s5_pytest.log
for k in 1 2 3 4 6; do DM_STREAM_CHUNKS=$k timeout 300 python tools/time_strip_e2e.py c3 8; done 2>&1 | tee gpurun_out/r2s5_strip.log
for k in 1 3; do DM_STREAM_CHUNKS=$k timeout 300 python tools/time_strip_e2e.py c2 1; done 2>&1 | tee -a gpurun_out/r2s5_strip.log
