#!/bin/bash
# round-2 session 1: tests + smoke + default bench on one GPU
mkdir -p gpurun_out
{ free -g | head -2; nproc; nvidia-smi -L; } > gpurun_out/r2s1_box.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2s1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s1_pytest.log
tail -5 gpurun_out/r2s1_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2s1_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s1_bench.json 2> gpurun_out/r2s1_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2s1_bench.json
