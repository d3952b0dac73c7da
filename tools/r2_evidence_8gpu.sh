#!/bin/bash
# 8 GPUs, final code: one-process multi-device tests, strong scaling of c3 at 8 / 4 / 2, c5 and c4 at 8
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cabi.py -q -p no:cacheprovider > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2g_pytest.log
run() { # name nproc args...
  local name=$1 np=$2; shift 2
  NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $np "$@" > gpurun_out/r2g_$name.out 2> gpurun_out/r2g_$name.err
  echo "$name rc=$?"; tail -n 1 gpurun_out/r2g_$name.out | cut -c 1-250
}
run c3_n8 8 --steps 20 --warmup 5
run c3_n4 4 --steps 20 --warmup 5 --no-cpu
run c3_n2 2 --steps 20 --warmup 5 --no-cpu
run c5_n8 8 --config c5 --steps 5 --warmup 3 --no-single
run c4_n8 8 --config c4 --steps 10 --warmup 3
