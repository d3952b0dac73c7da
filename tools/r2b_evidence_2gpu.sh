#!/bin/bash
# round 2, second session, final code on two GPUs: the one-process multi-device tests and the strong-scaling line of c3
mkdir -p gpurun_out
P=gpurun_out/r2z
timeout -k 10 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cabi.py -q -p no:cacheprovider > ${P}_pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -3 ${P}_pytest_2gpu.log
NCCL_DEBUG=INFO timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > ${P}_c3_n2.out 2> ${P}_c3_n2.err
echo "c3 n2 rc=$?"; tail -n 1 ${P}_c3_n2.out | cut -c 1-400
