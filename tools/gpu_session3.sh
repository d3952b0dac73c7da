#!/bin/bash
# 2 GPUs: multi-device tests, C program, bench --gpus 2 (c3 strong scaling), c4 over 2 ranks
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2s3_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cabi.py -q -s -p no:cacheprovider > gpurun_out/r2s3_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2s3_pytest.log
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2s3_bench_n2.out 2> gpurun_out/r2s3_bench_n2.err; echo "bench n2 rc=$?"
tail -n 1 gpurun_out/r2s3_bench_n2.out | cut -c 1-1500
tail -n 5 gpurun_out/r2s3_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config c4 --steps 5 --warmup 3 > gpurun_out/r2s3_bench_c4_n2.out 2> gpurun_out/r2s3_bench_c4_n2.err; echo "bench c4 n2 rc=$?"
tail -n 1 gpurun_out/r2s3_bench_c4_n2.out | cut -c 1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 0 > gpurun_out/r2s3_ref_n2.out 2>&1; echo "ref n2 rc=$?"
