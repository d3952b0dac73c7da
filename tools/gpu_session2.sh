#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2s2_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2s2_pytest.log
for c in c3 c4 c5; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2s2_bench_$c.json 2> gpurun_out/r2s2_bench_$c.err; echo "bench $c rc=$?"
  tail -c 600 gpurun_out/r2s2_bench_$c.err
done
