#!/bin/bash
# tests after the tile-range / batch streaming / row-argmax changes, c4 and c2 bench lines, ncu evidence for round 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2s6_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2s6_pytest.log
timeout 600 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2s6_bench_c4.json 2> gpurun_out/r2s6_bench_c4.err; echo "bench c4 rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s6_bench_c2.json 2> gpurun_out/r2s6_bench_c2.err; echo "bench c2 rc=$?"
# ncu: launch list of the default bench command (kernel shares), then --set full of one C2 step and of the T=128 / ws=5 correlation kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --sustain 0 --stage-seconds 0.01 --no-cpu --no-parity > gpurun_out/r2_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dm_ --launch-skip 27 --launch-count 9 -o gpurun_out/r2_c2_step -f python bench.py --steps 2 --warmup 3 --sustain 0 --stage-seconds 0.01 --no-cpu --no-parity > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"
DM_T=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_correlation --launch-skip 1 --launch-count 1 -o gpurun_out/r2_t128_corr -f python tools/profile_pool.py 16 4 > gpurun_out/r2_ncu_t128.log 2>&1; echo "ncu t128 rc=$?"
DM_T=32 DM_WS=5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_correlation --launch-skip 1 --launch-count 1 -o gpurun_out/r2_t32ws5_corr -f python tools/profile_pool.py 3136 4 > gpurun_out/r2_ncu_t32.log 2>&1; echo "ncu t32 rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -5
