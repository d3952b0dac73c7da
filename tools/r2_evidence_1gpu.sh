#!/bin/bash
# round-2 evidence with the final code, one GPU: tests, smoke, the four configs, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_c2.json 2> gpurun_out/r2f_bench_c2.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "bench ref rc=$?"
for c in c3 c4 c5; do timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2f_bench_$c.json 2> gpurun_out/r2f_bench_$c.err; echo "bench $c rc=$?"; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --sustain 0 --stage-seconds 0.01 --no-cpu --no-parity > gpurun_out/r2f_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dm_ --launch-skip 27 --launch-count 9 -o gpurun_out/r2f_c2_step -f python bench.py --steps 2 --warmup 3 --sustain 0 --stage-seconds 0.01 --no-cpu --no-parity > gpurun_out/r2f_ncu_full.log 2>&1; echo "ncu full rc=$?"
DM_T=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_correlation --launch-skip 1 --launch-count 1 -o gpurun_out/r2f_t128_corr -f python tools/profile_pool.py 16 4 > gpurun_out/r2f_ncu_t128.log 2>&1; echo "ncu t128 rc=$?"
DM_T=32 DM_WS=5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_correlation --launch-skip 1 --launch-count 1 -o gpurun_out/r2f_t32ws5_corr -f python tools/profile_pool.py 3136 4 > gpurun_out/r2f_ncu_t32.log 2>&1; echo "ncu t32 rc=$?"
