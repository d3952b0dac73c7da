#!/bin/bash
# round 2, second session, FINAL code (tensor stores on) on one GPU: tests, smoke, every config, the reference arm, ncu launch list + --set full passes
mkdir -p gpurun_out
P=gpurun_out/r2x
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider > ${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${P}_pytest.log
timeout -k 10 300 python __graft_entry__.py smoke > ${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 ${P}_smoke.log
timeout -k 10 600 python bench.py --steps 20 --warmup 5 > ${P}_bench_c2.json 2> ${P}_bench_c2.err; echo "bench c2 rc=$?"
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_ref.json 2> ${P}_bench_ref.err; echo "bench ref rc=$?"
for c in c3 c4 c5 c1; do timeout -k 10 900 python bench.py --config $c --steps 5 --warmup 3 > ${P}_bench_$c.json 2> ${P}_bench_$c.err; echo "bench $c rc=$?"; done
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_launches.csv python bench.py --steps 2 --warmup 3 --sustain 0 --stage-seconds 0.01 --no-cpu --no-parity > ${P}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:dm_ --launch-skip 24 --launch-count 8 -o ${P}_c2_step -f python bench.py --steps 2 --warmup 3 --sustain 0 --stage-seconds 0.01 --no-cpu --no-parity > ${P}_ncu_full.log 2>&1; echo "ncu full rc=$?"
for f in ${P}_bench_c2.json ${P}_bench_c3.json ${P}_bench_c4.json ${P}_bench_c5.json ${P}_bench_c1.json; do python - $f <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['roofline']['stage_ms'], d['gpu_launches'], (d.get('parity') or {}).get('ok'))
except Exception as e:
    print(sys.argv[1], 'no line', e)
PY
done
DM_T=128 timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:dm_correlation --launch-skip 1 --launch-count 1 -o ${P}_t128_corr -f python tools/profile_pool.py 16 4 > ${P}_ncu_t128.log 2>&1; echo "ncu t128 rc=$?"
