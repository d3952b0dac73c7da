#!/bin/bash
# round 2, second session: the last code change (32-float staging at image_size 128) -- full tests, the C5 line, the T = 128 kernel under ncu
mkdir -p gpurun_out
P=gpurun_out/r2y
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider > ${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${P}_pytest.log
timeout -k 10 300 python __graft_entry__.py smoke > ${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 ${P}_smoke.log
timeout -k 10 900 python bench.py --config c5 --steps 5 --warmup 3 > ${P}_bench_c5.json 2> ${P}_bench_c5.err; echo "bench c5 rc=$?"
timeout -k 10 600 python bench.py --steps 20 --warmup 5 > ${P}_bench_c2.json 2> ${P}_bench_c2.err; echo "bench c2 rc=$?"
DM_T=128 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:dm_correlation --launch-skip 1 --launch-count 1 -o ${P}_t128_corr -f python tools/profile_pool.py 16 4 > ${P}_ncu_t128.log 2>&1; echo "ncu t128 rc=$?"
for f in ${P}_bench_c2.json ${P}_bench_c5.json; do python - $f <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['roofline']['stage_ms'], d['gpu_launches'], (d.get('parity') or {}).get('ok'))
except Exception as e:
    print(sys.argv[1], 'no line', e)
PY
done
