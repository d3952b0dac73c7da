#!/bin/bash
# second session of round 2, call 4: persistent warp-per-parent first aggregation for small maps
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2b4_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b4_pytest.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['roofline']['stage_ms'], d['gpu_launches'])
except Exception as e:
    print('no line', e)
PY
}
i=0
for v in "DM_X=0" "DM_FIRST_CTA=1" "DM_X=0"; do
  i=$((i+1)); echo "== c4 [$v]"; env $v timeout 300 python bench.py --config c4 --steps 5 --warmup 3 --sustain 0 --no-cpu --no-parity --stage-seconds 0.1 > gpurun_out/r2b4_c4_$i.json 2> gpurun_out/r2b4_c4_$i.err; echo "rc=$?"; show gpurun_out/r2b4_c4_$i.json
done
i=0
for v in "DM_X=0" "DM_FIRST_CTA=1"; do
  i=$((i+1)); echo "== c1 [$v]"; env $v timeout 300 python bench.py --config c1 --steps 50 --warmup 5 --sustain 0 --no-cpu --stage-seconds 0.05 > gpurun_out/r2b4_c1_$i.json 2> gpurun_out/r2b4_c1_$i.err; echo "rc=$?"; show gpurun_out/r2b4_c1_$i.json
done
