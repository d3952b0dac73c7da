#!/bin/bash
# second session of round 2, call 1: parity with the merged launches / PDL, then A/B timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2b1_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b1_pytest.log
for v in "" "DM_AGG_NO_MERGE=1" "DM_AGG_BUDGET_KB=64" "DM_AGG_BUDGET_KB=16"; do
  echo "== aggregate $v"; env $v timeout 120 python tools/time_aggregate.py 2>&1 | tail -3
done
B="--steps 20 --warmup 5 --sustain 0 --no-cpu --no-parity"
i=0
for v in "" "DM_PDL=0" "DM_PDL=0 DM_DESC_SPLIT=1 DM_AGG_NO_MERGE=1" ""; do
  i=$((i+1)); echo "== c2 [$v]"; env $v timeout 300 python bench.py $B > gpurun_out/r2b1_c2_$i.json 2> gpurun_out/r2b1_c2_$i.err; echo "rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2b1_c2_$i.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['roofline']['stage_ms'], d['gpu_launches'])
PY
done
i=0
for v in "" "DM_PDL=0 DM_DESC_SPLIT=1 DM_AGG_NO_MERGE=1"; do
  i=$((i+1)); echo "== c4 [$v]"; env $v timeout 300 python bench.py --config c4 --steps 5 --warmup 3 --sustain 0 --no-cpu --no-parity > gpurun_out/r2b1_c4_$i.json 2> gpurun_out/r2b1_c4_$i.err; echo "rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2b1_c4_$i.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['roofline']['stage_ms'], d['gpu_launches'])
PY
done
