#!/bin/bash
# second session of round 2, call 3: first aggregation with per-warp cooperative child statistics and hoisted loads
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "fused or c2_bench or c4_pair or oracle_tile or t128 or scene_c2 or image_cut_solver or batch_of_pairs or chunked" > gpurun_out/r2b3_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b3_pytest.log
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
for v in "DM_X=0" "DM_FIRST_OLD=1" "DM_FIRST_PPC=2" "DM_FIRST_PPC=4" "DM_X=0"; do
  i=$((i+1)); echo "== c4 [$v]"; env $v timeout 300 python bench.py --config c4 --steps 5 --warmup 3 --sustain 0 --no-cpu --no-parity --stage-seconds 0.1 > gpurun_out/r2b3_c4_$i.json 2> gpurun_out/r2b3_c4_$i.err; echo "rc=$?"; show gpurun_out/r2b3_c4_$i.json
done
B="--steps 30 --warmup 5 --sustain 0 --no-cpu --no-parity --stage-seconds 0.05"
i=0
for v in "DM_X=0" "DM_FIRST_OLD=1" "DM_X=0" "DM_FIRST_OLD=1"; do
  i=$((i+1)); echo "== c2 [$v]"; env $v timeout 200 python bench.py $B > gpurun_out/r2b3_c2_$i.json 2> gpurun_out/r2b3_c2_$i.err; echo "rc=$?"; show gpurun_out/r2b3_c2_$i.json
done
i=0
for v in "DM_X=0" "DM_FIRST_OLD=1"; do
  i=$((i+1)); echo "== c5 [$v]"; env $v timeout 300 python bench.py --config c5 --steps 2 --warmup 1 --sustain 0 --no-cpu --no-parity --stage-seconds 0.1 > gpurun_out/r2b3_c5_$i.json 2> gpurun_out/r2b3_c5_$i.err; echo "rc=$?"; show gpurun_out/r2b3_c5_$i.json
done
