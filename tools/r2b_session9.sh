#!/bin/bash
# second session of round 2, call 9: the pair regions of the pooled epilogue leave by 2-D tensor stores (DM_CORR_TMA_STORE)
mkdir -p gpurun_out
DM_CORR_TMA_STORE=1 timeout -k 10 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "pooled_epilogue or repeatable or c2_bench or fused_path_equals or scene_c2" > gpurun_out/r2b9_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b9_pytest.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['roofline']['stage_ms'], d['gpu_launches'], (d.get('parity') or {}).get('ok'))
except Exception as e:
    print('no line', e)
PY
}
B="--steps 30 --warmup 5 --sustain 0 --no-cpu --stage-seconds 0.05"
i=0
for v in "DM_X=0" "DM_CORR_TMA_STORE=1" "DM_X=0" "DM_CORR_TMA_STORE=1"; do
  i=$((i+1)); echo "== c2 [$v]"; env $v timeout -k 10 120 python bench.py $B > gpurun_out/r2b9_c2_$i.json 2> gpurun_out/r2b9_c2_$i.err; echo "rc=$?"; show gpurun_out/r2b9_c2_$i.json
done
