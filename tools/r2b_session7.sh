#!/bin/bash
# second session of round 2, call 7: 32-float staging rows for the pooled epilogue at image_size 128
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "pooled_epilogue or repeatable or t128 or tcgen05_correlation_agrees" > gpurun_out/r2b7_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b7_pytest.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d['value'],2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['roofline']['stage_ms'], d['gpu_launches'], (d.get('parity') or {}).get('ok'))
except Exception as e:
    print('no line', e)
PY
}
i=0
for v in "DM_X=0" "DM_CORR_NO_WIDE=1" "DM_X=0" "DM_CORR_NO_WIDE=1"; do
  i=$((i+1)); echo "== c5 [$v]"; env $v timeout -k 10 300 python bench.py --config c5 --steps 2 --warmup 1 --sustain 0 --no-cpu --no-parity --stage-seconds 0.1 > gpurun_out/r2b7_c5_$i.json 2> gpurun_out/r2b7_c5_$i.err; echo "rc=$?"; show gpurun_out/r2b7_c5_$i.json
done
