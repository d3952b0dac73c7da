#!/bin/bash
# second session of round 2, call 10: tensor stores at image_size 128 (own regions) and 32 (pair regions)
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "pooled_epilogue or repeatable or c4_pair or t128_tiles or variants_agree or fused_path_equals or image_cut_solver_vs" > gpurun_out/r2b10_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b10_pytest.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['roofline']['stage_ms'], d['gpu_launches'], (d.get('parity') or {}).get('ok'))
except Exception as e:
    print('no line', e)
PY
}
i=0
for v in "DM_X=0" "DM_CORR_NO_TMA_STORE=1"; do
  i=$((i+1)); echo "== c4 [$v]"; env $v timeout -k 10 200 python bench.py --config c4 --steps 5 --warmup 3 --sustain 0 --no-cpu --no-parity --stage-seconds 0.1 > gpurun_out/r2b10_c4_$i.json 2> gpurun_out/r2b10_c4_$i.err; echo "rc=$?"; show gpurun_out/r2b10_c4_$i.json
done
i=0
for v in "DM_X=0" "DM_CORR_NO_TMA_STORE=1"; do
  i=$((i+1)); echo "== c5 [$v]"; env $v timeout -k 10 200 python bench.py --config c5 --steps 2 --warmup 1 --sustain 0 --no-cpu --no-parity --stage-seconds 0.1 > gpurun_out/r2b10_c5_$i.json 2> gpurun_out/r2b10_c5_$i.err; echo "rc=$?"; show gpurun_out/r2b10_c5_$i.json
done
