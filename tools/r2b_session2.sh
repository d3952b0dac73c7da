#!/bin/bash
# second session of round 2, call 2: descriptor shuffle sums (parity + timing), PDL launch by launch, first-aggregation CTA size, stream chunks
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "descriptor or correlation_and_pyramid or c4_pair or c2_bench or fused_path_equals or correction_slots or image_cut_solver_vs" > gpurun_out/r2b2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b2_pytest.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['roofline']['stage_ms'], d['gpu_launches'])
except Exception as e:
    print('no line', e)
PY
}
B="--steps 30 --warmup 5 --sustain 0 --no-cpu --no-parity --stage-seconds 0.05"
i=0
for v in "DM_PDL=0" "DM_PDL=1" "DM_PDL=2" "DM_PDL=4" "DM_PDL=8" "DM_PDL=16" "DM_PDL=0" "DM_STREAM_CHUNKS=2" "DM_STREAM_CHUNKS=4" "DM_STREAM_CHUNKS=1"; do
  i=$((i+1)); echo "== c2 [$v]"; env $v timeout 200 python bench.py $B > gpurun_out/r2b2_c2_$i.json 2> gpurun_out/r2b2_c2_$i.err; echo "rc=$?"; show gpurun_out/r2b2_c2_$i.json
done
i=0
for v in "DM_PDL=0" "DM_FIRST_THREADS=32" "DM_PDL=0"; do
  i=$((i+1)); echo "== c4 [$v]"; env $v timeout 300 python bench.py --config c4 --steps 5 --warmup 3 --sustain 0 --no-cpu --no-parity --stage-seconds 0.1 > gpurun_out/r2b2_c4_$i.json 2> gpurun_out/r2b2_c4_$i.err; echo "rc=$?"; show gpurun_out/r2b2_c4_$i.json
done
