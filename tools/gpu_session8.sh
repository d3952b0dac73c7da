#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2s8_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2s8_pytest.log
{ python tools/time_pool.py 64 15 225; python tools/time_pool.py 128 15 32; python tools/time_pool.py 32 5 3136; } 2>&1 | tee gpurun_out/r2s8_pool.log
python tools/time_c4_e2e.py 2>&1 | tee gpurun_out/r2s8_c4.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s8_bench_c2.json 2> gpurun_out/r2s8_bench_c2.err; echo "bench c2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2s8_bench_c2.json').read().strip().splitlines()[-1])
print('c2 value %.1f ms %.3f e2e %.1f sustained %.1f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['sustained']['value']), d['roofline']['stage_ms'], d['parity']['ok'], d['parity']['max_subpix_rel'], d['parity']['score_max_abs'], d['clocks'])
PY
