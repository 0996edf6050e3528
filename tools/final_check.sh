#!/bin/bash
# Round-end check (run under gpurun): the full GPU test suite, smoke(), and the driver-style bench line.
set -x
cd ${GRAFT_REPO_ROOT:-.}
R=${1:-r02x}
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_pytest_gpu.log
tail -4 gpurun_out/${R}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/${R}_smoke.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/${R}_bench_final.json 2> gpurun_out/${R}_bench_final.err; echo "bench rc=$?"
( time python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 ) > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err; echo "reference rc=$?"
python - <<PY
import json
d=json.loads([x for x in open('gpurun_out/${R}_bench_final.json') if x.startswith('{')][-1])
b=d.get('batch1') or {}; t=d.get('train') or {}
print('value %.0f %.2f ms e2e %.0f %.2f ms'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']), 'path', round(d['path_frac_of_peak'],3), 'fc2', round(d['roofline']['frac'],3), 'b1 %.3f'%b.get('ms_per_image',0), 'train %.2f'%t.get('ms_per_step',0), 'c3', (d.get('c3_strong') or {}).get('ms_per_global_batch'), d['clocks'])
print(d['kernel_ms_per_step'])
r=json.loads([x for x in open('gpurun_out/${R}_bench_reference.json') if x.startswith('{')][-1])
print('reference', r.get('value'), r.get('cpu_baseline'))
PY
