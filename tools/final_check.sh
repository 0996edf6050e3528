#!/bin/bash
# Round-end check (run under gpurun): the full GPU test suite, the plain bench line, then an interleaved A/B of the
# CUDA-graph replay of encode_images (RADVLM_B200_GRAPH=0 launches every kernel eagerly).
set -x
cd ${GRAFT_REPO_ROOT:-.}
R=${1:-r02r}
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_pytest_gpu.log
tail -4 gpurun_out/${R}_pytest_gpu.log
( time python bench.py ) > gpurun_out/${R}_bench_final.json 2> gpurun_out/${R}_bench_final.err; echo "bench rc=$?"
for i in 1 2; do
  RADVLM_B200_GRAPH=0 python bench.py --no-cpu-baseline --no-c3 --train-steps 0 > gpurun_out/${R}_bench_graph0_$i.json 2>> gpurun_out/${R}_ab.err
  RADVLM_B200_GRAPH=1 python bench.py --no-cpu-baseline --no-c3 --train-steps 0 > gpurun_out/${R}_bench_graph1_$i.json 2>> gpurun_out/${R}_ab.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${R}_bench_*.json')):
    try:
        d=json.loads([x for x in open(f) if x.startswith('{')][-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    b=d.get('batch1') or {}
    print(f, round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d.get('cuda_graph'), round(b.get('ms_per_image',0),3), b.get('encode_only'), d['clocks']['sm_mhz'])
PY
