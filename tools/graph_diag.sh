#!/bin/bash
# Diagnostics of the CUDA-graph replay of encode_images: per-step device / host times with the graph path on and off.
set -x
cd ${GRAFT_REPO_ROOT:-.}
R=${1:-r02s}
for i in 1 2 3; do
  RADVLM_BENCH_STEP_TIMES=1 RADVLM_B200_GRAPH=1 python bench.py --steps 20 --no-cpu-baseline --no-c3 --train-steps 0 > gpurun_out/${R}_graph1_$i.json 2> gpurun_out/${R}_graph1_$i.err
  RADVLM_BENCH_STEP_TIMES=1 RADVLM_B200_GRAPH=0 python bench.py --steps 20 --no-cpu-baseline --no-c3 --train-steps 0 > gpurun_out/${R}_graph0_$i.json 2> gpurun_out/${R}_graph0_$i.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${R}_graph*.json')):
    try:
        d=json.loads([x for x in open(f) if x.startswith('{')][-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    b=d.get('batch1') or {}
    print(f, round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d.get('cuda_graph'), round(b.get('ms_per_image',0),3), b.get('encode_only'), d['clocks']['sm_mhz'])
PY
grep -h "step times" gpurun_out/${R}_graph1_*.err | cut -c1-400
