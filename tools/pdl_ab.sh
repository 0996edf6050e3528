#!/bin/bash
# A/B of programmatic dependent launch (RADVLM_B200_PDL=0 launches the GEMM / attention kernels the ordinary way):
# full GPU tests with PDL on, then interleaved bench pairs.
set -x
cd ${GRAFT_REPO_ROOT:-.}
R=${1:-r02v}
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_pytest_gpu.log
tail -4 gpurun_out/${R}_pytest_gpu.log
for i in 1 2; do
  RADVLM_B200_PDL=0 python bench.py --no-cpu-baseline --no-c3 --train-steps 6 > gpurun_out/${R}_bench_pdl0_$i.json 2>> gpurun_out/${R}_ab.err
  RADVLM_B200_PDL=1 python bench.py --no-cpu-baseline --no-c3 --train-steps 6 > gpurun_out/${R}_bench_pdl1_$i.json 2>> gpurun_out/${R}_ab.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${R}_bench_*.json')):
    try:
        d=json.loads([x for x in open(f) if x.startswith('{')][-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    b=d.get('batch1') or {}
    t=d.get('train') or {}
    print(f, 'value %.2f e2e %.2f'%(d['ms_per_step'], d['e2e']['ms_per_step']), 'b1 %.3f'%b.get('ms_per_image',0), {k:round(v,3) for k,v in (b.get('encode_only') or {}).items() if k.endswith('_ms')}, 'train %.2f'%t.get('ms_per_step',0), d['clocks']['sm_mhz'], 'prof %.2f sumk %.2f'%(d['profiled_pass']['ms_per_step'], d['profiled_pass']['sum_of_kernels_ms_per_step']))
PY
tail -5 gpurun_out/${R}_ab.err
