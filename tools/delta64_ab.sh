#!/bin/bash
# A/B of the 64-column bf16-delta epilogue (RV_DELTA_PF=2, main library) vs the 32-column ring (build/var_pf1).
set -x
cd ${GRAFT_REPO_ROOT:-.}
R=${1:-r02zz}
V=$PWD/build/var_pf1/libradvlm_b200.so
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_pytest_gpu.log
tail -4 gpurun_out/${R}_pytest_gpu.log
for i in 1 2; do
  RADVLM_B200_LIB=$V python bench.py --steps 8 --no-cpu-baseline --no-c3 --train-steps 0 > gpurun_out/${R}_bench_pf1_$i.json 2>> gpurun_out/${R}_ab.err
  python bench.py --steps 8 --no-cpu-baseline --no-c3 --train-steps 0 > gpurun_out/${R}_bench_pf2_$i.json 2>> gpurun_out/${R}_ab.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${R}_bench_*.json')):
    try:
        d=json.loads([x for x in open(f) if x.startswith('{')][-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    k=d['kernel_ms_per_step']; b=d.get('batch1') or {}
    print(f, 'value %.2f e2e %.2f'%(d['ms_per_step'], d['e2e']['ms_per_step']), 'out %.2f qkv %.2f fc1 %.2f fc2 %.2f attn %.2f'%(k['gemm_out'],k['gemm_qkv'],k['gemm_fc1'],k['gemm_fc2'],k['attention']), 'b1 %.3f'%b.get('ms_per_image',0), d['clocks']['sm_mhz'])
PY
tail -3 gpurun_out/${R}_ab.err
