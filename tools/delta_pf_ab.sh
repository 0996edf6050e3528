#!/bin/bash
# A/B of the residual-prefetch ring of the out_proj (bf16-delta) epilogue: main library vs build/var_nopf (-DRV_DELTA_PF=0).
set -x
cd ${GRAFT_REPO_ROOT:-.}
R=${1:-r02w}
V=$PWD/build/var_nopf/libradvlm_b200.so
python tools/encode_checksum.py > gpurun_out/${R}_sha_pf.txt 2> gpurun_out/${R}_sha.err
RADVLM_B200_LIB=$V python tools/encode_checksum.py > gpurun_out/${R}_sha_nopf.txt 2>> gpurun_out/${R}_sha.err
cat gpurun_out/${R}_sha_pf.txt gpurun_out/${R}_sha_nopf.txt
cmp gpurun_out/${R}_sha_pf.txt gpurun_out/${R}_sha_nopf.txt && echo "BIT-EQUAL" || echo "DIFFERENT"
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_pytest_gpu.log
tail -4 gpurun_out/${R}_pytest_gpu.log
for i in 1 2; do
  RADVLM_B200_LIB=$V python bench.py --no-cpu-baseline --no-c3 --no-batch1 --train-steps 0 > gpurun_out/${R}_bench_nopf_$i.json 2>> gpurun_out/${R}_ab.err
  python bench.py --no-cpu-baseline --no-c3 --no-batch1 --train-steps 0 > gpurun_out/${R}_bench_pf_$i.json 2>> gpurun_out/${R}_ab.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${R}_bench_*.json')):
    try:
        d=json.loads([x for x in open(f) if x.startswith('{')][-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    k=d['kernel_ms_per_step']
    print(f, 'value %.2f e2e %.2f'%(d['ms_per_step'], d['e2e']['ms_per_step']), 'out %.2f qkv %.2f fc1 %.2f fc2 %.2f attn %.2f'%(k['gemm_out'],k['gemm_qkv'],k['gemm_fc1'],k['gemm_fc2'],k['attention']), d['clocks']['sm_mhz'])
PY
tail -3 gpurun_out/${R}_ab.err
