#!/bin/bash
# 2-GPU check (gpurun --gpus 2): the multi-GPU pytest wrappers, then the driver-style bench launch at N = 2.
set -x
cd ${GRAFT_REPO_ROOT:-.}
R=${1:-r02u}
( time python -m pytest tests -m gpu -x -q -k "multi_gpu or scatter or peer" ) > gpurun_out/${R}_pytest_n2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_pytest_n2.log
tail -4 gpurun_out/${R}_pytest_n2.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 5 ) > gpurun_out/${R}_bench_n2.json 2> gpurun_out/${R}_bench_n2.err; echo "bench rc=$?"
tail -3 gpurun_out/${R}_bench_n2.err
python - <<PY
import json
d=json.loads([x for x in open('gpurun_out/${R}_bench_n2.json') if x.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e'], d.get('gather_check'), d.get('cuda_graph'))
print({k:d['train'][k] for k in ('value','ms_per_step')} if 'train' in d else None, (d.get('c3_strong') or {}).get('ms_per_global_batch'))
PY
