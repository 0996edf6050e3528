#!/bin/bash
# ncu launch list + --set full capture of one launch per tower kernel class of the current build (run under gpurun).
set -x
cd ${GRAFT_REPO_ROOT:-.}
R=${1:-r02z}
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c3 --no-batch1 --train-steps 0 > gpurun_out/${R}_bench_plain.json 2> gpurun_out/${R}_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c3 --no-batch1 --train-steps 0 > gpurun_out/${R}_ncu_list.log 2>&1
ncu --set full --clock-control none -k regex:"gemm_bf16_tn_2cta_sched|siglip_attention_pp" \
    --launch-skip 540 -c 5 -f -o gpurun_out/${R}_full_tower python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c3 --no-batch1 --train-steps 0 \
    > gpurun_out/${R}_ncu_full_tower.log 2>&1
ls -la gpurun_out | tail -8
