#!/bin/bash
# NOTE: gpurun copies back at most 64 MiB of gpurun_out/: the three --set full reports together exceed that, so run the
# script once per report (comment the others out) or delete earlier reports first.
# Round profile capture (run under gpurun): plain bench first, then the ncu launch list and --set full captures of one
# launch of every kernel class.  Numbers printed by the runs under ncu are never bench values.  Summaries for profiles/
# are made in the container afterwards (tools/ncu_summary.py launches | full | stalls | traffic).
set -x
R=${1:-r02}
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c3 --no-batch1 --train-steps 0 > gpurun_out/${R}_ncu_list.log 2>&1
# one launch of each tower kernel class of a middle layer (patch + 26 x [qkv, attention, out, fc1, fc2] per tower call)
ncu --set full --clock-control none -k regex:"gemm_bf16_tn_2cta_sched|siglip_attention_pp" \
    --launch-skip 540 -c 5 -f -o gpurun_out/${R}_full_tower python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c3 --no-batch1 --train-steps 0 \
    > gpurun_out/${R}_ncu_full_tower.log 2>&1
ncu --set full --clock-control none -k regex:"merge_splice_kernel|resample_fused|im2col" \
    --launch-skip 8 -c 4 -f -o gpurun_out/${R}_full_misc python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c3 --no-batch1 --train-steps 0 \
    > gpurun_out/${R}_ncu_full_misc.log 2>&1
# training mode: bench line, then one --set full launch of each backward kernel class
python bench.py --mode train --batch 4 --steps 3 --warmup 3 > gpurun_out/${R}_bench_train.json 2> gpurun_out/${R}_bench_train.err
ncu --set full --clock-control none \
    -k regex:"siglip_attention_bwd|layernorm_bwd|colsum|attn_delta|attn_dq_store" \
    --launch-skip 40 -c 6 -f -o gpurun_out/${R}_full_train python bench.py --mode train --batch 4 --steps 1 --warmup 3 > gpurun_out/${R}_ncu_full_train.log 2>&1
tail -c 600 gpurun_out/${R}_bench.json
