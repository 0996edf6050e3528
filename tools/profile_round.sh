#!/bin/bash
# Round profile capture (run under gpurun): plain bench first, then the ncu launch list and --set full captures of one
# launch of every kernel class.  Numbers printed by the runs under ncu are never bench values.
set -x
R=${1:-r01}
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16|siglip_attention|layernorm|im2col|resample" \
    -c 13 -f -o gpurun_out/${R}_full_tower python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu_full_tower.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"merge_splice_kernel" \
    -c 1 -f -o gpurun_out/${R}_full_merge python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu_full_merge.log 2>&1
# training mode: bench line, then one --set full launch of each backward kernel class
python bench.py --mode train --batch 4 --steps 3 --warmup 3 > gpurun_out/${R}_bench_train.json 2> gpurun_out/${R}_bench_train.err
ncu --set full --clock-control none --import-source on \
    -k regex:"siglip_attention_bwd|layernorm_bwd|gelu_fwd_bwd|colsum|attn_delta|attn_dq_store|gemm_bf16_tn_2cta_kernel<256, 7>|gemm_bf16_tn_2cta_kernel<256, 0>" \
    -c 14 -f -o gpurun_out/${R}_full_train python bench.py --mode train --batch 4 --steps 1 --warmup 3 > gpurun_out/${R}_ncu_full_train.log 2>&1
tail -c 600 gpurun_out/${R}_bench.json
