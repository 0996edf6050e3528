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
ncu --set full --clock-control none --import-source on -k regex:"merge_splice|cast_f32" \
    -c 2 -f -o gpurun_out/${R}_full_merge python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu_full_merge.log 2>&1
tail -c 600 gpurun_out/${R}_bench.json
