set -x
cd $GRAFT_REPO_ROOT
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02q_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02q_pytest_gpu.log
tail -5 gpurun_out/r02q_pytest_gpu.log
( time python bench.py ) > gpurun_out/r02q_bench_final.json 2> gpurun_out/r02q_bench_final.err; echo "bench rc=$?"
tail -3 gpurun_out/r02q_bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r02q_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c3 --no-batch1 --train-steps 0 > gpurun_out/r02q_ncu_list.log 2>&1
ncu --set full --clock-control none -k regex:"gemm_bf16_tn_2cta_sched|siglip_attention_pp" \
    --launch-skip 540 -c 5 -f -o gpurun_out/r02q_full_tower python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c3 --no-batch1 --train-steps 0 \
    > gpurun_out/r02q_ncu_full_tower.log 2>&1
ls -la gpurun_out | tail -12
python -c "
import json
l=[x for x in open('gpurun_out/r02q_bench_final.json') if x.startswith('{')][-1]
d=json.loads(l)
print({k:d[k] for k in ('value','ms_per_step','e2e','batch1','clocks','path_frac_of_peak') if k in d})
print(d.get('roofline',{}).get('frac'))
"
