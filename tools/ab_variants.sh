for v in poly0 poly4 poly6; do
  RADVLM_B200_LIB=$PWD/build/var_$v/libradvlm_b200.so python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v.json 2>/dev/null
  python - <<P
import json
d=json.loads(open("gpurun_out/ab_$v.json").read().strip().splitlines()[-1])
print("$v", round(d["value"]), round(d["ms_per_step"],2), "attn", round(d["kernel_ms_per_step"]["attention"],2), d["clocks"]["sm_mhz"])
P
done
