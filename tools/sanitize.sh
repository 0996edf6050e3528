#!/bin/bash
# compute-sanitizer passes (run under gpurun): memcheck, racecheck, synccheck, initcheck over smoke() (whole path forward
# + backward on a reduced tower) and memcheck / synccheck over the building-block GPU tests.  Logs -> gpurun_out/<R>_sanitizer_*.log
R=${1:-r02}
export PYTHONWARNINGS=ignore
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --log-file gpurun_out/${R}_sanitizer_${tool}_smoke.log \
      python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_sanitizer_${tool}_smoke.out 2>&1
  echo "$tool smoke rc=$?" | tee -a gpurun_out/${R}_sanitizer_summary.log
  tail -3 gpurun_out/${R}_sanitizer_${tool}_smoke.log | tee -a gpurun_out/${R}_sanitizer_summary.log
done
for tool in memcheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 --log-file gpurun_out/${R}_sanitizer_${tool}_blocks.log \
      python -m pytest tests/test_gpu_parity.py -q -x -m gpu \
      -k "gemm or qkv_split or layernorm or merge_splice_vs or video_merge_vs or scatter or encoder_small" \
      > gpurun_out/${R}_sanitizer_${tool}_blocks.out 2>&1
  echo "$tool blocks rc=$?" | tee -a gpurun_out/${R}_sanitizer_summary.log
  tail -3 gpurun_out/${R}_sanitizer_${tool}_blocks.log | tee -a gpurun_out/${R}_sanitizer_summary.log
  tail -2 gpurun_out/${R}_sanitizer_${tool}_blocks.out | tee -a gpurun_out/${R}_sanitizer_summary.log
done
