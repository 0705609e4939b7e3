#!/bin/bash
# tools/profile_r2.sh TAG -- the profile set of a build, run on the GPU box (gpurun): launch list with per-launch DRAM
# bytes / time / instruction / fp64-op counters of one 1024-sequence chunk, and `--set full` captures of the largest
# phase kernels at a mid span.  Outputs under gpurun_out/ (copied to profiles/ by hand).
TAG=${1:-r2}
OUT=gpurun_out
CMD="python bench.py --nseq 512 --steps 1 --warmup 0 --no-cpu-baseline --no-scan"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum
RELEM_LANES=1 ncu --metrics $M --clock-control none -c 520 --csv --log-file $OUT/${TAG}_launch_list.csv $CMD > $OUT/${TAG}_launch_list.log 2>&1
for ph in 7 13 4 0 1 5; do
  RELEM_LANES=1 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:phase_kernel<${ph}," \
     --launch-skip 20 --launch-count 1 -o $OUT/${TAG}_ph${ph} -f $CMD > $OUT/${TAG}_ph${ph}.log 2>&1
done
ls -la $OUT/${TAG}_*
