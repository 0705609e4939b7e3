#!/bin/bash
# tools/profile_r2.sh TAG -- the profile set of a build, run on the GPU box (gpurun).  Outputs under gpurun_out/
# (summaries are copied to profiles/ by hand):
#   ${TAG}_launch_list.csv   per-launch time / DRAM bytes / warp instructions of one 1024-sequence chunk (one lane)
#   ${TAG}_fp64_list.csv     per-launch fp64 arithmetic thread-instructions of a 128-sequence chunk (multi-pass metrics)
#   ${TAG}_ph<k>.ncu-rep     `--set full` captures of the largest phase kernels at a mid span
#   ${TAG}_ph4_tma.ncu-rep   the same inside-B launch of the bulk-copy staged build (A/B of LIN_SPLIT_TMA)
#   ${TAG}_vit.ncu-rep       the Viterbi kernel of a 2048-read scan
TAG=${1:-r2}
OUT=gpurun_out
CMD="python bench.py --nseq 512 --steps 1 --warmup 0 --no-cpu-baseline --no-scan"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum
RELEM_LANES=1 ncu --metrics $M --clock-control none -c 560 --csv --log-file $OUT/${TAG}_launch_list.csv $CMD > $OUT/${TAG}_launch_list.log 2>&1
F=gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum
RELEM_LANES=1 ncu --metrics $F --clock-control none -c 560 --csv --log-file $OUT/${TAG}_fp64_list.csv python bench.py --nseq 64 --steps 1 --warmup 0 --no-cpu-baseline --no-scan > $OUT/${TAG}_fp64_list.log 2>&1
for ph in 7 13 4 1; do
  RELEM_LANES=1 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:phase_kernel<\(int\)${ph}," \
     --launch-skip 20 --launch-count 1 -o $OUT/${TAG}_ph${ph} -f $CMD > $OUT/${TAG}_ph${ph}.log 2>&1
done
if [ -f rnaelem_b200/variants/librelem_tma.so ]; then
  RELEM_LIBRARY=rnaelem_b200/variants/librelem_tma.so RELEM_LANES=1 ncu --set full --import-source on --clock-control none --kernel-name-base demangled \
     -k "regex:phase_kernel<\(int\)4," --launch-skip 20 --launch-count 1 -o $OUT/${TAG}_ph4_tma -f $CMD > $OUT/${TAG}_ph4_tma.log 2>&1
fi
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:relem_viterbi_kernel --launch-count 1 \
   -o $OUT/${TAG}_vit -f python tools/scan_probe.py 2048 > $OUT/${TAG}_vit.log 2>&1
ls -la $OUT/${TAG}_*
