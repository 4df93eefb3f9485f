#!/bin/bash
# ncu evidence of round 2 (run AFTER the same command exited 0 without ncu): launch list with durations and DRAM
# bytes of every kernel of the timed steps, and one --set full capture of the dominant kernel (x and y sweep).
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 0 --strong-steps 0"
$CMD > gpurun_out/r2_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/ncu_launch_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:flux_ws -s 24 -c 2 -o gpurun_out/prof_r2f -f $CMD > gpurun_out/ncu_full_r2.log 2>&1
ls -la gpurun_out/launches_r2.csv gpurun_out/prof_r2f.ncu-rep
