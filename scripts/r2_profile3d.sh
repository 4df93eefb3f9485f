#!/bin/bash
# ncu evidence for the 3-D extension (run AFTER the same command exited 0 without ncu): launch list with durations and
# DRAM bytes of every kernel of 6 steps of Orszag-Tang 256^3, and one --set full capture of the three sweep kernels.
mkdir -p gpurun_out
CMD="python scripts/bench3d.py 256 2"
$CMD > gpurun_out/r2_3d_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_3d_plain.log; exit 1; }
tail -1 gpurun_out/r2_3d_plain.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_r2_3d.csv $CMD > gpurun_out/ncu_launch_r2_3d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep3 -s 12 -c 3 -o gpurun_out/prof_r2_3d -f $CMD > gpurun_out/ncu_full_r2_3d.log 2>&1
ls -la gpurun_out/launches_r2_3d.csv gpurun_out/prof_r2_3d.ncu-rep
