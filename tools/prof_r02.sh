#!/bin/bash
# Round-2 profiling pass (run on the GPU box from the repo root): launch list of a short bench run + one --set full
# capture of every kernel of the library at the C2 shape.  Outputs under gpurun_out/.
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_launches_bench.log 2>&1
# second warm pass of every kernel: skip the first pass' launches of the same kernels by capturing with -s
ncu --set full --clock-control none --import-source on \
    -k regex:'dag_|grad_|lsg_|viterbi|glat|decode|posterior|xchg' -s 12 -c 14 -f -o gpurun_out/r02_prof \
    python tools/prof_dp.py 2 --viterbi --lsg > gpurun_out/r02_prof.log 2>&1
ls -la gpurun_out/r02_prof.ncu-rep
