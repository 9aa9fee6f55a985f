#!/bin/bash
# Developer tool: instruction / local-memory / issue counters of the read kernel for library variants (one ncu launch each).
# usage (under gpurun): bash tools/ncu_counts.sh <coverage> lib1[@VAR=value] lib2 ...
cov=$1; shift
M=smsp__inst_executed.sum,smsp__sass_inst_executed_op_local_ld.sum,smsp__sass_inst_executed_op_local_st.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio
for spec in "$@"; do
  lib=${spec%%@*}; set_=${spec#*@}; [ "$set_" == "$spec" ] && set_=
  echo "== $spec"
  env RR_LIB=$lib ${set_//,/ } NODL=1 ncu --metrics $M --clock-control none -k regex:rr_read_kernel -s 1 -c 1 python tools/quick_perf.py 5000000 $cov 2>&1 | grep -E "smsp__|gpu__time" | awk '{print "   ", $1, $NF}'
done
