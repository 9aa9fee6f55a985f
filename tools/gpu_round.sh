#!/bin/bash
# Developer tool: one GPU-box pass = parity tests, both bench arms, ncu launch list, one full capture of the read kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-x}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $out/${tag}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench exit $?"
cat $out/${tag}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_ref.err; echo "ref exit $?"
cat $out/${tag}_bench_reference_arm.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
  --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rr_read_kernel -s 2 -c 1 -o $out/${tag}_read_kernel \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1; echo "ncu full exit $?"
ncu -i $out/${tag}_read_kernel.ncu-rep --page raw --csv > $out/${tag}_read_kernel_raw.csv 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -1 $out/${tag}_smoke.log
ls -la $out
