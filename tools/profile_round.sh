# final-state evidence of a round: default bench line, then the ncu launch list of a short run of the same bench
set -x
R=${1:-r1f}
python bench.py > gpurun_out/${R}_bench_default.json 2> gpurun_out/${R}_bench.err
echo rc_bench=$?
B="python bench.py --steps 20 --warmup 3 --no-cpu"
$B > gpurun_out/${R}_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_bench_launches.csv $B > gpurun_out/${R}_ncu_bench.log 2>&1
echo rc_launches=$?
