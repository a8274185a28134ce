set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile --batch-boxes 8192 --c4-rows-per-rank 1000000 --c5-boxes 1024"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fbbt_single_jacobi -s 4 -c 1 -o gpurun_out/r1_k1_c2 $CMD > gpurun_out/ncu_a.log 2>&1
python scripts/k1_probe.py 2500000 0 > gpurun_out/prof_stream_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:fbbt_single_jacobi -s 6 -c 1 -o gpurun_out/r1_k1_stream python scripts/k1_probe.py 2500000 0 > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rounds_rows_kernel -s 36 -c 1 -o gpurun_out/r1_k5_rows python scripts/k1_probe.py 2500000 1 > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep
