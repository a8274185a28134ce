# round-2 ncu captures: K3 on C3 (full size) and K3+K4 on a C5-shaped batch small enough for ncu's replays
set -x
C3="python bench.py --only C3 --no-cpu-baseline --no-parity --profile"
C5="python bench.py --only C5 --no-cpu-baseline --no-parity --profile --c5-cons 200000 --c5-boxes 1024"
$C3 > gpurun_out/prof_c3_plain.json 2> gpurun_out/prof_c3_plain.err || exit 1
$C5 > gpurun_out/prof_c5_plain.json 2> gpurun_out/prof_c5_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:fbbt_batch_reference -c 1 -o gpurun_out/${TAG}_k3_c3 $C3 > gpurun_out/ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fbbt_batch_reference -c 1 -o gpurun_out/${TAG}_k4_c5 $C5 > gpurun_out/ncu_c5.log 2>&1
ls -la gpurun_out/*.ncu-rep
