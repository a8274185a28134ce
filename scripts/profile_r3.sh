# final ncu evidence of round 2's last session (TAG=r3): launch list of the default bench command, K3 on C3 (TMA gather4),
# K3+K4 on the C5-shaped batch small enough for ncu's replays.  Every ncu command runs only after the same command has
# exited 0 without ncu.
set -x
TAG=${TAG:-r3}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --profile"
C3="python bench.py --only C3 --no-cpu-baseline --no-parity --profile"
C5="python bench.py --only C5 --no-cpu-baseline --no-parity --profile --c5-cons 200000 --c5-boxes 1024"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
$C3 > gpurun_out/prof_c3_plain.json 2> gpurun_out/prof_c3_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:fbbt_batch_reference -c 1 -o gpurun_out/${TAG}_k3_c3 $C3 > gpurun_out/ncu_c3.log 2>&1
$C5 > gpurun_out/prof_c5_plain.json 2> gpurun_out/prof_c5_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:fbbt_batch_reference -c 1 -o gpurun_out/${TAG}_k4_c5 $C5 > gpurun_out/ncu_c5.log 2>&1
ls -la gpurun_out/${TAG}_*
