# round-2 ncu evidence (one GPU).  Every ncu command runs only after the same command has exited 0 without ncu.
# K3 on C3 and K4 on the C5-shaped batch: scripts/profile_r2.sh (TAG=...); K5: scripts/profile_k5.sh
set -x
TAG=${TAG:-r2}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --profile"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || exit 1
# launch list of the whole default bench command (all five configurations at BASELINE size)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
# K1 on C2
C2="python bench.py --only C2 --steps 2 --warmup 3 --no-cpu-baseline --no-parity --profile"
ncu --set full --clock-control none --import-source on -k regex:fbbt_single_jacobi -s 4 -c 1 -o gpurun_out/${TAG}_k1_c2 $C2 > gpurun_out/ncu_k1.log 2>&1
ls -la gpurun_out/${TAG}_*
