# ncu captures of the per-round kernels (K5) on C4, one GPU: round 1 (every row due) and round 2 of the rows kernel, the
# variables kernel of round 1
set -x
TAG=${TAG:-r2}
C4="python bench.py --only C4 --no-cpu-baseline --no-parity --profile"
$C4 > gpurun_out/prof_c4_plain.json 2> gpurun_out/prof_c4_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:rounds_rows -c 2 -o gpurun_out/${TAG}_k5_rows $C4 > gpurun_out/ncu_k5_rows.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rounds_vars_list -c 1 -o gpurun_out/${TAG}_k5_vars $C4 > gpurun_out/ncu_k5_vars.log 2>&1
ls -la gpurun_out/*k5*.ncu-rep
