mkdir -p gpurun_out/r02
export DEISM_B200_GRAPHS=0
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-configs"
python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-configs > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02/p_launches_c5.csv $B --workload c5 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02/p_launches_c3.csv $B --workload c3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02/p_launches_c1.csv $B --workload c1 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"lc_main_kernel" -c 1 -s 2 -o gpurun_out/r02/p_lc_main_c5 $B --workload c5 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"lc_main_kernel" -c 2 -s 4 -o gpurun_out/r02/p_lc_main_c5_general $B --workload c5 --general-impedance > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"org_consume_kernel|org_btable2" -c 2 -s 4 -o gpurun_out/r02/p_org_c3 $B --workload c3 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"lc_mono_kernel" -c 1 -s 2 -o gpurun_out/r02/p_lc_mono_c1 $B --workload c1 > /dev/null 2>&1
ls -la gpurun_out/r02 | grep " p_"
