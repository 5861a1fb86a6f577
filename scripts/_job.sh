mkdir -p gpurun_out/r02
ncu --set full --import-source on --clock-control none -k regex:"org_consume_kernel|org_btable2" -c 2 -s 6 -o gpurun_out/r02/q_org_c3 python bench.py --workload c3 --steps 2 --warmup 2 --no-cpu-baseline --no-extras --no-configs > gpurun_out/r02/q_org.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:lc_main_kernel -c 2 -s 6 -o gpurun_out/r02/q_lc_c5_general python bench.py --workload c5 --general-impedance --steps 2 --warmup 2 --no-cpu-baseline --no-extras --no-configs > gpurun_out/r02/q_lcg.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02/q_launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-configs > gpurun_out/r02/q_l3.log 2>&1
ls -la gpurun_out/r02/ | grep q_
