set -x
BENCH="python bench.py --steps 20 --warmup 5"
$BENCH > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err || exit 1
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm_n1.json 2>/dev/null
LITE="python bench.py --steps 20 --warmup 5 --no-extras --no-e2e --no-cpu-baseline --no-parity"
$LITE > gpurun_out/r02_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv $LITE > gpurun_out/r02_ncu_launch.log 2>&1
python tools/prof_hm.py > gpurun_out/r02_plain_hm.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'hist_u8_planar_lane_pw|apply_u8_planar_vec|build_lut' -s 3 -c 6 -o gpurun_out/r02_prof_hm -f python tools/prof_hm.py > gpurun_out/r02_ncu_hm.log 2>&1
python tools/prof_macenko.py macenko > gpurun_out/r02_plain_mk.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'t_moments_kernel|t_resolve_kernel|mid_kernel|select_recover_kernel|apply_kernel' -s 9 -c 9 -o gpurun_out/r02_prof_mk -f python tools/prof_macenko.py macenko > gpurun_out/r02_ncu_mk.log 2>&1
python tools/prof_reinhard.py 32 > gpurun_out/r02_plain_rh.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'stats_kernel|finalize_kernel|apply_kernel' -s 3 -c 3 -o gpurun_out/r02_prof_rh -f python tools/prof_reinhard.py 32 > gpurun_out/r02_ncu_rh.log 2>&1
ls -la gpurun_out/r02_prof_*.ncu-rep
for c in c3 c5 reinhard; do python bench.py --config $c > gpurun_out/r02_bench_${c}_n1.json 2> gpurun_out/r02_bench_${c}_n1.err; done
