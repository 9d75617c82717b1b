set -x
cd $GRAFT_REPO_ROOT
K='regex:conv_igemm|conv_patch|convt_fused|conv_rowstack|conv_swap'
timeout 600 python bench.py --profile-out gpurun_out/r2_final_launch_profile_cuda_events.txt > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err
timeout 900 python bench.py --impl reference > gpurun_out/r2_final_bench_reference_arm.json 2> gpurun_out/r2_final_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 460 --csv --log-file gpurun_out/r2_final_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k "$K" -s 186 -c 62 --csv --log-file gpurun_out/r2_final_ncu_step.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 186 -c 62 -o /tmp/r2_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ncu3.log 2>&1
ncu -i /tmp/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_final_ncu_full_raw.csv 2>/dev/null
ls -la /tmp/r2_full.ncu-rep gpurun_out/
