# ncu --set full of the step's kernels on the default build -> gpurun_out/r02/ncu_step.ncu-rep
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
timeout 900 ncu --set full --clock-control none --import-source on \
   -k regex:'k_slab_main|k_km_assign_tiles|k_slab_fix|k_slab_bins' -s 12 -c 4 -f -o gpurun_out/r02/ncu_step \
   python profiles/tools/ds_kernel_time.py > gpurun_out/r02/ncu_step.log 2>&1
tail -3 gpurun_out/r02/ncu_step.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r02/launches_r02.csv python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-extras > gpurun_out/r02/ncu_launches.log 2>&1
tail -2 gpurun_out/r02/ncu_launches.log
