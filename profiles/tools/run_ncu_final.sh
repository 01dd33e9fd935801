# ncu --set full of the fused step's kernels on the shipped build (one invocation per kernel: each replays ~40x)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_slab_main' -s 8 -c 1 -f \
   -o gpurun_out/r02/ncu_slab_final python profiles/tools/ds_kernel_time.py > gpurun_out/r02/ncu_slab_final.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_km_assign_tiles' -s 2 -c 1 -f \
   -o gpurun_out/r02/ncu_assign_final python profiles/tools/ds_kernel_time.py > gpurun_out/r02/ncu_assign_final.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:'k_slab_fix|k_slab_bins|k_km_finalise' -s 16 -c 3 -f \
   -o gpurun_out/r02/ncu_small_final python profiles/tools/ds_kernel_time.py > gpurun_out/r02/ncu_small_final.log 2>&1
tail -1 gpurun_out/r02/ncu_slab_final.log gpurun_out/r02/ncu_assign_final.log gpurun_out/r02/ncu_small_final.log
