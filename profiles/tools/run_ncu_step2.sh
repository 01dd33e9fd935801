# ncu --set full of the fused step's two heavy kernels on the shipped build
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
timeout 200 python profiles/tools/ds_kernel_time.py 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on \
   -k regex:'k_slab_main' -s 8 -c 1 -f -o gpurun_out/r02/ncu_slab_final \
   python profiles/tools/ds_kernel_time.py > gpurun_out/r02/ncu_slab_final.log 2>&1
tail -2 gpurun_out/r02/ncu_slab_final.log
timeout 600 ncu --set full --clock-control none --import-source on \
   -k regex:'k_km_assign_tiles|k_slab_fix' -s 8 -c 2 -f -o gpurun_out/r02/ncu_assign_final \
   python profiles/tools/ds_kernel_time.py > gpurun_out/r02/ncu_assign_final.log 2>&1
tail -2 gpurun_out/r02/ncu_assign_final.log
