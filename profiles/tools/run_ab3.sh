cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
for v in $VARIANTS; do
  EVK_AB_EVENTS=3000000 EVK_LIB=$PWD/variants/libevk_$v.so timeout 120 python profiles/tools/ds_kernel_time.py 2>&1 | tail -1
done > gpurun_out/r02/ab3_small.txt 2>&1
cat gpurun_out/r02/ab3_small.txt
for v in $VARIANTS; do
  EVK_LIB=$PWD/variants/libevk_$v.so timeout 200 python profiles/tools/ds_kernel_time.py 2>&1 | tail -1
done > gpurun_out/r02/ab3.txt 2>&1
cat gpurun_out/r02/ab3.txt
