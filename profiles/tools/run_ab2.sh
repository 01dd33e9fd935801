cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
for v in r1 r1c2 p32; do
  EVK_AB_NOREP=1 EVK_LIB=$PWD/variants/libevk_$v.so timeout 200 python profiles/tools/ds_kernel_time.py 2>&1 | tail -1
done > gpurun_out/r02/ab2.txt 2>&1
cat gpurun_out/r02/ab2.txt
