set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
nvidia-smi -L
# sanity at small size first (hang guard), per variant
for v in p32 p23 p22 p32b p23b r1; do
  EVK_AB_EVENTS=3000000 EVK_LIB=$PWD/variants/libevk_$v.so timeout 120 python profiles/tools/ds_kernel_time.py 2>&1 | tail -1
done > gpurun_out/r02/ab1_small.txt 2>&1
cat gpurun_out/r02/ab1_small.txt
for v in r1 p32 p23 p22 p32b p23b; do
  EVK_LIB=$PWD/variants/libevk_$v.so timeout 200 python profiles/tools/ds_kernel_time.py 2>&1 | tail -1
done > gpurun_out/r02/ab1.txt 2>&1
cat gpurun_out/r02/ab1.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02/pytest1.txt 2>&1; tail -15 gpurun_out/r02/pytest1.txt
