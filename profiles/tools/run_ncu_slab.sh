# usage: run_ncu_slab.sh <variant> ...   -> gpurun_out/r02/ncu_<variant>.ncu-rep (+ raw csv)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
for v in "$@"; do
  lib=$PWD/variants/libevk_$v.so
  [ "$v" = default ] && lib=$PWD/event-camera-clustering-and-optical-flow-estimation_b200/libevk.so
  EVK_LIB=$lib timeout 600 ncu --set full --clock-control none --import-source on \
     -k regex:'k_slab_(main|pipe)' -s 2 -c 1 -f -o gpurun_out/r02/ncu_$v \
     python profiles/tools/slab_only.py > gpurun_out/r02/ncu_$v.log 2>&1
  tail -3 gpurun_out/r02/ncu_$v.log
done
