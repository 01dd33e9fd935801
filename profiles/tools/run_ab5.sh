# A/B of slab-kernel builds on one box: "lib" = the shipped libevk.so, else variants/libevk_<name>.so
cd $GRAFT_REPO_ROOT
for v in $VARIANTS $VARIANTS; do
  if [ $v = lib ]; then unset EVK_LIB; else export EVK_LIB=$PWD/variants/libevk_$v.so; fi
  timeout 200 python profiles/tools/ds_kernel_time.py 2>&1 | tail -1
done
