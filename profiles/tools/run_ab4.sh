cd $GRAFT_REPO_ROOT
for v in $VARIANTS $VARIANTS; do
  EVK_LIB=$PWD/variants/libevk_$v.so timeout 200 python profiles/tools/ds_kernel_time.py 2>&1 | tail -1
done
