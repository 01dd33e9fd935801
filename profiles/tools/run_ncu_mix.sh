cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
EVK_MIX_TRACE=1 timeout 300 python profiles/tools/mix_single.py 2>&1 | tail -4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_mix_hist|k_mix_scatter' -s 2 -c 2 -f -o gpurun_out/r02/ncu_mix python profiles/tools/mix_single.py > gpurun_out/r02/ncu_mix.log 2>&1
tail -2 gpurun_out/r02/ncu_mix.log
