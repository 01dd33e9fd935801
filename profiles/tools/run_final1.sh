# final single-GPU pass of round 2: tests (checked + shipped build), smoke, bench lines, ncu captures
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
EVK_LIB=$PWD/variants/libevk_checks.so timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py tests/test_gpu_fused.py tests/test_gpu_round2.py tests/test_gpu_properties.py -m gpu -q > gpurun_out/r02/pytest_checked_build.txt 2>&1; tail -2 gpurun_out/r02/pytest_checked_build.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02/pytest_gpu_final.txt 2>&1; tail -2 gpurun_out/r02/pytest_gpu_final.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02/bench_final.json 2> gpurun_out/r02/bench_final.err; tail -2 gpurun_out/r02/bench_final.err
timeout 600 python bench.py --impl reference > gpurun_out/r02/bench_final_ref.json 2>/dev/null
timeout 600 python bench.py --algo table --steps 5 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02/bench_algo_table.json 2>/dev/null
timeout 600 python bench.py --algo partition --steps 5 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02/bench_algo_partition.json 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on \
   -k regex:'k_slab_main' -s 8 -c 1 -f -o gpurun_out/r02/ncu_slab_final \
   python profiles/tools/ds_kernel_time.py > gpurun_out/r02/ncu_slab_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
   -k regex:'k_km_assign_tiles|k_slab_fix|k_slab_bins|k_km_finalise' -s 12 -c 4 -f -o gpurun_out/r02/ncu_assign_final \
   python profiles/tools/ds_kernel_time.py > gpurun_out/r02/ncu_assign_final.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r02/launches_final.csv python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-extras > gpurun_out/r02/ncu_launches.log 2>&1
tail -1 gpurun_out/r02/ncu_launches.log | cut -c1-300
python profiles/tools/unordered_time.py > gpurun_out/r02/unordered.jsonl 2> gpurun_out/r02/unordered.err; tail -3 gpurun_out/r02/unordered.jsonl | cut -c1-200
