# final single-GPU pass (after the key-base race fix): the regression test against the build with the race,
# the suites on the checked and the shipped build, smoke, the bench lines
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
echo "== regression test on the build WITH the race (expected: failed)"
EVK_LIB=$PWD/variants/libevk_bug.so timeout 600 python -m pytest tests/test_gpu_stress.py -m gpu -q -k one_tile_bins 2>&1 | tail -2
echo "== checked build"
EVK_LIB=$PWD/variants/libevk_checks.so timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py tests/test_gpu_fused.py tests/test_gpu_round2.py tests/test_gpu_properties.py -m gpu -q > gpurun_out/r02/pytest_checked_build.txt 2>&1; tail -2 gpurun_out/r02/pytest_checked_build.txt
echo "== shipped build"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02/pytest_gpu_final.txt 2>&1; tail -2 gpurun_out/r02/pytest_gpu_final.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02/bench_final.json 2> gpurun_out/r02/bench_final.err; tail -1 gpurun_out/r02/bench_final.err
timeout 600 python bench.py --impl reference > gpurun_out/r02/bench_final_ref.json 2>/dev/null
timeout 600 python bench.py --algo partition --steps 5 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02/bench_algo_partition.json 2>/dev/null
python profiles/tools/ds_kernel_time.py | tail -1
python -c "
import json; d=json.load(open('gpurun_out/r02/bench_final.json')); print(round(d['value']), round(d['ms_per_step'],4), d['roofline']['frac'], d['roofline']['kernel_ms'], d['step_roofline']['frac'], 'e2e', d['e2e']['ms_per_step'])"
