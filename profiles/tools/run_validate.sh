cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
echo "== checked slab build (EVK_SLAB_CHECKS=1) against parity / stress / fused / round-2 tests"
EVK_LIB=$PWD/variants/libevk_checks.so timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py tests/test_gpu_fused.py tests/test_gpu_round2.py tests/test_gpu_properties.py -m gpu -q > gpurun_out/r02/pytest_checked_build.txt 2>&1; tail -4 gpurun_out/r02/pytest_checked_build.txt
echo "== full GPU suite, shipped build"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02/pytest_gpu_final.txt 2>&1; tail -4 gpurun_out/r02/pytest_gpu_final.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02/bench_b.json 2> gpurun_out/r02/bench_b.err; tail -2 gpurun_out/r02/bench_b.err; python -c "
import json; d=json.load(open('gpurun_out/r02/bench_b.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['roofline']['kernel_ms'], d['step_roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], d['e2e_serial']['ms_per_step'], d['e2e_centroids']['ms_per_step'], d['h2d_only']['ms_per_step'])
print(d['extra_keys']['c5']); print(d['extra_keys']['c1']['gpu_ms_per_step'], d['cpu_baseline'])"
timeout 600 python bench.py --algo table --steps 5 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02/bench_algo_table.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02/bench_algo_table.json')); print('table', d['ms_per_step'], d['ds_algo'], d['roofline']['kernel'], d['roofline']['kernel_ms'])"
