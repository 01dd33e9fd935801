# checked slab build against the parity / stress tests, then the full GPU suite and smoke() on the shipped build
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
EVK_LIB=$PWD/variants/libevk_checks.so timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py tests/test_gpu_fused.py tests/test_gpu_round2.py tests/test_gpu_properties.py -m gpu -q > gpurun_out/r02/pytest_checked_build.txt 2>&1; tail -4 gpurun_out/r02/pytest_checked_build.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02/pytest_gpu_final.txt 2>&1; tail -4 gpurun_out/r02/pytest_gpu_final.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
