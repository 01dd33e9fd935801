cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
timeout 300 python profiles/tools/ds_kernel_time.py 2>&1 | tail -1
timeout 900 python profiles/tools/unordered_time.py > gpurun_out/r02/unordered.jsonl 2> gpurun_out/r02/unordered.err; tail -3 gpurun_out/r02/unordered.err; cat gpurun_out/r02/unordered.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -4
