cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "d3_d4 or ref_hash or realloc" 2>&1 | tail -5
timeout 600 python profiles/tools/d3_time.py > gpurun_out/r02/d3_time.jsonl 2>gpurun_out/r02/d3_time.err; tail -2 gpurun_out/r02/d3_time.err; cat gpurun_out/r02/d3_time.jsonl
