cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02/pytest2.txt 2>&1; tail -15 gpurun_out/r02/pytest2.txt
timeout 900 python bench.py > gpurun_out/r02/bench_a.json 2> gpurun_out/r02/bench_a.err; tail -3 gpurun_out/r02/bench_a.err; cat gpurun_out/r02/bench_a.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02/bench_ref_a.json 2>/dev/null; cat gpurun_out/r02/bench_ref_a.json
