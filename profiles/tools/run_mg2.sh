cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
N=${NGPU:-2}
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "$N" -s > gpurun_out/r02/mg_pytest_n$N.txt 2>&1; grep -E "mg ok|passed|failed|Error|error|assert" gpurun_out/r02/mg_pytest_n$N.txt | tail -20
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-extras > gpurun_out/r02/bench_n${N}_time.json 2> gpurun_out/r02/bench_n${N}_time.err; tail -2 gpurun_out/r02/bench_n${N}_time.err; cut -c1-330 gpurun_out/r02/bench_n${N}_time.json
EVK_MIX_TRACE=1 timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-extras --owner mix64 > gpurun_out/r02/bench_n${N}_mix64.json 2> gpurun_out/r02/bench_n${N}_mix64.err; grep "evk mix64 rank 0" gpurun_out/r02/bench_n${N}_mix64.err | tail -2; tail -2 gpurun_out/r02/bench_n${N}_mix64.err; cut -c1-330 gpurun_out/r02/bench_n${N}_mix64.json
if [ -n "$C4" ]; then
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 2 --config c4 > gpurun_out/r02/bench_n${N}_c4.json 2> gpurun_out/r02/bench_n${N}_c4.err; tail -2 gpurun_out/r02/bench_n${N}_c4.err; cut -c1-330 gpurun_out/r02/bench_n${N}_c4.json
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 2 --config c4 --owner mix64 > gpurun_out/r02/bench_n${N}_c4_mix64.json 2> gpurun_out/r02/bench_n${N}_c4_mix64.err; tail -2 gpurun_out/r02/bench_n${N}_c4_mix64.err; cut -c1-330 gpurun_out/r02/bench_n${N}_c4_mix64.json
fi
