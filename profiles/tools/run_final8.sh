# final 8-GPU pass of round 2: N = 1 on the same box, the multi-GPU parity tests, weak scaling (C3 per GPU) and C4
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
N=${NGPU:-8}
python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02/bench_n1_samebox_as_n$N.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02/bench_n1_samebox_as_n$N.json')); print('N=1', round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['stage_ms'].items() if isinstance(v,float)})"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-extras > gpurun_out/r02/bench_n${N}_time.json 2> gpurun_out/r02/bench_n${N}_time.err; tail -1 gpurun_out/r02/bench_n${N}_time.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/r02/bench_n${N}_time.json')); print('N=$N', round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['stage_ms'].items() if isinstance(v,float)}, 'e2e', d['e2e']['ms_per_step'], d['h2d_only']['GBps_all_gpus'])"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "$N" -s > gpurun_out/r02/mg_pytest_n$N.txt 2>&1; grep -c "mg ok" gpurun_out/r02/mg_pytest_n$N.txt; grep -E "passed|failed" gpurun_out/r02/mg_pytest_n$N.txt | tail -2
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 2 --config c4 > gpurun_out/r02/bench_n${N}_c4.json 2> gpurun_out/r02/bench_n${N}_c4.err; cut -c1-330 gpurun_out/r02/bench_n${N}_c4.json
