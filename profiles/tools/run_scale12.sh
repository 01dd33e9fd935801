cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1', round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['stage_ms'].items() if isinstance(v,float)})"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
$TR bench.py --gpus 2 --steps 20 --warmup 3 --no-extras --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=2', round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['stage_ms'].items() if isinstance(v,float)}, 'e2e', d['e2e']['ms_per_step'], d['h2d_only'])"
