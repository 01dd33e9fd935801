cd $GRAFT_REPO_ROOT
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
EVK_MIX_TRACE=1 timeout 600 $TR bench.py --gpus $N --steps 4 --warmup 2 --no-extras --owner mix64 --e2e-steps 1 2>&1 >/dev/null | grep "evk mix64 rank 0" | tail -2
echo "-- all records written locally (timing experiment, results wrong)"
EVK_MIX_TIMING_LOCAL_ONLY=1 EVK_MIX_TRACE=1 timeout 600 $TR bench.py --gpus $N --steps 4 --warmup 2 --no-extras --owner mix64 --e2e-steps 1 2>&1 >/dev/null | grep "evk mix64 rank 0" | tail -2
