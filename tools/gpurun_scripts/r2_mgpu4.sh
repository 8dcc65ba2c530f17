# usage: bash tools/gpurun_scripts/r2_mgpu4.sh N     (final build: the driver's own multi-GPU command only)
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
( time timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 ) > gpurun_out/r3n_weak1s_$N.log 2>gpurun_out/r3n_weak1s_$N.err
echo finished
