# usage: bash tools/gpurun_scripts/r2_mgpu3.sh N     (final build: the driver's own multi-GPU command, then config 5)
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
( time timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 ) > gpurun_out/r3f_weak1s_$N.log 2>gpurun_out/r3f_weak1s_$N.err
( time timeout 600 $TR bench.py --impl reference --gpus $N --steps 5 --warmup 1 ) > gpurun_out/r3f_ref_$N.log 2>gpurun_out/r3f_ref_$N.err
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --clip-samples 144000 --batch 1024 --scaling strong --no-cpu-baseline --no-second-mode > gpurun_out/r3f_strong9s_$N.log 2>gpurun_out/r3f_strong9s_$N.err
echo finished
