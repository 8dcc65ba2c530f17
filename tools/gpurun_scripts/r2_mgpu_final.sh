# usage: bash tools/gpurun_scripts/r2_mgpu_final.sh N
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-second-mode > gpurun_out/r2v_weak1s_$N.log 2>gpurun_out/r2v_weak1s_$N.err
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --clip-samples 144000 --batch 1024 --scaling strong --no-cpu-baseline --no-second-mode > gpurun_out/r2v_strong9s_$N.log 2>gpurun_out/r2v_strong9s_$N.err
echo finished
