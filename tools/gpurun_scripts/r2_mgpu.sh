# usage: bash tools/r2_mgpu.sh N TAG
N=$1; TAG=$2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --clip-samples 144000 --batch 1024 --scaling strong --no-cpu-baseline --no-second-mode > gpurun_out/${TAG}_strong9s_$N.log 2>gpurun_out/${TAG}_strong9s_$N.err
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --no-cpu-baseline --no-second-mode > gpurun_out/${TAG}_strong1s_$N.log 2>gpurun_out/${TAG}_strong1s_$N.err
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-second-mode > gpurun_out/${TAG}_weak1s_$N.log 2>gpurun_out/${TAG}_weak1s_$N.err
echo finished
