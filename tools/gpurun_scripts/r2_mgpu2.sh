N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
for m in cnn-trad-fpool3 res8 res26; do
timeout 600 $TR bench.py --gpus $N --model $m --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline --no-second-mode > gpurun_out/r2r_${m}_2.log 2>gpurun_out/r2r_${m}_2.err
done
timeout 600 python -m pytest tests/test_dist_cpu.py -q > gpurun_out/r2r_dist.log 2>&1
echo finished
