mkdir -p gpurun_out
for m in res15 res8 res26 res15_narrow cnn-trad-fpool3; do
P=bf16
timeout 400 python bench.py --model $m --precision $P --steps 10 --warmup 3 --no-cpu-baseline --no-second-mode --no-parity --gpu-eager-bar > gpurun_out/r2o_bench_$m.log 2>gpurun_out/r2o_bench_$m.err
done
echo finished
