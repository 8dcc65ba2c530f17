mkdir -p gpurun_out
for m in res15 res15_narrow; do
HONK2_TC_DEBUG=1 timeout 300 python bench.py --model $m --precision bf16 --steps 1 --warmup 1 --no-cpu-baseline --no-second-mode --no-parity > gpurun_out/r2t_dbg_$m.log 2>gpurun_out/r2t_dbg_$m.err
done
echo finished
