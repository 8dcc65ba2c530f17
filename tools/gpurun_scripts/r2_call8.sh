mkdir -p gpurun_out
for m in res8 res26; do
HONK2_TC_DEBUG=1 timeout 300 python bench.py --model $m --precision bf16 --steps 1 --warmup 1 --no-cpu-baseline --no-second-mode --no-parity > gpurun_out/r2i_dbg_$m.log 2>gpurun_out/r2i_dbg_$m.err
done
echo finished
