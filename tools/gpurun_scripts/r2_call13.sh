mkdir -p gpurun_out
for r in 2 3 4 5; do
HONK2_CNN_RING=$r HONK2_TC_DEBUG=1 timeout 300 python bench.py --model cnn-trad-fpool3 --precision bf16 --steps 1 --warmup 1 --no-second-mode --no-cpu-baseline --no-parity --chunk 8192 > gpurun_out/r2n_dbg_$r.log 2>gpurun_out/r2n_dbg_$r.err
done
echo finished
