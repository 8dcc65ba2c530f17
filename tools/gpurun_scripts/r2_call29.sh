# bench.py after the 16-bit-clip change: the driver's single-GPU command and the reference arm
mkdir -p gpurun_out
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 ) > gpurun_out/r3i_bench_default.log 2> gpurun_out/r3i_bench_default.err
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 ) > gpurun_out/r3i_bench_ref.log 2> gpurun_out/r3i_bench_ref.err
timeout 600 python bench.py --model cnn-trad-fpool3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3i_bench_cnn.log 2> gpurun_out/r3i_bench_cnn.err
echo finished
