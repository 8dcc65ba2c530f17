mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "cnn" > gpurun_out/r2d_pytest_cnn.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_pytest_cnn.log
timeout 300 python bench.py --model cnn-trad-fpool3 --precision bf16 --steps 10 --warmup 3 --no-second-mode --no-cpu-baseline --chunk 8192 > gpurun_out/r2d_bench_cnn.log 2>gpurun_out/r2d_bench_cnn.err
HONK2_TC_DEBUG=1 timeout 300 python bench.py --model cnn-trad-fpool3 --precision bf16 --steps 1 --warmup 1 --no-second-mode --no-cpu-baseline --chunk 8192 > gpurun_out/r2d_dbg.log 2>gpurun_out/r2d_dbg.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cnn_tc_fused -s 2 -c 1 -o gpurun_out/r2d_cnn_fused python bench.py --model cnn-trad-fpool3 --precision bf16 --steps 1 --warmup 3 --no-second-mode --no-cpu-baseline --chunk 8192 > gpurun_out/r2d_ncu.log 2>&1
echo finished
