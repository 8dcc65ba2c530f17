mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "cnn" > gpurun_out/r2b_pytest_cnn.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_pytest_cnn.log
timeout 600 python bench.py --model cnn-trad-fpool3 --precision bf16 --steps 10 --warmup 3 --no-second-mode > gpurun_out/r2b_bench_cnn.log 2>gpurun_out/r2b_bench_cnn.err
echo finished
