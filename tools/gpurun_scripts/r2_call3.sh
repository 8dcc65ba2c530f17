mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "cnn" > gpurun_out/r2c_pytest_cnn.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_pytest_cnn.log
for c in 1024 2048 4096 8192; do
timeout 300 python bench.py --model cnn-trad-fpool3 --precision bf16 --steps 10 --warmup 3 --no-second-mode --no-cpu-baseline --chunk $c > gpurun_out/r2c_bench_cnn_$c.log 2>gpurun_out/r2c_bench_cnn_$c.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c_cnn_launches.csv python bench.py --model cnn-trad-fpool3 --precision bf16 --steps 1 --warmup 3 --no-second-mode --no-cpu-baseline > gpurun_out/r2c_ncu1.log 2>&1
echo finished
