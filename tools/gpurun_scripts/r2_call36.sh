# CNN family fp32: row-tile convolution kernel + split-K first Linear + sub-batch 1024: tests and throughput
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cnn or fp32" > gpurun_out/r3q_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r3q_pytest.log
B="python bench.py --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline --no-second-mode --no-parity --no-other-configs --model cnn-trad-fpool3 --batch 4096"
HONK2_F32_RESIDENT=0 timeout 600 $B > gpurun_out/r3q_bench_cnn_0.log 2> gpurun_out/r3q_bench_cnn_0.err
HONK2_F32_RESIDENT=1 timeout 600 $B > gpurun_out/r3q_bench_cnn_1.log 2> gpurun_out/r3q_bench_cnn_1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r3q_launches_cnn_fp32.csv $B --steps 1 > gpurun_out/r3q_ncu.log 2>&1
echo finished
