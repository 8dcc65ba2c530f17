# fp32 ResNet path with equal sub-batches: tests and throughput
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fp32 or resident or outside" > gpurun_out/r3w_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r3w_pytest.log
B="python bench.py --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline --no-second-mode --no-parity --no-other-configs"
timeout 600 $B --model res15 --batch 2048 > gpurun_out/r3w_bench_res15.log 2> gpurun_out/r3w_bench_res15.err
echo finished
