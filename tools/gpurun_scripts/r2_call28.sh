# fp32 row kernel: three staging buffers (one barrier per chunk) + batched skip loads: tests and throughput
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fp32 or resident or empty_batch or weights_follow" > gpurun_out/r3h_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r3h_pytest.log
B="python bench.py --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline --no-second-mode --no-parity --no-other-configs"
for model in res15 res15_narrow; do
timeout 600 $B --model $model --batch 2048 > gpurun_out/r3h_bench_${model}.log 2> gpurun_out/r3h_bench_${model}.err
done
echo finished
