# conv_0 + pool pre-pass without the pad channels: packed-strip tests, res8 / res26 throughput
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "packed or golden or permutation" > gpurun_out/r3v_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r3v_pytest.log
for model in res8 res26; do
timeout 600 python bench.py --model $model --steps 10 --warmup 3 --no-cpu-baseline --no-second-mode --no-parity --no-other-configs > gpurun_out/r3v_bench_${model}.log 2> gpurun_out/r3v_bench_${model}.err
done
echo finished
