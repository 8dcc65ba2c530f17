# fp32 convolution kernels on shapes outside the zoo; bench line with hey_snips res26 in other_configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "outside_the_zoo" > gpurun_out/r3m_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r3m_pytest.log
( time timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/r3m_bench_default.log 2> gpurun_out/r3m_bench_default.err
echo finished
