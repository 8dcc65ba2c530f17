# bench.py contract test on the GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bench.py -x -q > gpurun_out/r3r_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r3r_pytest.log
echo finished
