# two GPUs: the model classes under torch.nn.DataParallel (run/test.py:69-70)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "data_parallel or weights_follow or host_pipeline" > gpurun_out/r3k_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r3k_pytest.log
echo finished
