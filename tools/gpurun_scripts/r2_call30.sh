# two GPUs: the model classes under torch.nn.DataParallel (run/test.py:69-70), then the driver's 2-GPU bench command
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "data_parallel" > gpurun_out/r3j_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r3j_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519"
( time timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/r3j_weak1s_2.log 2>gpurun_out/r3j_weak1s_2.err
echo finished
