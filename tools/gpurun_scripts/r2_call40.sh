# PCM16 streaming front-end: tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stream or pcm16 or mfcc" > gpurun_out/r3u_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r3u_pytest.log
echo finished
