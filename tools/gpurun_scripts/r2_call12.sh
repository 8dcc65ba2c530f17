mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "cnn" > gpurun_out/r2m_pytest_cnn.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_pytest_cnn.log
timeout 300 python bench.py --model cnn-trad-fpool3 --precision bf16 --steps 10 --warmup 3 --no-second-mode --no-cpu-baseline --chunk 8192 > gpurun_out/r2m_bench_cnn.log 2>gpurun_out/r2m_bench_cnn.err
HONK2_TC_DEBUG=1 timeout 300 python bench.py --model cnn-trad-fpool3 --precision bf16 --steps 1 --warmup 1 --no-second-mode --no-cpu-baseline --no-parity --chunk 8192 > gpurun_out/r2m_dbg.log 2>gpurun_out/r2m_dbg.err
echo finished
