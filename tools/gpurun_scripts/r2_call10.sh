mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "mfcc or stream or wave or packed" > gpurun_out/r2k_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_pytest.log
timeout 120 python tools/fe_only.py > gpurun_out/r2k_fe.log 2>&1
for m in res8 res26; do
timeout 300 python bench.py --model $m --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline --no-second-mode --no-parity > gpurun_out/r2k_bench_$m.log 2>gpurun_out/r2k_bench_$m.err
done
echo finished
