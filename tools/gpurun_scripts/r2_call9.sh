mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "packed or res8 or res26 or golden" > gpurun_out/r2j_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_pytest.log
for m in res8 res26; do
timeout 300 python bench.py --model $m --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench_$m.log 2>gpurun_out/r2j_bench_$m.err
done
echo finished
