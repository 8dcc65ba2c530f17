mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "narrow or label_counts or odd_map or packed" > gpurun_out/r2s_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_pytest.log
for m in res15_narrow res8_narrow res26_narrow; do
timeout 300 python bench.py --model $m --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_bench_$m.log 2>gpurun_out/r2s_bench_$m.err
done
echo finished
