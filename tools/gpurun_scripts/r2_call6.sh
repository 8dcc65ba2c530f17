mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_res15.log 2>gpurun_out/r2f_bench_res15.err
for m in res8 res26 res15_narrow res8_narrow res26_narrow; do
timeout 300 python bench.py --model $m --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_$m.log 2>gpurun_out/r2f_bench_$m.err
done
timeout 400 python bench.py --model hey_snips_res26 --precision bf16 --clip-samples 144000 --batch 1024 --steps 5 --warmup 3 --no-cpu-baseline --no-second-mode > gpurun_out/r2f_bench_heysnips26.log 2>gpurun_out/r2f_bench_heysnips26.err
timeout 400 python bench.py --model res15 --precision bf16 --clip-samples 144000 --batch 1024 --steps 5 --warmup 3 --no-cpu-baseline --no-second-mode > gpurun_out/r2f_bench_res15_9s.log 2>gpurun_out/r2f_bench_res15_9s.err
echo finished
