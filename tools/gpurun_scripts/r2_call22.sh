# fp32 resident-weight convolution kernel: parity tests, then old vs new throughput (res15 and res15-narrow, fp32)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fp32 or resident or empty_batch or weights_follow" > gpurun_out/r3a_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r3a_pytest.log
B="python bench.py --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline --no-second-mode --no-parity"
for model in res15 res15_narrow res8; do
for r in 0 1; do
HONK2_F32_RESIDENT=$r timeout 600 $B --model $model --batch 2048 > gpurun_out/r3a_bench_${model}_$r.log 2> gpurun_out/r3a_bench_${model}_$r.err
done
done
echo finished
