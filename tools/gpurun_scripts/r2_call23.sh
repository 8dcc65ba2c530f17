# fp32 resident-weight convolution kernel + PCM16 front-end: tests, then fp32 old vs new, then the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fp32 or resident or empty_batch or weights_follow or pcm16 or host_pipeline or mfcc" > gpurun_out/r3b_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r3b_pytest.log
B="python bench.py --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline --no-second-mode --no-parity"
for model in res15 res8 res26; do
for r in 0 1; do
HONK2_F32_RESIDENT=$r timeout 600 $B --model $model --batch 2048 > gpurun_out/r3b_bench_${model}_$r.log 2> gpurun_out/r3b_bench_${model}_$r.err
done
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3b_bench_default.log 2> gpurun_out/r3b_bench_default.err
timeout 600 python bench.py --model cnn-trad-fpool3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3b_bench_cnn.log 2> gpurun_out/r3b_bench_cnn.err
echo finished
