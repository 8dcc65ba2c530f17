# ncu --set full: CNN fp32 conv_1 (conv_row_f32_kernel<4>) and the int16 PCM front-end
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_row_f32_kernel -s 7 -c 1 -o gpurun_out/r3s_cnn_f32_conv1 -f python bench.py --precision fp32 --model cnn-trad-fpool3 --batch 4096 --steps 1 --warmup 3 --no-cpu-baseline --no-second-mode --no-parity --no-other-configs > gpurun_out/r3s_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:mfcc_kernel -s 12 -c 1 -o gpurun_out/r3s_mfcc_pcm16 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-second-mode --no-parity --no-other-configs > gpurun_out/r3s_ncu2.log 2>&1
echo finished
