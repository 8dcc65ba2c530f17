# ncu --set full of the sweep kernel on the 9 s clips (config 5 shape), one launch of 296 clips
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:resnet_tc_sweep -s 2 -c 1 -o gpurun_out/r3o_sweep_9s -f python bench.py --clip-samples 144000 --batch 296 --steps 1 --warmup 3 --no-cpu-baseline --no-second-mode --no-parity --no-other-configs > gpurun_out/r3o_ncu.log 2>&1
echo finished
