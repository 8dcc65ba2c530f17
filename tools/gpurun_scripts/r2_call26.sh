# full validation of the current build: the whole GPU suite, smoke(), the default bench exactly as the driver runs it,
# the reference arm, and the launch list of the default command
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r3e_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r3e_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3e_smoke.log 2>&1
echo "rc=$?" >> gpurun_out/r3e_smoke.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 ) > gpurun_out/r3e_bench_default.log 2> gpurun_out/r3e_bench_default.err
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 ) > gpurun_out/r3e_bench_ref.log 2> gpurun_out/r3e_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r3e_ncu.log 2>&1
echo finished
