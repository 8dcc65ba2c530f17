mkdir -p gpurun_out
python - > gpurun_out/r2x_props.log 2>&1 <<'PY'
import ctypes, torch
p = torch.cuda.get_device_properties(0)
print(p)
print('L2', p.L2_cache_size)
rt = ctypes.CDLL('libcudart.so.12')
v = ctypes.c_int()
for name, attr in (('cudaDevAttrMaxPersistingL2CacheSize', 108), ('cudaDevAttrMaxAccessPolicyWindowSize', 109), ('cudaDevAttrL2CacheSize', 38)):
    rt.cudaDeviceGetAttribute(ctypes.byref(v), attr, 0)
    print(name, v.value, v.value / 2**20)
PY
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-second-mode --no-parity"
for mb in -1 32 48 64 72 -1 64; do
HONK2_TC_L2_PERSIST_MB=$mb timeout 300 $B > gpurun_out/r2x_bench_$mb.log 2>gpurun_out/r2x_bench_$mb.err
python - <<PY >> gpurun_out/r2x_summary.log
import json
try:
    d=json.loads(open('gpurun_out/r2x_bench_$mb.log').read().strip().splitlines()[-1]); r=d['roofline']
    print('persist MB $mb', round(d['value']), 'kernel ms', round(r['avg_launch_ms'],3), d['clocks']['sm_mhz'])
except Exception as e:
    print('persist MB $mb failed', e)
PY
grep PersistingL2 gpurun_out/r2x_bench_$mb.err | head -1 >> gpurun_out/r2x_summary.log
done
for mb in 48 64; do
HONK2_TC_L2_PERSIST_MB=$mb timeout 600 ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:resnet_tc_sweep -s 2 -c 1 --csv --log-file gpurun_out/r2x_ncu_$mb.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-second-mode --no-parity > gpurun_out/r2x_ncu_$mb.log 2>&1
done
echo finished
