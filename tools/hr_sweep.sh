python tools/bench_lookup.py --nwave 100001 --models 1 --fused-models 10000 --steps 3 2>/dev/null | tail -1 > gpurun_out/r02_hr_1e5_m10000.json; cut -c1-200 gpurun_out/r02_hr_1e5_m10000.json
python tools/bench_lookup.py --nwave 300001 --models 1 --fused-models 1000 --steps 3 2>/dev/null | tail -1 > gpurun_out/r02_hr_3e5_m1000.json; cut -c1-200 gpurun_out/r02_hr_3e5_m1000.json
timeout 900 python tools/bench_lookup.py --nwave 1000001 --models 1 --fused-models 1000 --steps 2 2>/dev/null | tail -1 > gpurun_out/r02_hr_1e6_m1000.json; cut -c1-200 gpurun_out/r02_hr_1e6_m1000.json
python -c "
import json
for f in ('1e5_m10000','3e5_m1000','1e6_m1000'):
    try:
        d=json.load(open('gpurun_out/r02_hr_%s.json'%f)); print(f, d['init_s'], d['lookup'][0], d['fused_eclipse'])
    except Exception as e: print(f, 'failed', e)
"
