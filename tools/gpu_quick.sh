# quick GPU check: parity tests, bench, optional ncu of the solve kernel.  usage: bash tools/gpu_quick.sh <tag> [ncu]
tag=$1
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --no-cpu > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'launches',d['gpu_launches'])
for k,v in list(d['kernel_profile'].items())[:8]: print(k,v)
PY
if [ "$2" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_solve_oct -s 13 -c 1 -o gpurun_out/${tag}_solve_r0 python tools/prof_driver.py 65536 > gpurun_out/${tag}_ncu_solve.log 2>&1
fi
