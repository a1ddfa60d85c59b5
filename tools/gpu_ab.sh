# A/B of an environment switch inside ONE gpurun call (boxes differ by a few per cent in host speed): alternates the two settings.
# usage: bash tools/gpu_ab.sh <tag> <ENVVAR> <valueA> <valueB> [rounds]
tag=$1; var=$2; a=$3; b=$4; n=${5:-3}
for i in $(seq 1 $n); do
  for v in $a $b; do
    env $var=$v timeout 300 python bench.py --no-cpu --no-profile --steps 10 --warmup 3 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$var=$v', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2))
" | tee -a gpurun_out/${tag}_ab.log
  done
done
