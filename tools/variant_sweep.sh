#!/bin/bash
# Runs bench.py (no CPU leg) on the default build, on every library under variants/ and on env-switch variants of the default build;
# one JSON per variant in gpurun_out/.  Scratch helper for A/B sessions on the GPU box.
python bench.py --no-cpu --steps 5 --warmup 3 > gpurun_out/var_default.json 2> gpurun_out/var_default.err
echo "== default"; python tools/bench_short.py gpurun_out/var_default.json 8
for so in variants/libtg_*.so; do
  [ -e "$so" ] || continue
  n=$(basename $so .so); n=${n#libtg_}
  TG_LIB=$PWD/$so python bench.py --no-cpu --steps 5 --warmup 3 > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err
  echo "== $n"; python tools/bench_short.py gpurun_out/var_$n.json 8
done
for e in "$@"; do
  n=$(echo $e | tr '= ' '__')
  env $e python bench.py --no-cpu --steps 5 --warmup 3 > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err
  echo "== $e"; python tools/bench_short.py gpurun_out/var_$n.json 8
done
