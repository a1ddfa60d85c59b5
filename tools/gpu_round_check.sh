# Full evidence run on one B200 (under gpurun): parity tests, both bench arms, the ncu launch list of the bench command and
# ncu --set full captures of the dominant kernels.  usage: bash tools/gpu_round_check.sh <tag>
tag=${1:-final}
nproc
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 400 gpurun_out/${tag}_bench.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-profile > gpurun_out/${tag}_ncu_bench.log 2>&1
# ncu --set full of one LOADED launch of each main kernel (launch-skip counts launches of that kernel only)
for spec in "k_solve_thread:2:solve_thread" "SetupMellingerFn<.int.0>:2:SetupHead" "SetupMellingerFn<.int.1>:2:SetupRow" "CoefCostGradFn:2:CoefCostGradFn" "ExtremaRawFn<.int.1>:0:ExtremaRawFn1"; do
  pat=${spec%%:*}; rest=${spec#*:}; skip=${rest%%:*}; name=${rest#*:}
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$pat" -s $skip -c 1 -o gpurun_out/${tag}_$name python tools/prof_driver.py 65536 1 > gpurun_out/${tag}_ncu_$name.log 2>&1
done
# general-shape path (tg_solve_linear_batch_nd): bench beside the oracle, ncu --set full of its three kernels at N = 12
timeout 300 python tools/bench_general_shape.py > gpurun_out/${tag}_general_shape.json 2> gpurun_out/${tag}_general_shape.err
for spec in "RecordHeadFn<.int.12>:GenRecordHead12" "RecordHrowFn<.int.12>:GenRecordHrow12" "SolveFn<.int.12>:GenSolve12"; do
  pat=${spec%%:*}; name=${spec#*:}
  TG_GS_B=16384 timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$pat" -s 2 -c 1 -o gpurun_out/${tag}_$name python tools/bench_general_shape.py > gpurun_out/${tag}_ncu_$name.log 2>&1
done
# gpurun brings back at most 64 MiB: keep the CSV pages of every report, drop the reports themselves
for rep in gpurun_out/${tag}_*.ncu-rep; do
  ncu -i $rep --page raw --csv > ${rep%.ncu-rep}_raw.csv 2>/dev/null
  ncu -i $rep --page source --csv --print-source cuda,sass > ${rep%.ncu-rep}_sass.csv 2>/dev/null
  rm -f $rep
done
(timeout 900 compute-sanitizer --tool memcheck python tools/prof_driver.py 256 2>&1 | tail -3) > gpurun_out/${tag}_memcheck.log
(timeout 900 compute-sanitizer --tool racecheck python tools/prof_driver.py 64 2 2>&1 | tail -3) > gpurun_out/${tag}_racecheck.log
cat gpurun_out/${tag}_memcheck.log gpurun_out/${tag}_racecheck.log
ls -la gpurun_out | grep ${tag}
