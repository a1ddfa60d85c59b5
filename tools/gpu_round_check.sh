set -x
nproc
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/s3_pytest.log 2>&1; tail -3 gpurun_out/s3_pytest.log
timeout 600 python bench.py > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; tail -c 600 gpurun_out/s3_bench.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s3_bench_ref.json 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s3_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-profile > gpurun_out/s3_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_solve_oct -s 30 -c 2 -o gpurun_out/s3_solve python tools/prof_driver.py 65536 > gpurun_out/s3_ncu_solve.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ExtremaRawFn.1 -s 3 -c 2 -o gpurun_out/s3_extrema python tools/prof_driver.py 65536 > gpurun_out/s3_ncu_extrema.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"CoefCostFn|SetupMellingerFn" -s 20 -c 2 -o gpurun_out/s3_coef python tools/prof_driver.py 65536 > gpurun_out/s3_ncu_coef.log 2>&1
ls -la gpurun_out
