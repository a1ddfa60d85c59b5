tag=$1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:ExtremaRawFn<.int.1>" -s 2 -c 1 -o gpurun_out/${tag}_extrema1 python tools/prof_driver.py 65536 > gpurun_out/${tag}_ncu_extrema1.log 2>&1
tail -3 gpurun_out/${tag}_ncu_extrema1.log
