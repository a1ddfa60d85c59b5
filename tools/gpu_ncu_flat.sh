# ncu --set full captures of the flat kernels (names need the demangled base to tell the functors apart)
tag=$1
for pat in CoefCostFn SetupMellingerFn "ExtremaRawFn<1>" "ExtremaRawFn<2>"; do
  name=$(echo $pat | tr -d '<>')
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$pat" -s 1 -c 1 -o gpurun_out/${tag}_$name python tools/prof_driver.py 65536 > gpurun_out/${tag}_ncu_$name.log 2>&1
done
ls -la gpurun_out | tail -8
