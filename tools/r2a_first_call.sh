mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_gpu_tests.log 2>&1; echo rc=$? >> gpurun_out/r2a_gpu_tests.log
V=$PWD/thermoelasticity_fem_b200/build/variants
( echo "== select (default build): staged order, bank-aware"; python tools/asm_ab.py 100 512,620,8,1 512,620,8,1,1;
  echo "== tabload (round-1 table loads), bank-aware"; TEFEM_LIB=$V/libtefem_tabload.so python tools/asm_ab.py 100 512,620,8,1,1;
  echo "== pitched lists, 448-block tiles, bank-aware"; TEFEM_PITCH_LISTS=1 python tools/asm_ab.py 100 448,540,8,1,1;
  echo "== 128 threads, 256-block tiles, bank-aware"; TEFEM_ASM_THREADS=128 python tools/asm_ab.py 100 256,310,8,1,1;
  echo "== persm768, bank-aware"; TEFEM_LIB=$V/libtefem_persm768.so python tools/asm_ab.py 100 512,620,8,1,1 ) > gpurun_out/r2a_k1_variants.log 2>&1
python tools/asm_ab.py 100 512,620,8,1,1 > gpurun_out/r2a_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:assemble_kernel -s 10 -c 1 -o gpurun_out/r2a_K python tools/asm_ab.py 100 512,620,8,1,1 > gpurun_out/r2a_ncuK.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:assemble_kernel -s 24 -c 1 -o gpurun_out/r2a_MKD python tools/asm_ab.py 100 512,620,8,1,1 > gpurun_out/r2a_ncuMKD.log 2>&1
tail -3 gpurun_out/r2a_gpu_tests.log; cat gpurun_out/r2a_k1_variants.log
