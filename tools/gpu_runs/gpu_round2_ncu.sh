O=gpurun_out/r02ncu; mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_umma16 -c 12 -o $O/umma16_final python tools/one_step.py --math-mode 5 --steps 1 > $O/ncu.log 2>&1; tail -n 2 $O/ncu.log | cut -c 1-200
