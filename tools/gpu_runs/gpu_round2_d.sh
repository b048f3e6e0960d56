mkdir -p gpurun_out/r02d
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_umma16 -c 10 -f -o gpurun_out/r02d/umma16_mode5 python tools/one_step.py --math-mode 5 --steps 1 > gpurun_out/r02d/ncu_umma16.log 2>&1
tail -3 gpurun_out/r02d/ncu_umma16.log
timeout 600 ncu --set full --clock-control none -k regex:vtrace_from_softmax -c 2 -f -o gpurun_out/r02d/vtrace python tools/one_step.py --vtrace > gpurun_out/r02d/ncu_vtrace.log 2>&1
tail -3 gpurun_out/r02d/ncu_vtrace.log
ls -la gpurun_out/r02d
