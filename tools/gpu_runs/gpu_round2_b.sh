mkdir -p gpurun_out/r02b
python tools/diag_mode5.py > gpurun_out/r02b/diag_mode5.log 2>&1
( time python -m pytest tests/test_gpu_umma16.py tests/test_gpu_vtrace_property.py -q -x --timeout 900 ) > gpurun_out/r02b/tests_umma16.log 2>&1
python bench.py --steps 20 --warmup 5 --math-mode 5 --no-agent-api --no-cpu-baseline > gpurun_out/r02b/bench_mode5.json 2> gpurun_out/r02b/bench_mode5.err
tail -5 gpurun_out/r02b/tests_umma16.log; tail -c 1500 gpurun_out/r02b/bench_mode5.err; head -c 300 gpurun_out/r02b/bench_mode5.json
