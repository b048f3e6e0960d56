mkdir -p gpurun_out/r02a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a/smi.txt 2>&1
( time python -m pytest tests -m gpu -q -x --timeout 600 ) > gpurun_out/r02a/gputests.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02a/bench_n1.json 2> gpurun_out/r02a/bench_n1.err
python bench.py --steps 20 --warmup 5 --batch 4 > gpurun_out/r02a/bench_n1_b4.json 2> gpurun_out/r02a/bench_n1_b4.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02a/bench_ref.json 2> gpurun_out/r02a/bench_ref.err
python bench.py --impl reference --steps 20 --warmup 5 --batch 4 > gpurun_out/r02a/bench_ref_b4.json 2> gpurun_out/r02a/bench_ref_b4.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_step.py --modes 2 3 > gpurun_out/r02a/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02a/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_step.py --modes 2 > gpurun_out/r02a/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02a/sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 3 python tools/sanitize_step.py --modes 2 > gpurun_out/r02a/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?" >> gpurun_out/r02a/sanitizer_synccheck.log
tail -3 gpurun_out/r02a/gputests.log; tail -c 600 gpurun_out/r02a/bench_n1.err; head -c 400 gpurun_out/r02a/bench_n1.json; tail -2 gpurun_out/r02a/sanitizer_*.log
