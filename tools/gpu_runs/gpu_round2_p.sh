O=gpurun_out/r02p; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_learner.py tests/test_gpu_api.py -x -q > $O/tests.log 2>&1; tail -n 3 $O/tests.log
for v in 1 0; do DRL_B200_SIDE2=$v python bench.py --steps 40 --warmup 5 --no-agent-api --no-cpu-baseline > $O/bench_side2_$v.json 2> $O/bench_side2_$v.err; python -c "
import json; d=json.load(open('$O/bench_side2_$v.json')); print('side2 $v', d['value'], d['ms_per_step'], d['e2e']['value'])"; done
timeout 300 python tools/timeline.py --math-mode 5 > $O/timeline.txt 2>&1; head -n 2 $O/timeline.txt
timeout 300 python tools/two_handles_probe.py > $O/two_handles.txt 2>&1; cat $O/two_handles.txt
