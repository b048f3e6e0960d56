O=gpurun_out/r02y; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_learner.py tests/test_gpu_api.py -x -q > $O/tests.log 2>&1; tail -n 5 $O/tests.log
run() { tag=$1; shift; env "$@" python bench.py --steps 300 --warmup 10 --no-agent-api --no-cpu-baseline > $O/bench_$tag.json 2> $O/bench_$tag.err; python -c "
import json; d=json.load(open('$O/bench_$tag.json')); print('$tag', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']))"; tail -c 300 $O/bench_$tag.err; }
run default X=1
run emb0 DRL_B200_EMB_SIDE2=0
run default2 X=1
timeout 300 python tools/timeline.py --math-mode 5 > $O/timeline.txt 2>&1; head -n 2 $O/timeline.txt
