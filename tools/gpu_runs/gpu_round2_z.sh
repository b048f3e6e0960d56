O=gpurun_out/r02z; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q > $O/gputests.log 2>&1 ) 2> $O/gputests.time; tail -n 3 $O/gputests.log
run() { tag=$1; shift; env "$@" python bench.py --steps 300 --warmup 10 --no-agent-api --no-cpu-baseline > $O/bench_$tag.json 2> $O/bench_$tag.err; python -c "
import json; d=json.load(open('$O/bench_$tag.json')); print('$tag', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), [k for k in d['kernels_ms'] if k[0] in ('heads_out_fwd','vtrace_losses')])"; tail -c 300 $O/bench_$tag.err; }
run default X=1
run default2 X=1
timeout 300 python tools/timeline.py --math-mode 5 > $O/timeline.txt 2>&1; head -n 2 $O/timeline.txt
