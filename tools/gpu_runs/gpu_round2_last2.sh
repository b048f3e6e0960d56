O=gpurun_out/r02last2; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; tail -n 3 $O/tests.log
python bench.py --steps 300 --warmup 10 --no-agent-api --no-cpu-baseline > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print(round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), [k for k in d['kernels_ms'] if k[0]=='heads_out_fwd'])"
