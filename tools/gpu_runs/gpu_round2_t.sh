O=gpurun_out/r02t; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_learner.py tests/test_gpu_api.py -x -q > $O/tests.log 2>&1; tail -n 3 $O/tests.log
run() { tag=$1; shift; env "$@" python bench.py --steps 300 --warmup 10 --no-agent-api --no-cpu-baseline > $O/bench_$tag.json 2> $O/bench_$tag.err; python -c "
import json; d=json.load(open('$O/bench_$tag.json')); print('$tag', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), [k for k in d['kernels_ms'] if k[0] in ('conv2_fwd','conv3_fwd','conv2_wgrad','conv3_wgrad')])"; }
run default X=1
run c2f2 DRL_B200_C2F=2
run cw8 DRL_B200_CW8=1
run c2f2_cw8 DRL_B200_C2F=2 DRL_B200_CW8=1
run side2off DRL_B200_SIDE2=0
run default2 X=1
DRL_B200_C2F=2 DRL_B200_CW8=1 timeout 600 python -m pytest tests/test_gpu_learner.py -x -q > $O/tests_8w.log 2>&1; tail -n 3 $O/tests_8w.log
