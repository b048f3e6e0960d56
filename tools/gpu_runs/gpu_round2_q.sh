O=gpurun_out/r02q; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_learner.py tests/test_gpu_api.py tests/test_gpu_umma.py -x -q > $O/tests.log 2>&1; tail -n 3 $O/tests.log
run() { tag=$1; shift; env "$@" python bench.py --steps 40 --warmup 5 --no-agent-api --no-cpu-baseline > $O/bench_$tag.json 2> $O/bench_$tag.err; python -c "
import json; d=json.load(open('$O/bench_$tag.json')); print('$tag', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), [k for k in d['kernels_ms'] if k[0].startswith('heads_l')])"; }
run default X=1
run ks1 DRL_B200_HEADS_KS=1
run late DRL_B200_LSTMW_LATE=1
run emb0 DRL_B200_EMB_SIDE2=0
run late_ks1 DRL_B200_LSTMW_LATE=1 DRL_B200_HEADS_KS=1
timeout 300 python tools/timeline.py --math-mode 5 > $O/timeline.txt 2>&1; head -n 2 $O/timeline.txt
DRL_B200_LSTMW_LATE=1 timeout 300 python tools/timeline.py --math-mode 5 > $O/timeline_late.txt 2>&1; head -n 2 $O/timeline_late.txt
timeout 300 python tools/two_handles_probe.py > $O/two_handles.txt 2>&1; cat $O/two_handles.txt
