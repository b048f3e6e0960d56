O=gpurun_out/r02r; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_learner.py tests/test_gpu_api.py tests/test_gpu_umma.py tests/test_gpu_apex.py -x -q > $O/tests.log 2>&1; tail -n 3 $O/tests.log
run() { tag=$1; shift; env "$@" python bench.py --steps 300 --warmup 10 --no-agent-api --no-cpu-baseline > $O/bench_$tag.json 2> $O/bench_$tag.err; python -c "
import json; d=json.load(open('$O/bench_$tag.json')); print('$tag', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']))"; }
run default X=1
run wimg DRL_B200_C1_WIMG=1
run late DRL_B200_LSTMW_LATE=1
run emb0 DRL_B200_EMB_SIDE2=0
run default2 X=1
run side2off DRL_B200_SIDE2=0
run late2 DRL_B200_LSTMW_LATE=1
run emb0b DRL_B200_EMB_SIDE2=0
run ks1 DRL_B200_HEADS_KS=1
run chunks1 DRL_B200_DCOL_CHUNKS=1
run chunks2 DRL_B200_DCOL_CHUNKS=2
run chunks4 DRL_B200_DCOL_CHUNKS=4
run tc DRL_B200_HEADS_TC=1
run tc2 DRL_B200_HEADS_TC=1
DRL_B200_HEADS_TC=1 timeout 600 python -m pytest tests/test_gpu_learner.py -x -q > $O/tests_tc.log 2>&1; tail -n 3 $O/tests_tc.log
DRL_B200_HEADS_TC=1 timeout 300 python tools/timeline.py --math-mode 5 > $O/timeline_tc.txt 2>&1; head -n 2 $O/timeline_tc.txt
timeout 300 python tools/timeline.py --math-mode 5 > $O/timeline.txt 2>&1; head -n 2 $O/timeline.txt
