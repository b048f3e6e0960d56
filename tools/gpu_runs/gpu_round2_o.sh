# final-state validation on one B200: full GPU suite, smoke, the driver's bench lines, ncu launch list + full capture of
# the conv1 TMA kernels (profiles/r02_*)
O=gpurun_out/r02o; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q > $O/gputests.log 2>&1 ) 2> $O/gputests.time; tail -n 3 $O/gputests.log; tail -n 4 $O/gputests.time
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 400 $O/bench_n1.err
timeout 600 python bench.py --batch 4 > $O/bench_n1_b4.json 2> $O/bench_n1_b4.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 600 python bench.py --impl reference --batch 4 --steps 5 --warmup 3 > $O/bench_ref_b4.json 2> $O/bench_ref_b4.err
for f in bench_n1 bench_n1_b4 bench_ref bench_ref_b4; do python -c "
import json; d=json.loads(open('$O/$f.json').read().strip().splitlines()[-1]); print('$f', d.get('value'), d.get('ms_per_step'), d.get('e2e',{}).get('value'), d.get('gpu_launches'), d.get('clocks'))"; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-agent-api > $O/launches.log 2>&1; tail -n 2 $O/launches.log | cut -c 1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv1_fwd_tma|conv1_wgrad_tma" -c 2 -o $O/conv1_tma python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-agent-api > $O/ncu_conv1.log 2>&1; tail -n 2 $O/ncu_conv1.log | cut -c 1-200
timeout 300 python tools/timeline.py --math-mode 5 > $O/timeline.txt 2>&1; head -n 3 $O/timeline.txt
ls -la $O
for c in 1 0; do DRL_B200_C2F=$c python bench.py --steps 40 --warmup 5 --no-agent-api --no-cpu-baseline > $O/bench_c2f$c.json 2> $O/bench_c2f$c.err; python -c "
import json; d=json.load(open('$O/bench_c2f$c.json')); print('c2f $c', d['value'], d['ms_per_step'], [k for k in d['kernels_ms'] if k[0]=='conv2_fwd'])"; done
