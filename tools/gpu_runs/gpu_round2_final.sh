# final-state validation on one B200 (profiles/r02_final_*)
O=gpurun_out/r02final; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q > $O/gputests.log 2>&1 ) 2> $O/gputests.time; tail -n 3 $O/gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 400 $O/bench_n1.err
timeout 600 python bench.py --batch 4 > $O/bench_n1_b4.json 2> $O/bench_n1_b4.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err
for f in bench_n1 bench_n1_b4 bench_ref; do python -c "
import json; d=json.loads(open('$O/$f.json').read().strip().splitlines()[-1]); print('$f', d.get('value'), d.get('ms_per_step'), d.get('e2e',{}).get('value'), d.get('gpu_launches'), d.get('e2e_agent_api'), d.get('roofline',{}).get('kernel'), d.get('roofline',{}).get('frac'))"; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-agent-api > $O/launches.log 2>&1; tail -n 1 $O/launches.log | cut -c 1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_bulk16|Umma16CfgILi64ELi2ELi2ELi8ELi1|Umma16CfgILi64ELi4ELi1ELi8ELi0" -c 6 -o $O/new_kernels python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-agent-api > $O/ncu_new.log 2>&1; tail -n 2 $O/ncu_new.log | cut -c 1-200
timeout 300 python tools/timeline.py --math-mode 5 > $O/timeline.txt 2>&1; head -n 2 $O/timeline.txt
ls -la $O | head -30
