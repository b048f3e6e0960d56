mkdir -p gpurun_out/r02h
( time python -m pytest tests -m gpu -q --timeout 900 ) > gpurun_out/r02h/gputests.log 2>&1
tail -8 gpurun_out/r02h/gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02h/smoke.log 2>&1; tail -3 gpurun_out/r02h/smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02h/bench_n1.json 2> gpurun_out/r02h/bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02h/bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_agent_api']['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'])"
