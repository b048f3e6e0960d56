O=gpurun_out/r02last; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_learner.py tests/test_gpu_api.py tests/test_gpu_umma16.py -x -q > $O/tests.log 2>&1; tail -n 3 $O/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -n 1 $O/smoke.log
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print(round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), d['gpu_launches'])"
