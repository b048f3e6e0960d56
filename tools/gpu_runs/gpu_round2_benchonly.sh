O=gpurun_out/r02bench; mkdir -p $O
timeout 300 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; python -c "
import json; d=json.load(open('$O/bench_n1.json')); print(round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), d['gpu_launches'], d['roofline']['kernel'], round(d['roofline']['frac'],4), d['clocks'])"
