mkdir -p gpurun_out/r02l
for pdl in 0 1 2 3; do
DRL_B200_PDL=$pdl python bench.py --steps 40 --warmup 5 --no-agent-api --no-cpu-baseline > gpurun_out/r02l/bench_pdl$pdl.json 2> gpurun_out/r02l/bench_pdl$pdl.err
python -c "
import json; d=json.load(open('gpurun_out/r02l/bench_pdl$pdl.json')); print('pdl $pdl', d['value'], d['ms_per_step'], d['e2e']['value'])"
tail -c 300 gpurun_out/r02l/bench_pdl$pdl.err
done
python bench.py --steps 20 --warmup 5 --scaling strong --no-agent-api --no-cpu-baseline > gpurun_out/r02l/bench_n1_strong.json 2> gpurun_out/r02l/bench_n1_strong.err
python -c "
import json; d=json.load(open('gpurun_out/r02l/bench_n1_strong.json')); print('strong n1', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['global_batch'])"
tail -c 300 gpurun_out/r02l/bench_n1_strong.err
