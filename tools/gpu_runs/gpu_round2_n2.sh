O=gpurun_out/r02n2; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 50 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
python -c "
import json; d=json.load(open('$O/bench_n2.json')); print('n2', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), d.get('replicas_identical'), d.get('reduce_matches_nccl'), d.get('reduce_vs_nccl_max_rel_err'))"
tail -c 300 $O/bench_n2.err
( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > $O/multi_tests.log 2>&1 ) 2> $O/multi_tests.time; tail -n 3 $O/multi_tests.log
python bench.py --steps 50 --warmup 5 --no-agent-api --no-cpu-baseline > $O/bench_n1_samebox.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/bench_n1_samebox.json')); print('n1 same box', round(d['value']), round(d['ms_per_step'],4))"
