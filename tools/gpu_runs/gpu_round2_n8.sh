O=gpurun_out/r02n8; mkdir -p $O
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 50 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err
python -c "
import json; d=json.load(open('$O/bench_n$N.json')); print('n$N', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), d.get('replicas_identical'), d.get('reduce_matches_nccl'), d.get('reduce_vs_nccl_max_rel_err'))"
tail -c 300 $O/bench_n$N.err
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 50 --warmup 5 --no-agent-api --no-cpu-baseline > $O/bench_n1_samebox_$N.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/bench_n1_samebox_$N.json')); print('n1 same box', round(d['value']), round(d['ms_per_step'],4))"
