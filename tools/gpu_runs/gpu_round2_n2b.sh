O=gpurun_out/r02n2b; mkdir -p $O
for c in 16 32 64 128 0; do
DRL_B200_PEER_EARLY_CTAS=$c timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 100 --warmup 5 --no-agent-api --no-cpu-baseline > $O/bench_n2_c$c.json 2> $O/bench_n2_c$c.err
python -c "
import json; d=json.load(open('$O/bench_n2_c$c.json')); print('early ctas $c', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), d.get('replicas_identical'), d.get('reduce_matches_nccl'))"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/timeline.py --math-mode 5 --back-to-back > $O/timeline_n2.txt 2>&1; head -n 3 $O/timeline_n2.txt
