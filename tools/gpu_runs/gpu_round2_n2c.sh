O=gpurun_out/r02n2c; mkdir -p $O
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 100 --warmup 5 --no-agent-api --no-cpu-baseline > $O/bench_n2_$tag.json 2> $O/bench_n2_$tag.err
python -c "
import json; d=json.load(open('$O/bench_n2_$tag.json')); print('$tag', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), d.get('replicas_identical'), d.get('reduce_matches_nccl'))"; }
run base X=1
run emb2 DRL_B200_EMB_SIDE2=1
run emb2_c32 DRL_B200_EMB_SIDE2=1 DRL_B200_PEER_EARLY_CTAS=32
run base2 X=1
