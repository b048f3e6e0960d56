mkdir -p gpurun_out/r02i
N=${NGPU:-2}
for coll in peer nccl; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --collective $coll > gpurun_out/r02i/bench_n${N}_$coll.json 2> gpurun_out/r02i/bench_n${N}_$coll.err
python -c "
import json; d=json.load(open('gpurun_out/r02i/bench_n${N}_$coll.json')); print('$coll', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('replicas_identical'), d.get('reduce_matches_nccl'), d.get('reduce_vs_nccl_max_rel_err'))"
tail -c 300 gpurun_out/r02i/bench_n${N}_$coll.err
done
