mkdir -p gpurun_out/r02j
N=${NGPU:-2}
( time python -m pytest tests/test_gpu_multi.py -q -x --timeout 900 ) > gpurun_out/r02j/multi_tests_n$N.log 2>&1
tail -6 gpurun_out/r02j/multi_tests_n$N.log
for ctas in 64 0; do
DRL_B200_PEER_EARLY_CTAS=$ctas python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --collective peer > gpurun_out/r02j/bench_n${N}_early$ctas.json 2> gpurun_out/r02j/bench_n${N}_early$ctas.err
python -c "
import json; d=json.load(open('gpurun_out/r02j/bench_n${N}_early$ctas.json')); print('early_ctas $ctas', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('replicas_identical'), d.get('reduce_matches_nccl'), d.get('reduce_vs_nccl_max_rel_err'))"
tail -c 300 gpurun_out/r02j/bench_n${N}_early$ctas.err
done
