mkdir -p gpurun_out/r02k
N=8
( time python -m pytest tests/test_gpu_multi.py -q -x --timeout 900 -k "fused_peer_exchange" ) > gpurun_out/r02k/multi_tests_n$N.log 2>&1
tail -5 gpurun_out/r02k/multi_tests_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02k/bench_n8.json 2> gpurun_out/r02k/bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 1000 --warmup 5 > gpurun_out/r02k/bench_n8_soak1000.json 2> gpurun_out/r02k/bench_n8_soak1000.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02k/bench_n4.json 2> gpurun_out/r02k/bench_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 --scaling strong > gpurun_out/r02k/bench_n8_strong.json 2> gpurun_out/r02k/bench_n8_strong.err
for f in bench_n8 bench_n8_soak1000 bench_n4 bench_n8_strong; do python -c "
import json; d=json.load(open('gpurun_out/r02k/$f.json')); print('$f', d['n_gpus'], d['steps'], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('replicas_identical'), d.get('reduce_matches_nccl'), d.get('reduce_vs_nccl_max_rel_err'), d['config']['global_batch'])"; tail -c 200 gpurun_out/r02k/$f.err; done
