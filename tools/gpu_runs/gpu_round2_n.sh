# 4 GPUs: the N=4 point of the strong-scaling curve + racecheck/memcheck of the 2-GPU fused exchange after the wait_peers fix
mkdir -p gpurun_out/r02n
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 5 --scaling strong > gpurun_out/r02n/bench_n4_strong.json 2> gpurun_out/r02n/bench_n4_strong.err
python -c "
import json; d=json.load(open('gpurun_out/r02n/bench_n4_strong.json')); print('strong n4', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['global_batch'], d.get('replicas_identical'), d.get('reduce_matches_nccl'))"
tail -c 300 gpurun_out/r02n/bench_n4_strong.err
export DRL_B200_PEER_TIMEOUT_S=120
for tool in racecheck memcheck; do
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 --no-python \
  compute-sanitizer --tool $tool --log-file gpurun_out/r02n/sanitizer_peer_${tool}_rank%q{RANK}.log \
  python tools/sanitize_step.py --peer --modes 5 --steps 3 > gpurun_out/r02n/sanitizer_peer_$tool.out 2>&1
echo "$tool rc=$?"; tail -n 3 gpurun_out/r02n/sanitizer_peer_$tool.out; tail -n 2 gpurun_out/r02n/sanitizer_peer_${tool}_rank*.log
done
