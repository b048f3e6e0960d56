# 2 GPUs: compute-sanitizer on the fused peer exchange + the N=2 point of the strong-scaling curve
mkdir -p gpurun_out/r02m
export DRL_B200_PEER_TIMEOUT_S=120
for tool in memcheck racecheck; do
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 --no-python \
  compute-sanitizer --tool $tool --log-file gpurun_out/r02m/sanitizer_peer_${tool}_rank%q{RANK}.log \
  python tools/sanitize_step.py --peer --modes 5 --steps 3 > gpurun_out/r02m/sanitizer_peer_$tool.out 2>&1
echo "$tool rc=$?"; tail -3 gpurun_out/r02m/sanitizer_peer_$tool.out; tail -2 gpurun_out/r02m/sanitizer_peer_${tool}_rank*.log
done
unset DRL_B200_PEER_TIMEOUT_S
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 5 --scaling strong > gpurun_out/r02m/bench_n2_strong.json 2> gpurun_out/r02m/bench_n2_strong.err
python -c "
import json; d=json.load(open('gpurun_out/r02m/bench_n2_strong.json')); print('strong n2', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['global_batch'], d.get('multi_gpu'))"
tail -c 300 gpurun_out/r02m/bench_n2_strong.err
