mkdir -p gpurun_out/r02c
CUDA_LAUNCH_BLOCKING=1 python tools/diag_mode5b.py > gpurun_out/r02c/diag_b.log 2>&1
tail -5 gpurun_out/r02c/diag_b.log
python - > gpurun_out/r02c/parity5.log 2>&1 <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import parity
for (B, T, kw) in ((4, 20, {}), (32, 20, {"layers": False}), (3, 7, {}), (4, 20, {"layers": False, "steps": 3, "use_cuda_graph": True})):
    errs = parity.compare_step(B, T=T, math_mode=5, **kw)
    bad = parity.failures(errs)
    top = sorted(((v, k) for k, v in errs.items() if not k.startswith(("kink", "lr"))), reverse=True)[:16]
    print("== step mode 5 B=%d T=%d %s: %d entries, %d failures; worst:" % (B, T, kw, len(errs), len(bad)))
    for v, k in top:
        print("   %-28s %.3e" % (k, v))
    print("   kink flip fraction %.2e max |x| at flip %.2e" % (errs["kink/flip_fraction"], errs["kink/max_abs_at_flip"]), flush=True)
PY
cat gpurun_out/r02c/parity5.log | head -60
python bench.py --steps 20 --warmup 5 --math-mode 5 --no-agent-api --no-cpu-baseline > gpurun_out/r02c/bench_mode5.json 2> gpurun_out/r02c/bench_mode5.err
tail -c 800 gpurun_out/r02c/bench_mode5.err; python -c "
import json; d=json.load(open('gpurun_out/r02c/bench_mode5.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['kernels_ms'])"
