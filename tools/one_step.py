"""A couple of eager (no CUDA graph) learner steps at the bench shape, for ncu captures of individual kernels.
    ncu --set full -k regex:gemm_umma16 -c 12 -o out python tools/one_step.py --math-mode 5"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--math-mode", type=int, default=0)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--vtrace", action="store_true", help="run the stand-alone V-trace kernel at [65536,18,18] instead")
    a = ap.parse_args()
    if a.vtrace:
        import torch
        import bench
        print(bench.vtrace_roofline(torch, bench.measured_peaks(), reps=2))
        return
    from distributed_reinforcement_learning_b200.learner import NativeLearner
    from distributed_reinforcement_learning_b200.model import impala_actor_critic as model
    import bench
    b = bench.synth_batch(a.batch, 1234)
    eng = NativeLearner(batch=a.batch, trajectory=20, num_action=18, math_mode=a.math_mode, use_cuda_graph=False)
    eng.set_params(model.init_params(seed=0))
    for s in range(a.steps):
        eng.stage(s % 2, *[b[k] for k in bench.FIELDS])
        out = eng.step(s % 2)
    print(out)
    eng.close()


if __name__ == "__main__":
    main()
