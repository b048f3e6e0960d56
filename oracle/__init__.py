"""CPU oracle for the IMPALA learner hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement of the reference algorithm
(chagmgang/distributed_reinforcement_learning @ 1890ce4):

  * ``oracle.vtrace_np``    -- NumPy restatement of ``optimizer/vtrace.py`` (all 126 lines)
  * ``oracle.impala_torch`` -- PyTorch-CPU restatement (float64 = truth, float32 = CPU baseline)
                               of ``model/impala_actor_critic.py``, ``agent/impala.py:31-100,132-148``
                               with TensorFlow 1.14 kernel semantics.
  * ``oracle.synthetic``    -- the seeded synthetic trajectories/parameters of SURVEY.md section 8(d).
  * ``oracle.vtrace_c``     -- plain-C (gcc) restatement of ``optimizer/vtrace.py`` (``oracle/c/vtrace_c.c``), a second
                               independent implementation that pins ``vtrace_np``
  * ``oracle.apex_torch``, ``oracle.per_np`` -- the Ape-X learner step and the prioritized replay memory
  * ``oracle.r2d2_torch``, ``oracle.a3c_torch`` -- the R2D2 and the A3C learner steps

PARITY UNPINNED: the reference's arithmetic lives in tensorflow==1.14.0 (README.md:14,
Dockerfile:2), which is not installable in this image (no wheel for CPython 3.12, no network),
and the reference ships no tests, golden vectors or fixtures.  The oracle is therefore pinned
only by (1) float64-NumPy vs float64/float32-torch self-consistency, (2) the analytic
known-answer cases of SURVEY.md Appendix C, (3) an independent O(T^2) closed-form V-trace,
(4) autograd vs the hand-derived head gradients, and (5) committed golden fixtures generated
by ``tests/golden/make_golden.py`` from this oracle.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package -- as the checker or the timed CPU baseline,
never as the product path.  Nothing under ``distributed_reinforcement_learning_b200/`` imports it.
"""
