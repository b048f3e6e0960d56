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
  * ``oracle.tf1_shim``, ``oracle.ref_exec`` -- TF 1.14 API stand-in + loader that EXECUTE the unmodified reference files

PIN (round 2): the reference is Python over tensorflow==1.14.0 (README.md:14, Dockerfile:2), which is not installable
in this image (no wheel for CPython 3.12, no network), and it ships no tests, golden vectors or fixtures.  The
reference's OWN FILES are nevertheless executed here: ``oracle/tf1_shim/tensorflow`` implements the slice of the TF1
API those files touch as a deferred graph over torch-CPU, ``oracle/ref_exec.py`` imports the unmodified
``agent/impala.py``, ``optimizer/vtrace.py``, ``model/impala_actor_critic.py``, ``distributed_queue/buffer_queue.py``
(and the Ape-X / R2D2 / A3C files) from the checkout, and ``tests/test_oracle_refexec.py`` asserts that every restatement in
this package equals the executed reference (V-trace taps, losses, all gradients, optimizer steps with slots, variable
names / sharing, parameter_sync, queue order) to ~1e-12 in float64.  ``tests/golden/*.npz`` are written from those
executed-reference runs (``tests/golden/make_golden.py``).  What remains restated rather than executed is TF's own op
kernels (conv2d / dense / LSTMCell / scan / RMSProp / Adam / clip_by_global_norm semantics, listed in the shim's
docstring); ``tests/golden/make_golden_tf1.py`` regenerates or checks the same goldens under a real TF 1.14 where one
exists.  Further pins: the analytic known-answer cases of SURVEY.md Appendix C, an independent O(T^2) closed-form
V-trace, a plain-C restatement, autograd vs hand-derived head gradients, torch.nn.LSTMCell / torch.optim.Adam /
explicit-loop conv cross-checks.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package -- as the checker or the timed CPU baseline,
never as the product path.  Nothing under ``distributed_reinforcement_learning_b200/`` imports it.
"""
