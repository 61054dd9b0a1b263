"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the openai/baselines PPO2 / DQN learner hot path
(SURVEY.md section 8a).  Nothing in ``baselines_b200/`` may import this package:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and there only as the checker or as the timed
CPU baseline -- never as the thing shipped.

Parity pinning status (see DESIGN.md "Oracle"):

* ``oracle.gae``          -- pinned: checked against golden vectors produced by
  EXECUTING the reference's own ``baselines/ppo2/runner.py`` (Runner.run) in the
  build container (``oracle/gen_golden.py`` -> ``tests/golden/gae_*.npz``).
* ``oracle.segment_tree`` / ``oracle.replay`` -- pinned: known-answer vectors of
  ``baselines/common/tests/test_segment_tree.py`` plus golden traces produced by
  executing the reference's ``segment_tree.py`` / ``deepq/replay_buffer.py``.
* ``oracle.nets`` / ``oracle.ppo2`` / ``oracle.deepq`` (the TF1 graph: conv / fc /
  softmax-xent / clip_by_global_norm / Adam) -- PARITY UNPINNED at the TF boundary:
  TensorFlow 1.x (``tensorflow<2``, reference Dockerfile:14) is a third-party
  dependency that is absent from /root/reference and from this image.  The
  restatement follows the reference call sites line by line (citations in each
  docstring) and is cross-checked with float64 finite differences.  Two pieces of
  it ARE pinned by executing reference code under a TensorFlow import stub
  (``tests/golden/init_adam.npz``): ``ortho_init`` (``a2c/utils.py:20-35``, numpy +
  SVD) for every nature_cnn / mlp / head shape, and the Adam update through the
  reference's numpy statement of TF-Adam (``common/mpi_adam.py:25-42``; m, v exact).
* ``oracle.frame_stack`` and the host helpers (VecNormalize, schedules,
  explained_variance) -- pinned: outputs of the executed reference classes
  (``frame_stack_*.npz``, ``vec_normalize_trace.npz``, ``host_misc.npz``).
"""
