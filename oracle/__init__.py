"""CPU oracle for the CACTO learning hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (NumPy fp64 + torch-CPU autograd), the algorithms of
the reference's hot path so that the CUDA path can be checked
against it.  Nothing under ``cacto_b200/`` imports it; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may, and there only as the checker / the timed CPU arm.

Pinning status (see DESIGN.md "Oracle"):

* ``oracle.per`` (segment tree + replay buffers): PINNED against the reference's own
  ``segment_tree.py`` / ``replay_buffer.ReplayBuffer`` executed in the build container
  (fixtures in ``tests/golden/per_*.npz`` made by ``tests/golden/make_golden.py``).
* ``oracle.rtg`` (RL_Solve reward-to-go): PINNED against ``RL.RL_AC.RL_Solve`` executed
  in the build container with a stub ``tensorflow`` module (same script).
* ``oracle.systems`` analytic systems (single integrator, car, car_park) and all reward
  functions: PINNED against ``environment.py`` executed with stub ``tensorflow`` /
  ``pinocchio`` modules (same script).
* ``oracle.robots`` / Pinocchio-backed systems (double integrator, manipulator, ur5):
  PARITY UNPINNED -- Pinocchio is not installable here and the reference ships no
  vectors.  Validated three independent ways instead (Lagrangian closed form, kinetic
  energy Jacobians, finite differences; tests/test_oracle_robots.py).
* ``oracle.nn`` (Keras/TF-2.11 semantics: Dense, LeakyReLU(0.3), SIREN, MSE with sample
  weights, nested GradientTape, Adam eps=1e-7): PARITY UNPINNED for losses/gradients --
  TensorFlow is absent; only layer shapes / initialiser ranges are pinned by the
  reference's archived ``.h5`` weight files.
"""
