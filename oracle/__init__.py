"""CPU oracle for the DeepHall walker-evaluation hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.  Nothing under ``deephall_b200/`` does.

It restates, in torch on the CPU (fp64 or fp32), the algorithm of the reference
(peterzjx/DeepHall, pure JAX/flax):

* ``oracle.psiformer``   <- deephall/networks/psiformer.py, networks/blocks.py
* ``oracle.laughlin``    <- deephall/networks/laughlin.py
* ``oracle.hamiltonian`` <- deephall/hamiltonian.py  (grad + full Hessian route)
* ``oracle.mcmc``        <- deephall/mcmc.py, deephall/train.py:40-54
* ``oracle.loss``        <- deephall/loss.py
* ``oracle.jets``        <- our own forward-Laplacian bookkeeping, checked
                            against ``oracle.hamiltonian`` (spec for the kernels)

Pinning status.  The reference cannot be imported here or on the GPU box (jax,
flax, kfac_jax, optax, chex and omegaconf are not installed and there is no
network), and it contains no native code to compile, so ``oracle/_ref`` does not
exist.  The oracle is pinned against every known answer the reference's own
tests hold for this path (tests/hamiltonian_test.py:42-76, tests/cli_test.py:41,
tests/train_test.py:46-48) -- see tests/test_oracle_known_answers.py.  The
third-party arithmetic (flax 0.10.2 Dense / DenseGeneral / MultiHeadAttention /
LayerNorm, jax 0.4.35 slogdet) is restated from its published semantics; the
reference holds no per-walker golden vectors for the Psiformer, so at the flax
boundary the status is: PARITY UNPINNED (structural invariants only).
"""
