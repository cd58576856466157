"""oracle/ -- CPU restatement of the PPO-Dash training hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain numpy / torch-CPU, the algorithms of the
reference files listed below.  It exists so that the CUDA product path in
``ppodash_b200`` can be checked on a machine that does not have
``/root/reference`` (the GPU box).  Nothing in ``ppodash_b200`` imports it;
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may.

Reference files restated (PKG = ppo-dash-training/pytorch-a2c-ppo-acktr-gail/a2c_ppo_acktr):
  PKG/storage.py:82-121   compute_returns           -> oracle/returns.py
  PKG/algo/ppo.py:35-37   advantage normalisation   -> oracle/returns.py
  PKG/storage.py:123-160  feed_forward_generator    -> oracle/minibatch.py
  PKG/storage.py:162-223  recurrent_generator       -> oracle/minibatch.py
  PKG/model.py:54-199     Policy / CNNBase / GRU    -> oracle/policy.py
  PKG/distributions.py    Categorical               -> oracle/policy.py
  PKG/algo/ppo.py:34-96   PPO.update                -> oracle/ppo_update.py
  PKG/envs.py:208-217 + openai/baselines RunningMeanStd (un-vendored, unpinned:
      ppo-dash-training/environment.yml:18)         -> oracle/running_mean_std.py

Pinning: the reference has no tests and no golden vectors, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: ``tests/golden/
make_golden.py`` imports the unmodified reference from /root/reference (with a
stub for its gym/baselines-dependent ``envs`` module), runs it on seeded
synthetic rollouts and commits the results as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every oracle function against them.
The one exception is ``running_mean_std.py``: openai/baselines is not in the
reference tree and not installable here, so that restatement is
**parity unpinned** (checked only against the pooled-moments property).
"""
