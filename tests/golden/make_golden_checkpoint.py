#!/usr/bin/env python
"""Generates tests/golden/checkpoint_008_outputs.npz and copies the checkpoint it was made from to tests/golden/ref_checkpoint_008.pt
(authoring container only: needs /root/reference).

The reference's own resume path (run.py:64-71) is executed: ``torch.load`` of ``ppo-dash-study/models/008_ra+rf+lshp-01/*.pt`` with the
reference's ``ppo`` package importable (001_baseline/ppo: the package the 2019 pickle names), then ``actor_critic.act(...,
deterministic=True)``, ``get_value`` and ``evaluate_actions`` on seeded observations, all on the CPU.  Modern torch needs two
compatibility shims for a torch-1.0 pickle, neither of which touches reference code: a stub for ``torch.nn.backends.thnn`` and the
``padding_mode`` attribute that today's ``Conv2d.forward`` reads.
"""
import os
import shutil
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

SRC = os.path.join(ref_loader.REF_ROOT, "ppo-dash-study/models/008_ra+rf+lshp-01/ObtRetro-reduced-frame-stack.pt")


def main():
    torch.set_num_threads(1)
    ref_loader.load_variant_b()                                    # registers the reference package under the name `ppo`
    thnn = types.ModuleType("torch.nn.backends.thnn")
    thnn._get_thnn_function_backend = lambda *a, **k: None
    sys.modules["torch.nn.backends.thnn"] = thnn
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        actor_critic, ob_rms = torch.load(SRC, map_location="cpu", weights_only=False)          # run.py:65-67
    for m in actor_critic.modules():
        if isinstance(m, torch.nn.Conv2d) and not hasattr(m, "padding_mode"):
            m.padding_mode = "zeros"
    C = actor_critic.base.main[0].weight.shape[1]
    g = torch.Generator().manual_seed(8)
    N = 6
    obs = torch.rand(N, C, 84, 84, generator=g)                  # study 008 feeds obs / 255 (NormalizeWrapper without a mean file)
    vobs = torch.zeros(N, 0)
    hxs = torch.zeros(N, 1)
    masks = torch.ones(N, 1)
    with torch.no_grad():
        value, action, logp, _ = actor_critic.act(obs, vobs, hxs, masks, deterministic=True)
        gv = actor_critic.get_value(obs, vobs, hxs, masks)
        ev, elp, ent, _ = actor_critic.evaluate_actions(obs, vobs, hxs, masks, action)
    out = os.path.join(HERE, "checkpoint_008_outputs.npz")
    np.savez_compressed(out, obs=obs.numpy(), value=value.numpy(), action=action.numpy(), logp=logp.numpy(), get_value=gv.numpy(),
                        eval_value=ev.numpy(), eval_logp=elp.numpy(), entropy=ent.numpy(),
                        sum_params=np.array([float(p.double().sum()) for p in actor_critic.parameters()]),
                        ob_rms_is_none=np.array(ob_rms is None))
    shutil.copyfile(SRC, os.path.join(HERE, "ref_checkpoint_008.pt"))
    print("wrote", out, os.path.getsize(out), "and ref_checkpoint_008.pt", os.path.getsize(SRC))


if __name__ == "__main__":
    main()
