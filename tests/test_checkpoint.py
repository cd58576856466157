"""Checkpoint interchange with the reference (run.py:64-71,251-262): a real 2019 checkpoint of study 008
(tests/golden/ref_checkpoint_008.pt, copied by tests/golden/make_golden_checkpoint.py) is read without the reference package, and the
policy it yields reproduces what the reference's own module computed on the CPU (tests/golden/checkpoint_008_outputs.npz)."""
import os

import numpy as np
import pytest
import torch

from ppodash_b200 import checkpoint as ck

HERE = os.path.dirname(os.path.abspath(__file__))
CKPT = os.path.join(HERE, "golden", "ref_checkpoint_008.pt")


def test_reference_checkpoint_loads_without_reference_package(golden):
    g = golden("checkpoint_008_outputs")
    pol, ob_rms = ck.load_reference_checkpoint(CKPT)
    assert ob_rms is None and bool(g["ob_rms_is_none"])
    assert pol.obs_shape == (4, 84, 84) and pol.num_actions == 8 and not pol.is_recurrent and pol.base.vector_obs_len == 0
    got = np.array([float(p.detach().double().sum()) for p in pol.parameters()])
    np.testing.assert_array_equal(got, g["sum_params"])              # same tensors, same order (state_dict keys of the reference)
    sd = ck.reference_state_dict(pol)
    assert list(sd)[:2] == ["base.main.0.weight", "base.main.0.bias"] and sd["dist.linear.weight"].shape == (8, 512)


def test_reference_checkpoints_all_load_when_reference_is_present():
    import glob
    files = sorted(glob.glob("/root/reference/ppo-dash-study/models/*/*.pt"))
    if not files:
        pytest.skip("authoring container only (/root/reference)")
    for f in files:
        pol, _ = ck.load_reference_checkpoint(f)
        assert pol.obs_shape[1:] == (84, 84)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
def test_reference_checkpoint_policy_matches_reference_outputs(golden, precision, tmp_path):
    g = golden("checkpoint_008_outputs")
    pol, _ = ck.load_reference_checkpoint(CKPT)
    pol = pol.to("cuda:0")
    pol.engine(precision)
    d = lambda k: torch.as_tensor(g[k]).to("cuda:0")
    N = g["obs"].shape[0]
    vobs, hxs, masks = torch.zeros(N, 0, device="cuda:0"), torch.zeros(N, 1, device="cuda:0"), torch.ones(N, 1, device="cuda:0")
    value, action, logp, _ = pol.act(d("obs"), vobs, hxs, masks, deterministic=True)
    tol = dict(rtol=1e-5, atol=1e-5)          # trained weights: values ~ 1, logits ~ 10; 1e-5 relative + 1e-5 absolute
    assert torch.equal(action.cpu(), torch.as_tensor(g["action"]))
    np.testing.assert_allclose(value.cpu().numpy(), g["value"], **tol)
    np.testing.assert_allclose(logp.cpu().numpy(), g["logp"], **tol)
    np.testing.assert_allclose(pol.get_value(d("obs"), vobs, hxs, masks).cpu().numpy(), g["get_value"], **tol)
    ev, elp, ent, _ = pol.evaluate_actions(d("obs"), vobs, hxs, masks, action)
    np.testing.assert_allclose(elp.cpu().numpy(), g["eval_logp"], **tol)
    np.testing.assert_allclose(ent.item(), float(g["entropy"]), **tol)
    # save as run.py:259-262 does (+ optimiser state), load back, same outputs; the optimiser state survives
    import ppodash_b200 as ppd
    agent = ppd.algo.PPO(pol, 0.1, 1, 1, 0.5, 0.01, lr=1e-4, eps=1e-5, max_grad_norm=0.5)
    path = str(tmp_path / "ObtRetro.pt")
    ck.save_checkpoint(path, pol, None, agent.optimizer)
    pol2, rms2, opt2 = ck.load_checkpoint(path)
    assert rms2 is None and opt2 is not None and "exp_avg" in opt2
    agent.optimizer.load_state_dict(opt2)                 # layout signature matches
    bad = dict(opt2, layout=opt2["layout"][:-1])
    with pytest.raises(ValueError):
        agent.optimizer.load_state_dict(bad)
    pol2 = pol2.to("cuda:0")
    pol2.engine(precision)
    assert torch.equal(pol2.get_value(d("obs"), vobs, hxs, masks), pol.get_value(d("obs"), vobs, hxs, masks))
