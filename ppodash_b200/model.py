"""Policy / CNNBase / NNBase / Categorical: drop-in for PKG/model.py and PKG/distributions.py.

Same constructor signatures, properties and ``state_dict`` keys as the reference
(``base.main.{0,2,4,7}.*``, ``base.gru.*``, ``base.critic_linear.*``, ``dist.linear.*``) and the
same initialisers consumed in the same order, so ``torch.manual_seed(s); Policy(...)`` produces
the reference's weights.  The modules are parameter containers: every forward (and, inside
``PPO.update``, backward) pass is executed by ``engine.PolicyEngine`` with the sm_100a kernels of
libppodash_b200.so.  There is no PyTorch/cuDNN fallback: calling ``act`` / ``get_value`` /
``evaluate_actions`` with the model on the CPU raises.

Out of scope: MLPBase (broken in the reference fork,
PKG/model.py:225-226), DiagGaussian / Bernoulli heads (Obstacle Tower is Discrete only).
"""
import torch
import torch.nn as nn

from . import engine as _engine


def init(module, weight_init, bias_init, gain=1):
    """PKG/utils.py:53-56."""
    weight_init(module.weight.data, gain=gain)
    bias_init(module.bias.data)
    return module


class Flatten(nn.Module):
    def forward(self, x):
        return x.view(x.size(0), -1)


class FixedCategorical:
    """Result of the Categorical head (PKG/distributions.py:18-27): holds the logits and offers
    ``sample`` / ``mode`` / ``log_probs`` / ``entropy`` / ``probs`` with the reference's shapes."""

    def __init__(self, logits, ld=None):
        self._z = logits            # [B, ld] device tensor; first A columns are logits
        self.num_actions = logits.shape[1] if ld is None else ld
        self._cache = None

    def _eval(self, actions=None):
        return _engine.categorical_eval(self._z, self.num_actions, actions)

    @property
    def probs(self):
        return self._eval()["probs"]

    @property
    def logits(self):
        z = self._z[:, :self.num_actions]
        return z - torch.logsumexp(z, dim=-1, keepdim=True)

    def mode(self):
        return self._eval()["mode"].unsqueeze(-1)

    def sample(self):
        # torch.multinomial on the kernel's probabilities: the same device RNG call
        # torch.distributions.Categorical.sample makes (not bit-reproducible across devices)
        return torch.multinomial(self.probs, 1, True)

    def log_probs(self, actions):
        return self._eval(actions.reshape(-1))["logp"].unsqueeze(-1)

    def entropy(self):
        return self._eval()["entropy"]


class Categorical(nn.Module):
    """PKG/distributions.py:54-68: Linear(num_inputs, num_outputs), orthogonal gain 0.01, zero bias."""

    def __init__(self, num_inputs, num_outputs):
        super().__init__()
        self.linear = init(nn.Linear(num_inputs, num_outputs), nn.init.orthogonal_,
                           lambda x: nn.init.constant_(x, 0), gain=0.01)

    def forward(self, x):
        raise RuntimeError("the Categorical head is evaluated by PolicyEngine (fused with the critic head)")


class NNBase(nn.Module):
    """PKG/model.py:82-109."""

    def __init__(self, recurrent, recurrent_input_size, hidden_size):
        super().__init__()
        self._hidden_size = hidden_size
        self._recurrent = recurrent
        if recurrent:
            self.gru = nn.GRU(recurrent_input_size, hidden_size)
            for name, param in self.gru.named_parameters():
                if 'bias' in name:
                    nn.init.constant_(param, 0)
                elif 'weight' in name:
                    nn.init.orthogonal_(param)

    @property
    def is_recurrent(self):
        return self._recurrent

    @property
    def recurrent_hidden_state_size(self):
        return self._hidden_size if self._recurrent else 1

    @property
    def output_size(self):
        return self._hidden_size


class CNNBase(NNBase):
    """PKG/model.py:169-199: conv 8x8/4 -> 4x4/2 -> 3x3/1 (32, 64, 32 channels), FC 1568 -> hidden,
    optional GRU(hidden + vector_obs_len -> hidden), critic head."""

    def __init__(self, num_inputs, vector_obs_len=0, recurrent=False, hidden_size=512):
        super().__init__(recurrent, hidden_size + vector_obs_len, hidden_size)
        if not recurrent and vector_obs_len != 0:
            # variant A concatenates the vector obs (model.py:195) but sizes the Categorical head for
            # `hidden_size` inputs (model.py:32); variant B drops the concat but sizes the critic for
            # hidden+V (S001/ppo/model.py:188).  Neither runs, so there is nothing to be parity with.
            raise NotImplementedError("feed-forward CNNBase with vector observations is shape-inconsistent "
                                      "in the reference (SURVEY.md 'Key structural fact'); use recurrent=True")
        self.num_inputs = num_inputs
        self.vector_obs_len = vector_obs_len
        relu_gain = nn.init.calculate_gain('relu')
        init_ = lambda m: init(m, nn.init.orthogonal_, lambda x: nn.init.constant_(x, 0), relu_gain)
        self.main = nn.Sequential(
            init_(nn.Conv2d(num_inputs, 32, 8, stride=4)), nn.ReLU(),
            init_(nn.Conv2d(32, 64, 4, stride=2)), nn.ReLU(),
            init_(nn.Conv2d(64, 32, 3, stride=1)), nn.ReLU(), Flatten(),
            init_(nn.Linear(32 * 7 * 7, hidden_size)), nn.ReLU())
        init_ = lambda m: init(m, nn.init.orthogonal_, lambda x: nn.init.constant_(x, 0))
        self.critic_linear = init_(nn.Linear(hidden_size, 1))
        self.train()

    def forward(self, visual_inputs, vector_inputs, rnn_hxs, masks):
        raise RuntimeError("CNNBase is evaluated through Policy (PolicyEngine); call Policy.act / "
                           "get_value / evaluate_actions")


class Policy(nn.Module):
    """PKG/model.py:15-79."""

    def __init__(self, obs_shape, action_space, base=None, base_kwargs=None, vector_obs_len=0):
        super().__init__()
        if base_kwargs is None:
            base_kwargs = {}
        if base is None:
            if len(obs_shape) == 3:
                base = CNNBase
            else:
                raise NotImplementedError("only image observations (CNNBase) are supported; MLPBase is broken "
                                          "in the reference fork (PKG/model.py:225-226)")
        if base is not CNNBase and not (isinstance(base, type) and issubclass(base, CNNBase)):
            raise NotImplementedError("only CNNBase is supported")
        self.obs_shape = tuple(obs_shape)
        self.base = base(obs_shape[0], vector_obs_len, **base_kwargs)
        if action_space.__class__.__name__ == "Discrete":
            self.dist = Categorical(self.base.output_size, action_space.n)
        else:
            raise NotImplementedError("only Discrete action spaces are supported (Obstacle Tower, SURVEY.md 2 #4)")
        self.num_actions = action_space.n
        self._engine = None

    # the engine holds device buffers: keep it out of pickles / deepcopies (torch.save(actor_critic), run.py:259)
    def __getstate__(self):
        d = self.__dict__.copy()
        d["_engine"] = None
        return d

    @property
    def is_recurrent(self):
        return self.base.is_recurrent

    @property
    def recurrent_hidden_state_size(self):
        """Size of rnn_hx."""
        return self.base.recurrent_hidden_state_size

    def engine(self, precision=None):
        if self._engine is None:
            self._engine = _engine.PolicyEngine(self)
        if precision is not None:
            self._engine.set_precision(precision)
        return self._engine

    def forward(self, visual_inputs, vector_inputs, rnn_hxs, masks):
        raise NotImplementedError

    def act(self, visual_inputs, vector_inputs, rnn_hxs, masks, deterministic=False):
        out = self.engine().forward(visual_inputs, vector_inputs, rnn_hxs, masks)
        dist = FixedCategorical(out["z"], self.num_actions)
        action = dist.mode() if deterministic else dist.sample()
        action_log_probs = dist.log_probs(action)
        return out["value"], action, action_log_probs, out["rnn_hxs"]

    def get_value(self, visual_inputs, vector_inputs, rnn_hxs, masks):
        return self.engine().forward(visual_inputs, vector_inputs, rnn_hxs, masks)["value"]

    def evaluate_actions(self, visual_inputs, vector_inputs, rnn_hxs, masks, action):
        """Forward-only evaluation (values are not attached to an autograd graph: PPO.update runs its
        own fused backward).  Returns (value [B,1], log_prob [B,1], mean entropy, rnn_hxs)."""
        out = self.engine().forward(visual_inputs, vector_inputs, rnn_hxs, masks)
        ev = _engine.categorical_eval(out["z"], self.num_actions, action.reshape(-1))
        return out["value"], ev["logp"].unsqueeze(-1), ev["entropy"].mean(), out["rnn_hxs"]
