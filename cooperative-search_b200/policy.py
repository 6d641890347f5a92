"""Batched agents: the reference's RNN agent network (network/base_net.py:5-47, without the conv front end) and the
action choice of Agents.choose_action (agent/agent.py:33-75) for every (env, agent) row of a vectorised env at once,
in one kernel launch (csrc/policy.cu).  The reference evaluates one (1, in) row per agent per step."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import CoopSearchError

# RNN.state_dict() key -> cs_policy_weights field
_KEYMAP = {"fc1.weight": "fc1_w", "fc1.bias": "fc1_b", "rnn.weight_ih": "w_ih", "rnn.weight_hh": "w_hh",
           "rnn.bias_ih": "b_ih", "rnn.bias_hh": "b_hh", "fc2.0.weight": "fc2a_w", "fc2.0.bias": "fc2a_b",
           "fc2.2.weight": "fc2b_w", "fc2.2.bias": "fc2b_b"}


class BatchedRNNAgents:
    """``state_dict``: the reference's ``{idx}_rnn_net_params.pkl`` contents (or any mapping with the same keys).
    ``choose_actions(obs)`` = Agents.choose_action for all envs and agents; hidden states and last actions are kept on
    the device between calls (``init_hidden()`` = policy.init_hidden + rollout.py:31)."""

    def __init__(self, state_dict, num_envs, n_agents, obs_dim=4, n_actions=3, last_action=True, reuse_network=True,
                 device=None, seed=0):
        if not torch.cuda.is_available():
            raise CoopSearchError("coopsearch_b200 needs a CUDA device (B200); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.num_envs, self.n_agents, self.obs_dim, self.n_actions = int(num_envs), int(n_agents), int(obs_dim), int(n_actions)
        self.seed = int(seed) & 0xFFFFFFFF
        host = {}
        for key, field in _KEYMAP.items():
            if key not in state_dict:
                raise CoopSearchError("state_dict has no %r" % key)
            host[field] = np.ascontiguousarray(torch.as_tensor(state_dict[key]).detach().cpu().numpy(), dtype=np.float32)
        in_dim = obs_dim + (n_actions if last_action else 0) + (n_agents if reuse_network else 0)
        if host["fc1_w"].shape != (64, in_dim):
            raise CoopSearchError("fc1.weight is %s, expected (64, %d)" % (host["fc1_w"].shape, in_dim))
        cfg = _lib.PolicyCfg(struct_size=C.sizeof(_lib.PolicyCfg), device=self.device.index, n_agents=n_agents, obs_dim=obs_dim,
                             n_actions=n_actions, hidden_dim=64, last_action=int(last_action), reuse_network=int(reuse_network))
        w = _lib.PolicyWeights(**{k: v.ctypes.data for k, v in host.items()})
        hp = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cs_policy_create(C.byref(cfg), C.byref(w), C.byref(hp)), "cs_policy_create")
        self._h = hp
        rows = self.num_envs * self.n_agents
        self.hidden = torch.zeros((self.num_envs, self.n_agents, 64), dtype=torch.float32, device=self.device)
        self.q = torch.empty((self.num_envs, self.n_agents, self.n_actions), dtype=torch.float32, device=self.device)
        self.actions = torch.full((self.num_envs, self.n_agents), 255, dtype=torch.uint8, device=self.device)
        self._rows, self._t = rows, 0

    def init_hidden(self, mask=None):
        """policy.init_hidden(1) + last_action = zeros (rollout.py:31-33) for all envs, or those with mask[e] != 0."""
        if mask is None:
            self.hidden.zero_()
            self.actions.fill_(255)
        else:
            m = torch.as_tensor(mask, device=self.device).bool()
            self.hidden[m] = 0
            self.actions[m] = 255

    def choose_actions(self, obs, avail=None, epsilon=0.0, evaluate=True):
        """obs: [E,n,obs_dim] float32 device tensor (env.get_obs()); avail: [E,n,A] or None.  Returns the [E,n] uint8
        action tensor (also kept as the next call's last action); ``self.q`` holds the action values."""
        obs = obs.contiguous()
        if tuple(obs.shape) != (self.num_envs, self.n_agents, self.obs_dim) or obs.dtype != torch.float32:
            raise CoopSearchError("obs must be float32 [num_envs, n_agents, obs_dim]")
        av = None
        if avail is not None:
            av = avail.to(device=self.device, dtype=torch.uint8).contiguous()
        io = _lib.PolicyIO(rows=self._rows, evaluate=int(bool(evaluate)), epsilon=float(epsilon), seed=self.seed, t=self._t,
                           obs=obs.data_ptr(), last_action=self.actions.data_ptr(), avail=av.data_ptr() if av is not None else None,
                           hidden=self.hidden.data_ptr(), q=self.q.data_ptr(), actions=self.actions.data_ptr())
        self._t = (self._t + 1) & 0xFFFFFFFF
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cs_policy_act(self._h, C.byref(io), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                       "cs_policy_act")
        return self.actions

    def __del__(self):
        try:
            if self._h:
                self.lib.cs_policy_destroy(self._h)
                self._h = None
        except Exception:
            pass
