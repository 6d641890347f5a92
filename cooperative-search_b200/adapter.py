"""E=1 adapter: makes a batched env look exactly like one reference env instance, so the reference's
``common/rollout.py:RolloutWorker`` (generate_episode :22-141, generate_replay :143-204) runs on it
unmodified: numpy float64 arrays of the reference shapes, Python scalars, fresh copies."""
import numpy as np
import torch

from ._lib import CoopSearchError


class SingleEnvAdapter:
    def __init__(self, vec_env):
        if vec_env.num_envs != 1:
            raise CoopSearchError("SingleEnvAdapter wraps a batched env created with num_envs=1")
        self.vec = vec_env
        self.n_agents = vec_env.n_agents

    def get_env_info(self):
        info = dict(self.vec.get_env_info())
        info.pop("n_envs", None)
        return info

    def reset(self, init=False):
        self.vec.reset(init=init)

    def get_obs(self):
        return self.vec.get_obs()[0].to(torch.float64).cpu().numpy()

    def get_state(self):
        return self.vec.get_state()[0].to(torch.float64).cpu().numpy()

    def get_avail_agent_actions(self, agent_id):
        return self.vec.get_avail_agent_actions(agent_id)[0].to(torch.float64).cpu().numpy()

    def step(self, act_list):
        if len(act_list) != self.n_agents:
            raise CoopSearchError('Act num mismatch agent')
        acts = np.array([int(a) for a in act_list], dtype=np.uint8).reshape(1, self.n_agents)   # ints, np ints, 0-d tensors
        out = self.vec.step(acts)
        reward = float(out[0][0].item())
        terminated = bool(out[1][0].item())
        info = out[2] if isinstance(out[2], str) else bool(out[2][0].item())     # win flag (flight) / '' (search)
        return reward, terminated, info

    @property
    def target_find(self):
        return int(self.vec.target_find[0].item())

    def close(self):
        self.vec.close()

    def render(self):
        pass
