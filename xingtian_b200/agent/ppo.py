"""Trajectory handling of the PPO agents (xt/agent/ppo/ppo.py:30-106, atari_ppo.py:31-51).

The env-step loop of xt.agent.Agent stays in the reference (CPU, out of scope).  These classes
restate only what decides the bytes the learner receives: which fields a transition carries and
what `get_trajectory` ships.  With ``device_gae=True`` (default) the numpy GAE loop of
``data_proc`` is skipped and the raw (value[T+1], reward, done) go to the learner, where
``xtb_gae`` computes adv / target_value in HBM; ``device_gae=False`` reproduces the reference
message byte for byte (host float64 GAE)."""
from collections import defaultdict

import numpy as np

from ..registry import Registers

GAMMA, LAM = 0.99, 0.95   # xt/agent/ppo/default_config.py


@Registers.agent
class PPO(object):
    def __init__(self, env=None, alg=None, agent_config=None, **kwargs):
        self.env, self.alg = env, alg
        self.agent_config = dict(agent_config or {})
        self.transition_data = defaultdict()
        self.trajectory = defaultdict(list)
        self.device_gae = bool(self.agent_config.get("device_gae", True))

    def infer_action(self, state, use_explore):
        predict_val = self.alg.predict(state)
        return self.handel_predict_value(state, predict_val)

    def handel_predict_value(self, state, predict_val):
        action, logp, value = predict_val[0][0], predict_val[1][0], predict_val[2][0]
        self.transition_data.update({"cur_state": state, "action": action, "logp": logp, "value": value})
        return action

    def handle_env_feedback(self, next_raw_state, reward, done, info, use_explore):
        self.transition_data.update({"reward": reward, "done": done, "info": info})
        return self.transition_data

    def add_to_trajectory(self, transition_data):
        for k, val in transition_data.items():
            self.trajectory[k].append(val)

    def get_trajectory(self, last_pred=None):
        self.trajectory["value"].append(last_pred[2][0])
        self.data_proc()
        return dict(self.trajectory)

    def data_proc(self):
        traj = self.trajectory
        state = np.asarray(traj["cur_state"])
        action = np.asarray(traj["action"])
        logp = np.asarray(traj["logp"])
        value = np.asarray(traj["value"])
        reward = np.asarray(traj["reward"])
        done = np.asarray(traj["done"])
        self.trajectory["cur_state"] = state
        self.trajectory["action"] = action
        self.trajectory["logp"] = logp
        if self.device_gae:
            self.trajectory["value"] = value.astype(np.float32)
            self.trajectory["reward"] = reward
            self.trajectory["done"] = done
            return
        next_value = value[1:]
        value = value[:-1]
        done = np.expand_dims(done, axis=1)
        reward = np.expand_dims(reward, axis=1)
        discount = ~done * GAMMA
        adv = reward + discount * next_value - value
        for j in range(len(adv) - 2, -1, -1):
            adv[j] += adv[j + 1] * discount[j] * LAM
        self.trajectory["adv"] = adv
        self.trajectory["old_value"] = value
        self.trajectory["target_value"] = adv + value
        del self.trajectory["value"]


@Registers.agent
class AtariPpo(PPO):
    """atari_ppo.py:31-51: rewards are sign-clipped while exploring."""

    def handle_env_feedback(self, next_raw_state, reward, done, info, use_explore):
        info.update({"eval_reward": reward})
        self.transition_data.update({"reward": np.sign(reward) if use_explore else reward, "done": done, "info": info})
        return self.transition_data
