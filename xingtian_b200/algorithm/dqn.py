"""DQN algorithm (xt/algorithm/dqn/dqn.py:36-148)."""
import numpy as np
import torch

from ..registry import Registers, import_config, model_builder
from .base import Algorithm
from .replay_buffer import DeviceReplayBuffer

# xt/algorithm/dqn/default_config.py
BATCH_SIZE = 32
BUFFER_SIZE = 100000
TARGET_UPDATE_FREQ = 1000
GAMMA = 0.99


@Registers.algorithm
class DQN(Algorithm):
    """Online + target network, device replay ring, fused TD-target/mse/Adam step."""

    def __init__(self, model_info, alg_config, **kwargs):
        import_config(globals(), alg_config)
        model_info = model_info["actor"]
        super().__init__(alg_name="dqn", model_info=model_info, alg_config=alg_config)
        self.target_actor = model_builder(model_info)
        self.update_target()     # keras builds the target with its own init; synced at first update in the
        # reference -- here both start equal, which only changes the first TARGET_UPDATE_FREQ steps
        obs_dt = torch.uint8 if self.actor.arch["input_dtype"] == "uint8" else torch.float32
        self.buff = DeviceReplayBuffer(BUFFER_SIZE, self.actor.state_dim, obs_dt, self.actor.device)
        self.double_dqn = alg_config.get("double_dqn", False)
        self._loss = torch.zeros(1, dtype=torch.float32, device=self.actor.device)

    def train(self, **kwargs):
        """dqn.py:61-103."""
        idx = self.buff.sample_indices(BATCH_SIZE)
        loss = self.train_on_indices(idx)
        return loss

    def train_on_indices(self, idx):
        obs, act, rew, nxt, done = self.buff.gather(idx)
        n = int(len(idx))
        self.actor.train_td_device(self.target_actor, obs, act, rew, nxt, done, n, GAMMA, self._loss,
                                   double_dqn=self.double_dqn)
        self.train_count += 1
        if self.train_count % TARGET_UPDATE_FREQ == 0:
            self.update_target()
        return float(self._loss.cpu()[0])

    def restore(self, model_name=None, model_weights=None):
        """dqn.py:105-119."""
        if model_weights is not None:
            self.actor.set_weights(model_weights)
            self.target_actor.set_weights(model_weights)
        else:
            self.actor.load_model(model_name)
            self.target_actor.load_model(model_name)

    def prepare_data(self, train_data, **kwargs):
        """dqn.py:121-138."""
        self.buff.add_batch(np.asarray(train_data["cur_state"]), np.asarray(train_data["action"]),
                            np.asarray(train_data["reward"]), np.asarray(train_data["next_state"]),
                            np.asarray(train_data["done"]))

    def update_target(self):
        """dqn.py:140-148: hard copy (device to device)."""
        self.target_actor.net.load_flat(self.actor.net.params)
