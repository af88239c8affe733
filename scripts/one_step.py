#!/usr/bin/env python
"""A few eager (non-graph) PPO SGD steps at B=320 for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import xingtian_b200 as xb
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = 320
info = {"actor": {"model_name": "PpoCnn", "state_dim": [84, 84, 4], "action_dim": 4, "input_dtype": "uint8",
                  "model_config": {"BATCH_SIZE": B, "ENTROPY_LOSS": 0.003, "LOSS_CLIPPING": 0.1, "LR": 0.00025, "NUM_SGD_ITER": 1,
                                   "hidden_sizes": [256], "action_type": "Categorical", "init_seed": 0, "use_cuda_graph": False}}}
alg = xb.alg_builder("PPO", info, {"instance_num": 1, "agent_num": 1})
m = alg.actor
n = B * nsteps
ro = m.rollout
ro.reserve(n)
ro.obs[:n].random_(0, 256)
ro.action[:n].random_(0, 4); ro.old_logp[:n].fill_(-1.38); ro.adv[:n].normal_(); ro.old_v[:n].normal_(); ro.target_v[:n].normal_()
ro.n = n
np.random.seed(0)
print("loss", m.train_device(n))
torch.cuda.synchronize()
