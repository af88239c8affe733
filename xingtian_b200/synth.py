"""Seeded synthetic rollouts of the shapes SURVEY.md 8(d) fixes (no environment, no dataset): used by bench.py and the
measurement scripts.  Product-side utility (the parity tests generate their own inputs)."""
import numpy as np


def ppo_rollout(seed, env_num, steps, state_dim=(84, 84, 4), action_dim=4):
    """Env-major [E*T, ...] PPO rollout: uint8 frames U[0,255], actions U{0..A-1}, rewards in {-1,0,1} with
    P=(.05,.9,.05), done ~ Bernoulli(1/200), values ~ N(0,1), behaviour log-probs from N(0,1) logits."""
    rng = np.random.default_rng(seed)
    n = env_num * steps
    obs = rng.integers(0, 256, size=(n,) + tuple(state_dim), dtype=np.uint8)
    action = rng.integers(0, action_dim, size=n).astype(np.int32)
    reward = rng.choice(np.array([-1.0, 0.0, 1.0]), size=n, p=[0.05, 0.9, 0.05])
    done = rng.random(n) < (1.0 / 200.0)
    value = rng.standard_normal((env_num, steps + 1, 1)).astype(np.float32)
    logits = rng.standard_normal((n, action_dim)).astype(np.float32)
    lsm = logits - np.log(np.exp(logits).sum(-1, keepdims=True))
    logp = np.take_along_axis(lsm, action[:, None].astype(np.int64), 1).astype(np.float32)
    return dict(obs=obs, action=action, reward=reward, done=done, value=value, logp=logp, logits=logits)


def replay_transitions(seed, count, state_dim=(84, 84, 4), action_dim=4):
    """`count` DQN transitions (s, a, r, s', done): frames U[0,255], sign-clipped rewards, done ~ Bernoulli(1/200)."""
    rng = np.random.default_rng(seed)
    obs = rng.integers(0, 256, size=(count,) + tuple(state_dim), dtype=np.uint8)
    nxt = rng.integers(0, 256, size=(count,) + tuple(state_dim), dtype=np.uint8)
    action = rng.integers(0, action_dim, size=count).astype(np.int32)
    reward = rng.choice(np.array([-1.0, 0.0, 1.0], np.float32), size=count, p=[0.05, 0.9, 0.05])
    done = rng.random(count) < (1.0 / 200.0)
    return dict(obs=obs, next_obs=nxt, action=action, reward=reward, done=done)
