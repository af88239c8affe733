"""Agent-side helpers (mirror of the trajectory post-processing in xt/agent/ppo)."""
from .ppo import PPO, AtariPpo  # noqa: F401
