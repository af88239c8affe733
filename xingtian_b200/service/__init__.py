"""Centralised batched inference service."""
from .predictor import BatchedPredictor, RemotePredict, make_msg  # noqa: F401
