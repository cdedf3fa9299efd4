"""Reference workloads that exercise the collectives end to end (smoke / examples)."""
from .ddp_mlp import DDPMLP, train_step  # noqa: F401
