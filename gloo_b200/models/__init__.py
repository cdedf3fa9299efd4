"""Reference workloads that exercise the collectives end to end (smoke / examples)."""
from .ddp_mlp import DDPMLP, train_step  # noqa: F401
from .tp_transformer import DenseBlock, TPBlock  # noqa: F401,E402
