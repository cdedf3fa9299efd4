"""Collective operations. `host` works on CPU buffers over the TCP transport; `cuda`
works on device buffers through the NVLink peer-memory kernels; `algorithms` are the
old-style (construct once, run many) host classes."""
from . import algorithms, cuda, host  # noqa: F401
