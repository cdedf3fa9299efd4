"""Collective operations. `host` works on CPU buffers over the TCP transport;
`cuda` works on device buffers through the NVLink peer-memory kernels."""
from . import cuda, host  # noqa: F401
