from .launch import create_device, init_context, spawn_threads  # noqa: F401
from .affinity import bind_to_gpu, gpu_local_cpus  # noqa: F401
