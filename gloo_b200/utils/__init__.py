from .launch import create_device, init_context, spawn_threads  # noqa: F401
