"""CPU / NUMA placement for one-process-per-GPU jobs.

Pinned host buffers are first-touched by the allocating thread, so a rank that runs on the
socket far from its GPU puts every H2D / D2H byte across the inter-socket link. The kernel
publishes, per PCI function, the CPUs that are local to it; binding the rank to those before
it allocates keeps the staging memory (and the transport's I/O thread) next to the GPU.
The reference leaves placement to the launcher (docs/readme: `numactl`); this is that step.
"""
from __future__ import annotations

import os
from typing import List, Optional


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-")
            cpus.extend(range(int(lo), int(hi) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_pci_address(device: int) -> Optional[str]:
    import torch

    p = torch.cuda.get_device_properties(device)
    try:
        return f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except AttributeError:
        return None


def gpu_local_cpus(device: int) -> List[int]:
    """CPUs on the NUMA node of ``device`` (empty when the platform does not say)."""
    addr = gpu_pci_address(device)
    if addr is None:
        return []
    try:
        with open(f"/sys/bus/pci/devices/{addr}/local_cpulist") as f:
            return _parse_cpulist(f.read())
    except OSError:
        return []


def bind_to_gpu(device: int) -> List[int]:
    """Restrict this process to the CPUs local to ``device``. Returns the CPU list that is in
    effect afterwards; never raises (a cpuset that excludes those CPUs, or a single-node
    machine that reports every CPU, simply leaves the affinity as it was)."""
    try:
        allowed = os.sched_getaffinity(0)
        local = [c for c in gpu_local_cpus(device) if c in allowed]
        if local and len(local) < len(allowed):
            os.sched_setaffinity(0, local)
        return sorted(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001 - placement is best effort by design
        return []
