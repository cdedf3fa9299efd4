"""Process-wide counters of the library and an optional Prometheus endpoint for them.

The reference has a stderr logger and nothing else ("Metrics / logging"); a serving
or training job wants to scrape how many collective kernels were launched, how much went through the
same-host single-copy path, and which tuning table is in effect.

    from gloo_b200.utils import metrics
    metrics.snapshot()                       # dict
    metrics.start_exporter(9400)             # http://host:9400/metrics (needs prometheus_client)
"""
from __future__ import annotations

from typing import Dict

from .. import _C


def snapshot() -> Dict[str, float]:
    """Current values of every counter / gauge the native library keeps."""
    out: Dict[str, float] = {}
    tcp = _C.tcp_stats()
    out["glb_tcp_cma_messages_total"] = float(tcp["cma_messages"])
    out["glb_tcp_cma_bytes_total"] = float(tcp["cma_bytes"])
    out["glb_tcp_spin_budget_us"] = float(tcp["spin_us"])
    cu = getattr(_C, "cuda", None)
    if cu is not None:
        out["glb_cuda_kernel_launches_total"] = float(cu.launch_count())
    return out


def info() -> Dict[str, str]:
    cfg = {k: str(v) for k, v in _C.build_config().items()}
    cu = getattr(_C, "cuda", None)
    if cu is not None:
        try:
            cfg["tuning_table"] = str(cu.tuning_source())
        except Exception:  # noqa: BLE001
            pass
    return cfg


def start_exporter(port: int, addr: str = "127.0.0.1"):
    """Serve ``snapshot()`` as Prometheus metrics (counters end in ``_total``, the rest are gauges) plus a
    ``glb_build_info`` gauge carrying ``info()`` as labels. Returns ``(server, thread)`` from prometheus_client."""
    from prometheus_client import start_http_server
    from prometheus_client.core import REGISTRY, CounterMetricFamily, GaugeMetricFamily

    class _Collector:
        def collect(self):
            for name, value in snapshot().items():
                if name.endswith("_total"):
                    fam = CounterMetricFamily(name[:-len("_total")], "gloo_b200 counter")
                    fam.add_metric([], value)
                else:
                    fam = GaugeMetricFamily(name, "gloo_b200 gauge")
                    fam.add_metric([], value)
                yield fam
            labels = info()
            g = GaugeMetricFamily("glb_build_info", "build configuration and tuning table", labels=list(labels))
            g.add_metric(list(labels.values()), 1.0)
            yield g

    collector = _Collector()
    REGISTRY.register(collector)
    try:
        return start_http_server(port, addr=addr)
    except Exception:
        REGISTRY.unregister(collector)
        raise
