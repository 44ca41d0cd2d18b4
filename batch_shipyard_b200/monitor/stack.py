"""Generated monitoring stack: Prometheus + Grafana (+ nginx) around the shipyard exporter.

What the reference ships as static files for its monitoring VM — ``heimdall/docker-compose.yml`` (prometheus, grafana and nginx
services, :1-79), the Prometheus configuration its bootstrap script writes and the 5 267-line
``heimdall/batch_shipyard_dashboard.json`` — is generated here from the exporter's own metric list, so a dashboard panel can never
reference a metric that is not exported (tests/test_services.py checks that).  ``monitor create`` writes the directory; with a
container runtime on the box ``docker compose up`` in it gives the same three-service stack the reference runs, without one the
exporter + ``prometheus.yml`` are usable by any Prometheus binary.
"""
from __future__ import annotations

import json
import os

# metric name -> (panel title, unit, legend) — the single source of truth shared with monitor.exporter
METRICS = {
    "shipyard_gpu_utilization_percent": ("GPU utilisation", "percent", "gpu {{gpu}}"),
    "shipyard_gpu_memory_used_bytes": ("HBM used", "bytes", "gpu {{gpu}}"),
    "shipyard_gpu_power_watts": ("GPU power", "watt", "gpu {{gpu}}"),
    "shipyard_gpu_sm_clock_mhz": ("SM clock", "rotmhz", "gpu {{gpu}}"),
    "shipyard_nvlink_tx_bytes_total": ("NVLink transmit", "Bps", "gpu {{gpu}}"),
    "shipyard_nvlink_rx_bytes_total": ("NVLink receive", "Bps", "gpu {{gpu}}"),
    "shipyard_pool_nodes": ("Pool nodes by state", "short", "{{pool}} {{state}}"),
    "shipyard_pool_slot_utilization_percent": ("Task slot utilisation", "percent", "{{pool}}"),
    "shipyard_job_tasks": ("Tasks by state", "short", "{{job}} {{state}}"),
    "shipyard_timing_events_total": ("Timing events (nodeprep / cascade / pull)", "short", "{{source}}:{{event}}"),
    "shipyard_collective_latency_us_bucket": ("Collective latency distribution", "µs", "{{op}} le={{le}}"),
    "shipyard_collective_latency_us_count": ("Collectives per second", "ops", "{{op}}"),
    "shipyard_collective_latency_us_sum": ("Mean collective latency", "µs", "{{op}}"),
}
_COUNTERS = {"shipyard_nvlink_tx_bytes_total", "shipyard_nvlink_rx_bytes_total", "shipyard_timing_events_total", "shipyard_collective_latency_us_count"}

ROWS = [
    ("GPUs", ["shipyard_gpu_utilization_percent", "shipyard_gpu_power_watts", "shipyard_gpu_memory_used_bytes", "shipyard_gpu_sm_clock_mhz"]),
    ("NVLink / NVSwitch", ["shipyard_nvlink_tx_bytes_total", "shipyard_nvlink_rx_bytes_total"]),
    ("Pools and jobs", ["shipyard_pool_nodes", "shipyard_pool_slot_utilization_percent", "shipyard_job_tasks", "shipyard_timing_events_total"]),
    ("Collectives", ["shipyard_collective_latency_us_count", "shipyard_collective_latency_us_sum", "shipyard_collective_latency_us_bucket"]),
]


def _expr(metric: str) -> str:
    if metric == "shipyard_collective_latency_us_sum":
        return "rate(shipyard_collective_latency_us_sum[1m]) / clamp_min(rate(shipyard_collective_latency_us_count[1m]), 1e-9)"
    if metric == "shipyard_collective_latency_us_bucket":
        return "histogram_quantile(0.99, sum by (le, op) (rate(shipyard_collective_latency_us_bucket[1m])))"
    return f"rate({metric}[1m])" if metric in _COUNTERS else metric


def dashboard() -> dict:
    panels, pid, y = [], 1, 0
    for row_title, metrics in ROWS:
        panels.append({"id": pid, "type": "row", "title": row_title, "collapsed": False, "gridPos": {"h": 1, "w": 24, "x": 0, "y": y}})
        pid += 1; y += 1
        for i, m in enumerate(metrics):
            title, unit, legend = METRICS[m]
            panels.append({"id": pid, "title": title, "type": "timeseries", "datasource": {"type": "prometheus", "uid": "shipyard-prometheus"},
                           "gridPos": {"h": 8, "w": 12, "x": (i % 2) * 12, "y": y + (i // 2) * 8},
                           "targets": [{"expr": _expr(m), "legendFormat": legend, "refId": "A"}],
                           "fieldConfig": {"defaults": {"unit": unit, "custom": {"fillOpacity": 10, "lineWidth": 1}}, "overrides": []},
                           "options": {"legend": {"displayMode": "table", "placement": "right", "calcs": ["lastNotNull", "max"]},
                                       "tooltip": {"mode": "multi"}}})
            pid += 1
        y += 8 * ((len(metrics) + 1) // 2)
    return {"title": "Shipyard B200", "uid": "shipyard-b200", "schemaVersion": 39, "version": 2, "refresh": "10s", "timezone": "browser",
            "time": {"from": "now-30m", "to": "now"}, "tags": ["shipyard", "b200"], "editable": True,
            "templating": {"list": [{"name": "pool", "type": "query", "datasource": {"type": "prometheus", "uid": "shipyard-prometheus"},
                                     "query": "label_values(shipyard_pool_nodes, pool)", "includeAll": True, "multi": True, "refresh": 2}]},
            "annotations": {"list": [{"name": "task completions", "datasource": {"type": "prometheus", "uid": "shipyard-prometheus"},
                                      "enable": True, "expr": "changes(shipyard_job_tasks{state=\"completed\"}[1m]) > 0", "iconColor": "green"}]},
            "panels": panels}


def prometheus_yml(file_sd: str, scrape_interval: int, exporter_port: int) -> str:
    return (f"global:\n  scrape_interval: {scrape_interval}s\n  evaluation_interval: {scrape_interval}s\n"
            "scrape_configs:\n"
            "  - job_name: shipyard                 # per-pool targets discovered by the exporter (file_sd rewritten only on change)\n"
            f"    file_sd_configs:\n      - files: ['{file_sd}']\n        refresh_interval: 10s\n"
            f"  - job_name: shipyard-exporter\n    static_configs:\n      - targets: ['127.0.0.1:{exporter_port}']\n")


def compose_yml(prom_port: int, grafana_port: int = 3000) -> str:
    """Three services, as /root/reference/heimdall/docker-compose.yml:1-79 (prometheus, grafana behind nginx); host networking
    because the exporter listens on the box itself."""
    return f"""# generated by `shipyard monitor create` — same service set as the reference's monitoring VM
services:
  prometheus:
    image: prom/prometheus:latest
    network_mode: host
    command: ["--config.file=/etc/prometheus/prometheus.yml", "--web.listen-address=:{prom_port}", "--storage.tsdb.retention.time=15d"]
    volumes:
      - ./prometheus.yml:/etc/prometheus/prometheus.yml:ro
      - ./file_sd.json:/etc/prometheus/file_sd.json:ro
      - prometheus-data:/prometheus
    restart: unless-stopped
  grafana:
    image: grafana/grafana:latest
    network_mode: host
    environment:
      GF_SERVER_HTTP_PORT: "{grafana_port}"
      GF_SECURITY_ADMIN_USER: ${{GF_SECURITY_ADMIN_USER:-admin}}
      GF_SECURITY_ADMIN_PASSWORD: ${{GF_SECURITY_ADMIN_PASSWORD:-admin}}
      GF_DASHBOARDS_DEFAULT_HOME_DASHBOARD_PATH: /var/lib/grafana/dashboards/shipyard_b200.json
    volumes:
      - ./grafana/provisioning:/etc/grafana/provisioning:ro
      - ./grafana/dashboards:/var/lib/grafana/dashboards:ro
      - grafana-data:/var/lib/grafana
    depends_on: [prometheus]
    restart: unless-stopped
  nginx:
    image: nginx:stable
    network_mode: host
    volumes:
      - ./nginx.conf:/etc/nginx/nginx.conf:ro
    depends_on: [grafana]
    restart: unless-stopped
volumes:
  prometheus-data: {{}}
  grafana-data: {{}}
"""


def nginx_conf(prom_port: int, grafana_port: int = 3000, listen: int = 8080) -> str:
    return (f"events {{}}\nhttp {{\n  server {{\n    listen {listen};\n"
            f"    location /grafana/ {{ proxy_pass http://127.0.0.1:{grafana_port}/; proxy_set_header Host $host; }}\n"
            f"    location /prometheus/ {{ proxy_pass http://127.0.0.1:{prom_port}/; }}\n"
            "    location / { return 302 /grafana/; }\n  }\n}\n")


def write_stack(directory: str, scrape_interval: int, exporter_port: int, prom_port: int) -> list[str]:
    """Write the whole stack directory; returns the files written (relative)."""
    files = {
        "prometheus.yml": prometheus_yml(os.path.join(directory, "file_sd.json"), scrape_interval, exporter_port),
        "docker-compose.yml": compose_yml(prom_port),
        "nginx.conf": nginx_conf(prom_port),
        "grafana/provisioning/datasources/prometheus.yml":
            f"apiVersion: 1\ndatasources:\n  - name: Prometheus\n    uid: shipyard-prometheus\n    type: prometheus\n    access: proxy\n"
            f"    url: http://127.0.0.1:{prom_port}\n    isDefault: true\n",
        "grafana/provisioning/dashboards/shipyard.yml":
            "apiVersion: 1\nproviders:\n  - name: shipyard\n    type: file\n    options:\n      path: /var/lib/grafana/dashboards\n",
        "grafana/dashboards/shipyard_b200.json": json.dumps(dashboard(), indent=1),
        "grafana_dashboard.json": json.dumps(dashboard(), indent=1),          # (flat copy kept for `monitor` users of round 1)
    }
    for rel, body in files.items():
        p = os.path.join(directory, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as f:
            f.write(body)
    sd = os.path.join(directory, "file_sd.json")
    if not os.path.exists(sd):
        with open(sd, "w") as f:
            f.write("[]\n")
    return sorted(files)
