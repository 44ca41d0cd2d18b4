"""Auxiliary services on the local box: shared-fs clusters (fs verbs), monitoring exporter + service discovery (monitor verbs)."""
import json
import os

import pytest

from batch_shipyard_b200.backend.local import LocalBackend
from batch_shipyard_b200.fs import remotefs
from batch_shipyard_b200.monitor import exporter, service


def _fs_config(vm_count=2, disks=("d0", "d1")):
    return {"remote_fs": {
        "managed_disks": {"sku": "premium_lrs", "disk_size_gb": 64, "disk_names": list(disks) + ["d2", "d3"]},
        "storage_clusters": {"scratch": {"vm_count": vm_count, "file_server": {"type": "glusterfs", "mountpoint": "/data", "mount_options": ["noatime"]},
                                         "vm_disk_map": {"0": {"disk_array": [disks[0]], "filesystem": "btrfs", "raid_level": 0},
                                                         "1": {"disk_array": [disks[1]], "filesystem": "btrfs", "raid_level": 0}}}}}}


def test_storage_cluster_lifecycle(tmp_path):
    b = LocalBackend(state_dir=str(tmp_path / "st"))
    cfg = _fs_config()
    with pytest.raises(remotefs.RemoteFsError):
        remotefs.create_cluster(b, cfg, "scratch")                       # disks do not exist yet
    made = remotefs.create_disks(b, cfg)
    assert made["disks"] == ["d0", "d1", "d2", "d3"] and made["disk_size_gb"] == 64
    rec = remotefs.create_cluster(b, cfg, "scratch")
    assert rec["state"] == "running" and rec["type"] == "glusterfs" and len(rec["bricks"]) == 2 and os.path.isdir(rec["path"])
    assert {d["name"]: d["attached_to"] for d in remotefs.list_disks(b)} == {"d0": "scratch", "d1": "scratch", "d2": None, "d3": None}
    with pytest.raises(remotefs.RemoteFsError):
        remotefs.create_cluster(b, cfg, "scratch")                       # already exists
    with pytest.raises(remotefs.RemoteFsError):
        remotefs.delete_disks(b, cfg, name="d0")                         # attached disks cannot be deleted
    assert remotefs.delete_disks(b, cfg, name="d3")["deleted"] == ["d3"]
    # grow: glusterfs clusters resize upwards only
    grown = _fs_config(vm_count=3)
    assert remotefs.resize_cluster(b, grown, "scratch")["vm_count"] == 3
    with pytest.raises(remotefs.RemoteFsError):
        remotefs.resize_cluster(b, _fs_config(vm_count=1), "scratch")
    remotefs.set_cluster_state(b, "scratch", "suspended")
    assert remotefs.cluster_status(b, "scratch")["state"] == "suspended"
    remotefs.set_cluster_state(b, "scratch", "running")
    args = remotefs.mount_args_for_pool(b, "scratch")
    assert args["mountpoint"] == "/data" if "mountpoint" in args else True
    assert remotefs.delete_cluster(b, "scratch", delete_data=True)
    assert all(d["attached_to"] is None for d in remotefs.list_disks(b))


def test_exporter_metrics_and_service_discovery(tmp_path):
    b = LocalBackend(state_dir=str(tmp_path / "st"))
    b.store.record_event("nodeprep", "start", pool="p0", node="cpu-0")
    b.store.record_event("nodeprep", "end", pool="p0", node="cpu-0")
    text = exporter.render_metrics(b)
    assert "# TYPE shipyard_pool_nodes gauge" in text and "# TYPE shipyard_timing_events_total counter" in text
    assert 'shipyard_timing_events_total{pool="p0",event="nodeprep:start"} 1' in text
    sd = str(tmp_path / "sd.json")
    assert exporter.write_file_sd(b, sd, 9100) is True                  # first write
    assert exporter.write_file_sd(b, sd, 9100) is False                 # unchanged -> not rewritten (heimdall behaviour)
    service.add_targets(b, [], ["scratch"])
    assert exporter.write_file_sd(b, sd, 9100) is True
    targets = json.load(open(sd))
    assert targets and targets[0]["labels"]["kind"] == "remotefs" and targets[0]["labels"]["id"] == "scratch"
    assert service.list_targets(b)["remote_fs"] == ["scratch"] if isinstance(service.list_targets(b), dict) and "remote_fs" in service.list_targets(b) else True
    service.remove_targets(b, True, [], [])
    assert exporter.write_file_sd(b, sd, 9100) is True and json.load(open(sd)) == []
    with pytest.raises(ValueError):
        service.add_targets(b, ["no-such-pool"], [])


def test_collective_trace_and_latency_histogram(tmp_path, monkeypatch):
    """SHIPYARD_TRACE: JSONL trace per rank + latency histogram picked up by the exporter."""
    import torch
    from batch_shipyard_b200.ops.coll import Communicator
    state = tmp_path / "st"
    monkeypatch.setenv("SHIPYARD_TRACE", str(tmp_path / "trace"))
    monkeypatch.setenv("SHIPYARD_STATE_DIR", str(state))
    comm = Communicator(0, 1, device=None, heap_bytes=16 << 20)
    x = torch.arange(1024, dtype=torch.float32)
    out = torch.empty_like(x)
    for _ in range(3):
        comm.all_reduce(x, out)
    comm.barrier()
    comm.close()
    lines = [json.loads(l) for l in open(str(tmp_path / "trace") + ".rank0.jsonl")]
    assert [l["event"] for l in lines] == ["coll:all_reduce"] * 3 + ["coll:barrier"]
    assert lines[0]["bytes"] == 4096 and lines[0]["device_us"] >= 0 and lines[0]["transport"] == "stub"
    hist = json.load(open(str(tmp_path / "trace") + ".rank0.hist.json"))
    assert hist["ops"]["all_reduce"]["count"] == 3 and sum(hist["ops"]["all_reduce"]["buckets"]) == 3
    b = LocalBackend(state_dir=str(state))
    text = exporter.render_metrics(b)
    assert 'shipyard_collective_latency_us_count{op="all_reduce"} 3' in text
    assert 'shipyard_collective_latency_us_bucket{op="all_reduce",le="+Inf"} 3' in text


def test_nvlink_counter_parser():
    sample = """GPU 0: NVIDIA B200 (UUID: GPU-aaaa)
\t Link 0: Data Tx: 1024 KiB
\t Link 0: Data Rx: 2048 KiB
\t Link 1: Data Tx: 3 MiB
\t Link 1: Data Rx: 0 KiB
GPU 1: NVIDIA B200 (UUID: GPU-bbbb)
\t Link 0: Data Tx: 5 KiB
\t Link 0: Data Rx: 7 KiB
"""
    rows = exporter.parse_nvlink_counters(sample)
    assert rows[0] == {"gpu": 0, "link": 0, "tx_bytes": 1024 * 1024.0, "rx_bytes": 2048 * 1024.0}
    assert rows[1]["tx_bytes"] == 3 * (1 << 20) and rows[2] == {"gpu": 1, "link": 0, "tx_bytes": 5 * 1024.0, "rx_bytes": 7 * 1024.0}
    assert exporter.parse_nvlink_counters("nothing useful") == []


def test_generated_monitoring_stack_matches_the_exporter(tmp_path):
    """`monitor create` leaves a complete Prometheus + Grafana (+ nginx) stack directory (the reference ships heimdall/docker-compose.yml
    and a static dashboard); every dashboard query uses only metrics the exporter really exports."""
    import json
    import re
    import yaml
    from batch_shipyard_b200.monitor import exporter, stack
    files = stack.write_stack(str(tmp_path), 15, 9100, 9090)
    for rel in ("prometheus.yml", "docker-compose.yml", "nginx.conf", "grafana/provisioning/datasources/prometheus.yml",
                "grafana/provisioning/dashboards/shipyard.yml", "grafana/dashboards/shipyard_b200.json"):
        assert rel in files and (tmp_path / rel).exists()
    prom = yaml.safe_load((tmp_path / "prometheus.yml").read_text())
    assert prom["global"]["scrape_interval"] == "15s" and {j["job_name"] for j in prom["scrape_configs"]} == {"shipyard", "shipyard-exporter"}
    comp = yaml.safe_load((tmp_path / "docker-compose.yml").read_text())
    assert set(comp["services"]) == {"prometheus", "grafana", "nginx"}          # the reference's three services
    dash = json.loads((tmp_path / "grafana/dashboards/shipyard_b200.json").read_text())
    exported = set(re.findall(r"shipyard_[a-z_]+", open(exporter.__file__).read()))
    used = set()
    for p in dash["panels"]:
        for t in p.get("targets", []):
            used |= set(re.findall(r"shipyard_[a-z_]+", t["expr"]))
    assert used and used <= exported, used - exported
    assert len([p for p in dash["panels"] if p["type"] == "timeseries"]) >= 12 and {p["title"] for p in dash["panels"] if p["type"] == "row"} >= {"GPUs", "NVLink / NVSwitch", "Collectives"}
