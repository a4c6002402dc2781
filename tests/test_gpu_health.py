"""resources/gpu_health.py: policy and NVML plumbing against a fake NVML module."""
from __future__ import annotations

import sys
import types

from infomesh_b200.resources import gpu_health as GH


def test_policy_marks_ecc_xid_and_lost_devices_unhealthy():
    ok = GH.GpuHealth(index=0, available=True)
    assert GH.assess(ok)
    assert not GH.assess(GH.GpuHealth(index=0, available=True, ecc_uncorrected=1))
    assert not GH.assess(GH.GpuHealth(index=0, available=True, row_remap_failure=True))
    assert not GH.assess(GH.GpuHealth(index=0, available=True, xid_events=2, last_xid=79))
    assert GH.assess(GH.GpuHealth(index=0, available=True, throttle_reasons=["sw_power_cap"], ecc_corrected=5))
    assert not GH.assess(GH.GpuHealth(index=0, available=False, error="lost: GPU is lost"))
    assert GH.assess(GH.GpuHealth(index=0, available=False, error="nvml unavailable"))


def _fake_nvml(state):
    m = types.ModuleType("pynvml")
    m.NVML_MEMORY_ERROR_TYPE_UNCORRECTED, m.NVML_MEMORY_ERROR_TYPE_CORRECTED, m.NVML_VOLATILE_ECC, m.NVML_TEMPERATURE_GPU = 1, 0, 0, 0
    m.nvmlEventTypeXidCriticalError, m.nvmlEventTypeDoubleBitEccError = 8, 2
    m.nvmlInit = lambda: None
    m.nvmlShutdown = lambda: None
    m.nvmlDeviceGetCount = lambda: 2
    m.nvmlDeviceGetHandleByIndex = lambda i: ("h", i)
    m.nvmlDeviceGetIndex = lambda h: h[1]
    m.nvmlDeviceGetName = lambda h: b"NVIDIA B200"
    m.nvmlDeviceGetTotalEccErrors = lambda h, kind, counter: state["ecc"].get((h[1], kind), 0)
    m.nvmlDeviceGetRetiredPagesPendingStatus = lambda h: 0
    m.nvmlDeviceGetRemappedRows = lambda h: (0, 0, 0, 0)
    m.nvmlDeviceGetCurrentClocksThrottleReasons = lambda h: state["reasons"].get(h[1], 0)
    m.nvmlDeviceGetTemperature = lambda h, s: 55
    m.nvmlDeviceGetMemoryInfo = lambda h: types.SimpleNamespace(used=3 * 2 ** 30)
    m.nvmlEventSetCreate = lambda: "set"
    m.nvmlDeviceRegisterEvents = lambda h, mask, s: None
    m.nvmlEventSetFree = lambda s: None

    def wait(_set, _timeout):
        if state["events"]:
            return state["events"].pop(0)
        raise RuntimeError("timeout")

    m.nvmlEventSetWait_v2 = wait
    return m


def test_monitor_reads_counters_and_xid_events(monkeypatch):
    state = {"ecc": {}, "reasons": {1: 0x4 | 0x40}, "events": []}
    monkeypatch.setitem(sys.modules, "pynvml", _fake_nvml(state))
    mon = GH.GpuHealthMonitor()
    hs = mon.poll()
    assert [h.index for h in hs] == [0, 1] and all(h.available and h.healthy for h in hs) and hs[0].name == "NVIDIA B200"
    assert hs[1].throttle_reasons == ["hw_thermal_slowdown", "sw_power_cap"] and hs[0].memory_used_mb == 3072
    state["ecc"][(0, 1)] = 2                                        # uncorrected ECC on device 0
    state["events"].append(types.SimpleNamespace(device=("h", 1), eventData=79))      # Xid 79 on device 1
    assert mon.unhealthy_devices() == [0, 1]
    s = mon.summary()
    assert not s["all_healthy"] and s["devices"][1]["last_xid"] == 79 and s["throttled"] == ["hw_thermal_slowdown"]
    mon.close()
    assert mon.poll()[0].available is False


def test_monitor_without_nvml_reports_unavailable(monkeypatch):
    broken = types.ModuleType("pynvml")

    def boom():
        raise RuntimeError("no driver")

    broken.nvmlInit = boom
    monkeypatch.setitem(sys.modules, "pynvml", broken)
    mon = GH.GpuHealthMonitor()
    hs = mon.poll()
    assert len(hs) == 1 and not hs[0].available and hs[0].healthy and mon.summary()["all_healthy"]


def test_apply_health_masks_unhealthy_peers_only():
    """HybridEngine.apply_health: unhealthy device indices become bits of the channels' degraded-mode status word (own rank and
    out-of-range indices excluded); nothing happens without degraded_ok or without a peer heap."""
    from types import SimpleNamespace

    import torch

    from infomesh_b200.engine.hybrid import HybridEngine

    class Mon:
        def __init__(self, bad):
            self.bad = bad

        def unhealthy_devices(self):
            return self.bad

    def fake(world, rank, degraded_ok=True, heap=True):
        e = HybridEngine.__new__(HybridEngine)
        e.cfg = SimpleNamespace(degraded_ok=degraded_ok)
        e.ctx = SimpleNamespace(world=world, rank=rank)
        e.heap = object() if heap else None
        e.ch_dense = SimpleNamespace(status=torch.zeros(1, dtype=torch.int32))
        e.ch_bm25 = SimpleNamespace(status=torch.zeros(1, dtype=torch.int32))
        return e

    e = fake(8, 2)
    assert e.apply_health(Mon([5, 2, 11, 0])) == [5, 0]
    assert int(e.ch_dense.status.item()) == (1 << 5) | 1 and int(e.ch_bm25.status.item()) == (1 << 5) | 1
    assert e.apply_health(Mon([])) == [] and int(e.ch_dense.status.item()) == (1 << 5) | 1        # sticky
    assert fake(8, 2, degraded_ok=False).apply_health(Mon([5])) == []
    assert fake(8, 2, heap=False).apply_health(Mon([5])) == []
