"""GPU health polling (SURVEY §5.3 "ECC / Xid polling"): the signal that decides whether a shard may keep serving.

The reference's failure detection is CPU-side only (peer time-outs, ``ResourceGovernor`` degrade levels,
infomesh/resources/governor.py:206-266).  The device plane adds what a GPU node can actually observe through NVML:

* uncorrected ECC errors (volatile double-bit count) and pending page retirements / row-remap failures,
* the last Xid-class critical event seen on the device (NVML event set: XidCriticalError, DoubleBitEccError),
* thermal / power throttle reasons and whether the device "fell off the bus" (any NVML call failing with GPU_IS_LOST).

``GpuHealthMonitor.poll()`` returns one :class:`GpuHealth` per device; ``healthy`` going False is what the serving engine
feeds into ``HybridConfig.degraded_ok`` (answer from the remaining shards with a ``degraded`` flag instead of trapping)
and what ``infomesh doctor`` / ``/gpu/stats`` report.  Without NVML (no driver, CPU box) every device reports
``available=False`` and nothing else changes."""
from __future__ import annotations

import threading
import time
from dataclasses import asdict, dataclass, field

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

# clocks-event (throttle) reason bits worth surfacing (nvmlClocksEventReason*)
_REASONS = {0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake", 0x4: "sw_power_cap"}
_FATAL_REASONS = {"hw_slowdown", "hw_thermal_slowdown", "hw_power_brake"}


@dataclass
class GpuHealth:
    index: int
    available: bool = False
    healthy: bool = True
    name: str = ""
    ecc_uncorrected: int = 0
    ecc_corrected: int = 0
    retired_pages_pending: bool = False
    row_remap_failure: bool = False
    last_xid: int = 0
    xid_events: int = 0
    throttle_reasons: list[str] = field(default_factory=list)
    temperature_c: int = 0
    memory_used_mb: int = 0
    error: str = ""
    checked_at: float = 0.0

    def to_dict(self) -> dict:
        return asdict(self)


def assess(h: GpuHealth) -> bool:
    """Policy: a device is unhealthy when it is lost, reports uncorrected ECC errors, has a row-remap failure, or saw a
    critical Xid since the monitor started.  Throttling alone is reported but does not take a shard out of service."""
    if not h.available:
        return not h.error.startswith("lost")
    return not (h.ecc_uncorrected > 0 or h.row_remap_failure or h.xid_events > 0)


class GpuHealthMonitor:
    def __init__(self, device_indices: list[int] | None = None):
        self._nvml = None
        self._handles: dict[int, object] = {}
        self._events = None
        self._xid: dict[int, tuple[int, int]] = {}        # index -> (last xid, count)
        self._lock = threading.Lock()
        self._last: dict[int, GpuHealth] = {}
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            n = pynvml.nvmlDeviceGetCount()
            for i in (device_indices if device_indices is not None else range(n)):
                if 0 <= i < n:
                    self._handles[i] = pynvml.nvmlDeviceGetHandleByIndex(i)
            self._arm_events()
        except Exception as exc:  # noqa: BLE001 -- no driver / no NVML: the monitor degrades to "unavailable"
            logger.debug("nvml_unavailable", error=str(exc))
            self._nvml = None

    # ------------------------------------------------------------------ Xid events
    def _arm_events(self) -> None:
        nv = self._nvml
        try:
            self._events = nv.nvmlEventSetCreate()
            mask = nv.nvmlEventTypeXidCriticalError | nv.nvmlEventTypeDoubleBitEccError
            for h in self._handles.values():
                try:
                    nv.nvmlDeviceRegisterEvents(h, mask, self._events)
                except Exception:  # noqa: BLE001 -- not supported on this device / in this container
                    pass
        except Exception:  # noqa: BLE001
            self._events = None

    def _drain_events(self) -> None:
        nv = self._nvml
        if self._events is None:
            return
        for _ in range(64):
            try:
                ev = nv.nvmlEventSetWait_v2(self._events, 0)
            except Exception:  # noqa: BLE001 -- timeout = nothing pending
                return
            for i, h in self._handles.items():
                try:
                    same = nv.nvmlDeviceGetIndex(ev.device) == i
                except Exception:  # noqa: BLE001
                    same = False
                if same:
                    last, cnt = self._xid.get(i, (0, 0))
                    self._xid[i] = (int(getattr(ev, "eventData", 0)) or last, cnt + 1)

    # ------------------------------------------------------------------ polling
    def _read(self, i: int, handle) -> GpuHealth:
        nv = self._nvml
        h = GpuHealth(index=i, available=True, checked_at=time.time())

        def q(fn, *a, default=None):
            try:
                return fn(handle, *a)
            except Exception as exc:  # noqa: BLE001
                if "LOST" in str(exc).upper() or "fallen off" in str(exc).lower():
                    h.available, h.error = False, f"lost: {exc}"
                return default

        name = q(nv.nvmlDeviceGetName, default="")
        h.name = name.decode() if isinstance(name, bytes) else str(name or "")
        h.ecc_uncorrected = int(q(nv.nvmlDeviceGetTotalEccErrors, nv.NVML_MEMORY_ERROR_TYPE_UNCORRECTED, nv.NVML_VOLATILE_ECC, default=0) or 0)
        h.ecc_corrected = int(q(nv.nvmlDeviceGetTotalEccErrors, nv.NVML_MEMORY_ERROR_TYPE_CORRECTED, nv.NVML_VOLATILE_ECC, default=0) or 0)
        pend = q(nv.nvmlDeviceGetRetiredPagesPendingStatus, default=0)
        h.retired_pages_pending = bool(pend)
        rows = q(nv.nvmlDeviceGetRemappedRows, default=None)
        if rows is not None and len(rows) >= 4:
            h.row_remap_failure = bool(rows[3])
        reasons = int(q(nv.nvmlDeviceGetCurrentClocksThrottleReasons, default=0) or 0)
        h.throttle_reasons = sorted(name for bit, name in _REASONS.items() if reasons & bit)
        h.temperature_c = int(q(nv.nvmlDeviceGetTemperature, nv.NVML_TEMPERATURE_GPU, default=0) or 0)
        mem = q(nv.nvmlDeviceGetMemoryInfo, default=None)
        h.memory_used_mb = int(mem.used // 2 ** 20) if mem is not None else 0
        h.last_xid, h.xid_events = self._xid.get(i, (0, 0))
        h.healthy = assess(h)
        return h

    def poll(self) -> list[GpuHealth]:
        with self._lock:
            if self._nvml is None:
                return [GpuHealth(index=i, available=False, error="nvml unavailable", checked_at=time.time()) for i in sorted(self._handles)] or \
                       [GpuHealth(index=0, available=False, error="nvml unavailable", checked_at=time.time())]
            self._drain_events()
            out = [self._read(i, h) for i, h in sorted(self._handles.items())]
            for h in out:
                prev = self._last.get(h.index)
                if prev is not None and prev.healthy and not h.healthy:
                    logger.warning("gpu_unhealthy", index=h.index, ecc_uncorrected=h.ecc_uncorrected, xid=h.last_xid, error=h.error)
                self._last[h.index] = h
            return out

    def unhealthy_devices(self) -> list[int]:
        return [h.index for h in self.poll() if not h.healthy]

    def summary(self) -> dict:
        hs = self.poll()
        return {"devices": [h.to_dict() for h in hs], "all_healthy": all(h.healthy for h in hs),
                "throttled": sorted({r for h in hs for r in h.throttle_reasons if r in _FATAL_REASONS})}

    def close(self) -> None:
        with self._lock:
            if self._nvml is not None:
                try:
                    if self._events is not None:
                        self._nvml.nvmlEventSetFree(self._events)
                    self._nvml.nvmlShutdown()
                except Exception:  # noqa: BLE001
                    pass
                self._nvml = None
