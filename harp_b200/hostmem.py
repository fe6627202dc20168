"""Host-side placement for the host->device ingest path (`HarpFitter.step(batch=...)` / `batch_u8=...`).

The per-step image upload runs at PCIe line rate only when the pinned staging pages live on the NUMA node the GPU hangs off;
from the far socket the copies cross the inter-socket link.  `gpu_numa_affinity(device)` restricts the calling process to the
CPUs of that node for the duration of a `with` block, so that pinned allocations made inside it (first touch) land there.
The reference has no equivalent: its DataLoader tensors are pageable and copied with `.to(device)` (optimize_sequence.py:448)."""
import contextlib
import os

import torch


def _pci_bus_id(device):
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    p = torch.cuda.get_device_properties(idx)
    if all(hasattr(p, k) for k in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
        return f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    import pynvml  # nvidia-ml-py
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(p.uuid)).encode()) if hasattr(p, "uuid") else pynvml.nvmlDeviceGetHandleByIndex(idx)
    bus = pynvml.nvmlDeviceGetPciInfo(h).busId
    bus = bus.decode() if isinstance(bus, bytes) else bus
    return bus[-12:].lower()  # nvml pads the domain to 8 hex digits


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def gpu_numa_node(device):
    """-> (node, cpus) of the GPU's PCIe root, or (None, None) when the platform does not say (single node, VM)."""
    try:
        bus = _pci_bus_id(device)
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None, None
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        return node, cpus
    except Exception:
        return None, None


@contextlib.contextmanager
def gpu_numa_affinity(device):
    """Run the block on the CPUs of the GPU's NUMA node (no-op when unknown); yields a dict describing what was done."""
    info = {"numa_node": None, "cpus": None}
    old = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    node, cpus = gpu_numa_node(device)
    if old is not None and node is not None and cpus and (cpus & old):
        os.sched_setaffinity(0, cpus & old)
        info = {"numa_node": node, "cpus": len(cpus & old)}
    try:
        yield info
    finally:
        if old is not None:
            os.sched_setaffinity(0, old)
