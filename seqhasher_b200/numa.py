"""Put a feeder process next to its GPU: pinned staging that lives on the GPU's own NUMA node is read by the DMA engine
without crossing the inter-socket link (each GPU has its own pinned staging and feeder thread; the
end-to-end limiter of an 8-GPU box is the host side of PCIe).  Linux only, best effort, no dependency beyond /sys."""
from __future__ import annotations

import os


def _read(path: str):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_cpulist(s: str) -> set[int]:
    out: set[int] = set()
    for part in s.split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            out.update(range(int(a), int(b) + 1))
        else:
            out.add(int(part))
    return out


def gpu_numa_node(pci_bus_id: str) -> int | None:
    """NUMA node of the PCI device `0000:1b:00.0` (as torch / nvidia-smi report it), or None when the kernel does not say."""
    bus = pci_bus_id.lower()
    if len(bus.split(":")[0]) == 8:          # nvidia-smi prints an 8-digit domain, sysfs uses 4
        bus = bus[4:]
    v = _read(f"/sys/bus/pci/devices/{bus}/numa_node")
    try:
        n = int(v) if v is not None else -1
    except ValueError:
        n = -1
    return n if n >= 0 else None


def bind_to_gpu_node(pci_bus_id: str) -> dict:
    """Restrict this process to the CPUs of the GPU's NUMA node (memory then follows first touch, and cudaHostAlloc's pages
    with it).  Returns what was done, for the benchmark record."""
    node = gpu_numa_node(pci_bus_id)
    info = {"pci": pci_bus_id, "numa_node": node, "bound": False}
    if node is None:
        return info
    cpus = _read(f"/sys/devices/system/node/node{node}/cpulist")
    if not cpus:
        return info
    want = _parse_cpulist(cpus)
    try:
        allowed = os.sched_getaffinity(0)
        use = want & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            info["bound"] = True
        info["cpus"] = len(use or allowed)
    except (AttributeError, OSError):
        pass
    return info


def n_numa_nodes() -> int:
    try:
        return len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
    except OSError:
        return 1
