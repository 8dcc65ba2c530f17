"""Bind a rank to the host NUMA node of its GPU.

One process per GPU copies its own shard of waveforms host->device every step (dist.py); on a two-socket B200 box the
eight ranks otherwise allocate their pinned buffers wherever the launcher happened to run and half of the copies
cross the inter-socket link.  ``bind_to_gpu_node`` restricts the calling process to the CPUs of the NUMA node the GPU's
PCIe root hangs off (sysfs), so that first-touch places the pinned staging buffers allocated AFTERWARDS on that node.
Everything here is best effort: any failure leaves the process as it was and is reported in the returned dict.
"""
import os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(device_index):
    """NUMA node of CUDA device `device_index` (None if sysfs does not say)."""
    import torch
    props = torch.cuda.get_device_properties(device_index)
    bdf = None
    if hasattr(props, "pci_bus_id") and hasattr(props, "pci_device_id"):
        domain = getattr(props, "pci_domain_id", 0)
        bdf = "%04x:%02x:%02x.0" % (domain, props.pci_bus_id, props.pci_device_id)
    if bdf is None:
        return None
    path = f"/sys/bus/pci/devices/{bdf}/numa_node"
    try:
        node = int(open(path).read().strip())
    except (OSError, ValueError):
        return None
    return node if node >= 0 else None


def bind_to_gpu_node(device_index):
    """-> {"node": n or None, "cpus": count, "bound": bool, "why": str}"""
    info = {"node": None, "cpus": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None,
            "bound": False, "why": ""}
    try:
        node = gpu_numa_node(device_index)
        if node is None:
            info["why"] = "no NUMA node in sysfs for this GPU"
            return info
        info["node"] = node
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if not target:
            info["why"] = "the node's CPUs are outside this process's affinity mask"
            return info
        if target != allowed:
            os.sched_setaffinity(0, target)
        info["cpus"] = len(target)
        info["bound"] = True
    except Exception as exc:   # best effort
        info["why"] = f"{type(exc).__name__}: {exc}"
    return info
