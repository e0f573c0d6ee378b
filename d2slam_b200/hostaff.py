"""Host-side placement helper for applications that drive the C ABI (bench.py, tools/): run the feeding threads and
allocate the pinned staging memory on the NUMA node the GPU hangs off, the equivalent of ``numactl --cpunodebind``.
Cross-socket pinned memory roughly halves the host->device bandwidth on a two-socket box."""
import os


def gpu_numa_cpus(device=0):
    """CPU ids of the NUMA node of CUDA device ``device`` (None when the topology cannot be read)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return cpus or None
    except Exception:
        return None


def bind_to_gpu_node(device=0):
    """Restrict this process (and the threads it creates from now on) to the GPU's NUMA node. Returns the cpu count
    now available, or None when nothing was changed."""
    cpus = gpu_numa_cpus(device)
    if not cpus:
        return None
    try:
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return len(allowed)
    except Exception:
        return None
