"""Host placement: run the rank on the CPU cores of the NUMA node its GPU hangs off, BEFORE any pinned memory is
allocated, so that the D2H drain (2 B per base) does not cross the inter-socket link.  Best effort: silently does
nothing when sysfs / nvidia-smi do not answer."""
import os
import subprocess


def gpu_numa_node(device_index):
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(device_index)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bus:
            return None
        if len(bus.split(":")[0]) == 8:  # nvidia-smi prints an 8-digit PCI domain, sysfs uses 4
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def node_cpus(node):
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        return cpus
    except Exception:
        return set()


def bind_to_gpu_numa(device_index):
    """Returns a short description of what was done (for logs / bench config)."""
    node = gpu_numa_node(device_index)
    if node is None:
        return "numa: unknown"
    cpus = node_cpus(node) & os.sched_getaffinity(0)
    if not cpus:
        return "numa: node %d has no usable cpus" % node
    try:
        os.sched_setaffinity(0, cpus)
        return "numa: bound to node %d (%d cpus)" % (node, len(cpus))
    except Exception as e:
        return "numa: bind failed (%s)" % e
