"""Host->device bandwidth from pinned memory allocated on each NUMA node (run under gpurun): is the 35 GB/s of the e2e path
a PCIe limit or a cross-socket one?"""
import glob
import json
import os
import time

import torch

dev = 0
props = torch.cuda.get_device_properties(dev)
bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
node_of_gpu = None
try:
    node_of_gpu = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
except Exception as e:  # noqa: BLE001
    node_of_gpu = f"? ({e})"
nodes = sorted(int(p.split("node")[-1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))


def cpus(node):
    out = []
    for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


allowed = sorted(os.sched_getaffinity(0))
res = {"gpu_bus": bus, "gpu_numa_node": node_of_gpu, "nodes": nodes, "allowed_cpus": len(allowed), "cpu_count": os.cpu_count(), "bw": {}}
d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for node in nodes:
    c = [x for x in cpus(node) if x in allowed]
    if not c:
        res["bw"][node] = "no allowed cpu on this node"
        continue
    os.sched_setaffinity(0, c)
    h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()      # first touch on this node
    h.fill_(1)
    torch.cuda.synchronize()
    best = 0
    for _ in range(5):
        t0 = time.perf_counter()
        d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        best = max(best, (1 << 30) / (time.perf_counter() - t0) / 1e9)
    res["bw"][node] = round(best, 1)
    del h
os.sched_setaffinity(0, allowed)
print(json.dumps(res))
