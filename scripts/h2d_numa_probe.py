#!/usr/bin/env python
"""Host-to-device copy rate from pinned memory as a function of where the pinned pages live (NUMA node of the allocating
thread): is the 24 GB/s the streamed search sees (bench.py e2e.h2d_gbs_per_rank) a PCIe limit or a placement effect?"""
import os, sys, json, subprocess, time
import torch

def cpus_of_node(n):
    try:
        txt = open("/sys/devices/system/node/node%d/cpulist" % n).read().strip()
    except OSError:
        return None
    out = []
    for part in txt.split(","):
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out

def main():
    res = {"nodes": {}, "topo": None}
    try:
        res["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout[-1500:]
    except Exception as e:
        res["topo"] = repr(e)
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id,pcie.link.gen.current,pcie.link.width.current", "--format=csv,noheader", "-i", "0"],
                             capture_output=True, text=True, timeout=20).stdout.strip()
        res["gpu0"] = bus
        bid = bus.split(",")[0].strip().lower()
        bid = bid[4:] if len(bid) > 12 else bid
        res["gpu0_numa_node"] = open("/sys/bus/pci/devices/%s/numa_node" % bid).read().strip()
    except Exception as e:
        res["gpu0_numa_node"] = repr(e)
    torch.cuda.set_device(0)
    nbytes = 1 << 30
    dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    all_cpus = sorted(os.sched_getaffinity(0))
    res["affinity"] = len(all_cpus)
    node = 0
    while True:
        cpus = cpus_of_node(node)
        if cpus is None:
            break
        use = [c for c in cpus if c in all_cpus]
        if use:
            os.sched_setaffinity(0, use)
            src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            src.fill_(7)
            best = 0.0
            for rep in range(4):
                torch.cuda.synchronize(); t = time.perf_counter()
                dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
                best = max(best, nbytes / (time.perf_counter() - t) / 1e9)
            res["nodes"][str(node)] = {"cpus": len(use), "h2d_gbs": round(best, 1)}
            del src
        node += 1
    os.sched_setaffinity(0, all_cpus)
    print(json.dumps(res))

main()
