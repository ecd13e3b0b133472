#!/usr/bin/env python3
"""Where the aggregate device-to-host wall of a multi-GPU box comes from (plain FASTQ into one
host tops out near 70-90 GB/s whatever the GPU count).  All visible GPUs copy 691 MB buffers into pinned host memory at
the same time; the pinned buffers are allocated (a) by the main thread, wherever the allocator puts them, (b) by one
thread per GPU that first binds itself to the CPUs local to its GPU (/sys/bus/pci/devices/<bdf>/local_cpulist), so that
first touch lands on the GPU's NUMA node.  Prints the NUMA layout, per-GPU alone and aggregate bandwidths as JSON."""
import glob
import json
import os
import threading
import time

import torch

n_gpu = torch.cuda.device_count()
nbytes = 691_000_000
out = {"n_gpus": n_gpu, "numa_nodes": len(glob.glob("/sys/devices/system/node/node[0-9]*")), "cpus": len(os.sched_getaffinity(0))}


def local_cpus(i):
    bdf = None
    if bdf is None:
        import pynvml
        pynvml.nvmlInit()
        bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(i)).busId
        bdf = bdf.decode() if isinstance(bdf, bytes) else bdf
    bdf = bdf.lower()
    if len(bdf.split(":")[0]) == 8:
        bdf = bdf[4:]
    try:
        txt = open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read().strip()
        node = open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip()
    except Exception as e:
        return None, str(e)
    cpus = set()
    for part in txt.split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    return cpus & os.sched_getaffinity(0), node


out["gpu_numa_node"] = [local_cpus(i)[1] for i in range(n_gpu)]
dev = [torch.empty(nbytes, dtype=torch.uint8, device="cuda:%d" % i) for i in range(n_gpu)]


def run(hosts, which, reps=8):
    streams = [torch.cuda.Stream(device=i) for i in range(n_gpu)]
    for i in which:
        with torch.cuda.stream(streams[i]):
            hosts[i].copy_(dev[i], non_blocking=True)
    for i in which:
        streams[i].synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in which:
            with torch.cuda.stream(streams[i]):
                hosts[i].copy_(dev[i], non_blocking=True)
    for i in which:
        streams[i].synchronize()
    return len(which) * reps * nbytes / (time.perf_counter() - t0) / 1e9


hosts_a = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(n_gpu)]
out["default_alloc"] = {"alone_GBps": [round(run(hosts_a, [i]), 1) for i in range(n_gpu)], "all_GBps": round(run(hosts_a, list(range(n_gpu))), 1)}
for k in (2, 4):
    if k < n_gpu:
        out["default_alloc"]["first_%d_GBps" % k] = round(run(hosts_a, list(range(k))), 1)
del hosts_a
hosts_b = [None] * n_gpu


def alloc(i):
    cpus, _ = local_cpus(i)
    if cpus:
        os.sched_setaffinity(0, cpus)
    hosts_b[i] = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    hosts_b[i].fill_(1)


th = [threading.Thread(target=alloc, args=(i,)) for i in range(n_gpu)]
for t in th:
    t.start()
for t in th:
    t.join()
out["numa_local_alloc"] = {"alone_GBps": [round(run(hosts_b, [i]), 1) for i in range(n_gpu)], "all_GBps": round(run(hosts_b, list(range(n_gpu))), 1)}
print(json.dumps(out, indent=1))
