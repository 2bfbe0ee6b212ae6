"""Where is the GPU attached, and does the NUMA node of a pinned buffer matter for PCIe copies?  Prints the GPU's
numa_node, the nodes' CPU lists, and H2D / D2H bandwidth of a 4 GB pinned buffer allocated with the calling thread
bound to each node in turn."""
import ctypes
import glob
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402


def cpulist(s):
    out = []
    for part in s.strip().split(','):
        if not part:
            continue
        a, _, b = part.partition('-')
        out.extend(range(int(a), int(b or a) + 1))
    return out


res = {}
bus = torch.cuda.get_device_properties(0).pci_bus_id if hasattr(torch.cuda.get_device_properties(0), 'pci_bus_id') else None
cudart = torch.cuda.cudart()
buf = ctypes.create_string_buffer(64)
try:
    libcudart = ctypes.CDLL('libcudart.so.12')
    libcudart.cudaDeviceGetPCIBusId(buf, 64, 0)
    busid = buf.value.decode().lower()
except Exception as e:  # noqa: BLE001
    busid = None
    res['busid_error'] = str(e)
res['busid'] = busid
if busid:
    for cand in (busid, busid[4:] if busid.startswith('0000') else busid):
        p = '/sys/bus/pci/devices/%s/numa_node' % cand
        if os.path.exists(p):
            res['gpu_numa_node'] = int(open(p).read())
            break
nodes = {}
for d in sorted(glob.glob('/sys/devices/system/node/node[0-9]*')):
    nodes[int(d.rsplit('node', 1)[1])] = cpulist(open(d + '/cpulist').read())
res['nodes'] = {k: '%d cpus (%s..%s)' % (len(v), v[0] if v else None, v[-1] if v else None) for k, v in nodes.items()}
res['affinity_at_start'] = len(os.sched_getaffinity(0))
allowed = os.sched_getaffinity(0)
nbytes = 4 << 30
dev = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
bw = {}
for node, cpus in nodes.items():
    use = set(cpus) & allowed
    if not use:
        bw[node] = 'no allowed cpu'
        continue
    os.sched_setaffinity(0, use)
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize()
    out = {}
    for name, (dst, src) in dict(h2d=(dev, host), d2h=(host, dev)).items():
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        out[name] = 3 * nbytes / (time.perf_counter() - t) / 1e9
    bw[node] = out
    del host
    os.sched_setaffinity(0, allowed)
res['pinned_copy_GBps_by_node_of_allocation'] = bw
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(REPO, 'gpurun_out'), exist_ok=True)
json.dump(res, open(os.path.join(REPO, 'gpurun_out', 'numa_probe.json'), 'w'), indent=1)
