"""Map the host -> device ceiling of this box: sustained pinned H2D rate per GPU when 1, 2, 4, ... GPUs copy at
the same time (one process per GPU, bound to the GPU's CPUs like the engine does, pinned buffer first-touched
by that process), plus what the system says about PCIe / NUMA placement.

    python tools/h2d_probe.py [--seconds 2.0] [--gb 4] [--out profiles/r02_h2d_matrix.json]

The end-to-end number of bench.py is ingest-bound (84 GB cross PCIe per step); this is its ceiling."""
import argparse
import glob
import json
import multiprocessing as mp
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(dev, nbytes, seconds, barrier, q, bind):
    import torch
    from solex_ser_recon_en_b200 import engine as E
    torch.cuda.set_device(dev)
    cpus = E._bind_to_gpu_numa_node(dev) if bind else None
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host.fill_(1)                                               # first touch on this process's CPUs
    d = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        d.copy_(host, non_blocking=True)
    st.synchronize()
    node = None
    try:                                                        # NUMA node that holds the pinned pages
        addr = host.data_ptr()
        for ln in open('/proc/self/numa_maps'):
            f = ln.split()
            if int(f[0], 16) <= addr < int(f[0], 16) + (1 << 40) and any(x.startswith('N') for x in f):
                a0 = int(f[0], 16)
                if a0 <= addr:
                    best = (a0, [x for x in f if x[0] == 'N' and '=' in x])
                    if node is None or best[0] > node[0]:
                        node = best
        node = node[1] if node else None
    except Exception:
        node = None
    barrier.wait()
    t0 = time.perf_counter()
    reps = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record()
        while time.perf_counter() - t0 < seconds:
            d.copy_(host, non_blocking=True)
            reps += 1
            if reps % 2 == 0:
                st.synchronize()
        e1.record()
    st.synchronize()
    q.put((dev, reps * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9, cpus, node))


def run_subset(devs, nbytes, seconds, bind):
    ctx = mp.get_context('spawn')
    barrier = ctx.Barrier(len(devs))
    q = ctx.Queue()
    ps = [ctx.Process(target=worker, args=(d, nbytes, seconds, barrier, q, bind)) for d in devs]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in devs]
    for p in ps:
        p.join()
    return sorted(res)


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout.strip()
    except Exception as e:
        return repr(e)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--seconds', type=float, default=2.0)
    ap.add_argument('--gb', type=float, default=4.0)
    ap.add_argument('--out', default='')
    ap.add_argument('--subsets', default='', help="e.g. '0,2;0,4' (default: 1, 2, 4, ... GPUs and a far pair)")
    a = ap.parse_args()
    import torch
    n = torch.cuda.device_count()
    nbytes = int(a.gb * (1 << 30))
    subsets = [[0]]
    if n >= 2:
        subsets += [[0, 1]]
    if n >= 3:
        subsets += [[0, 2]]
    if n >= 4:
        subsets += [[0, 1, 2, 3]]
    if n >= 8:
        subsets += [[0, 4], list(range(8))]
    if a.subsets:
        subsets = [[int(x) for x in grp.split(',')] for grp in a.subsets.split(';')]
    out = {'gpus_visible': n, 'host_cpus': os.cpu_count(), 'seconds_per_point': a.seconds, 'buffer_GB': a.gb, 'points': []}
    for devs in subsets:
        for bind in (True, False) if devs == list(range(n)) and n > 1 else (True,):
            r = run_subset(devs, nbytes, a.seconds, bind)
            out['points'].append({'gpus': devs, 'bound_to_gpu_cpus': bind,
                                  'per_gpu_GBps': {str(d): round(g, 2) for d, g, _, _ in r},
                                  'aggregate_GBps': round(sum(g for _, g, _, _ in r), 1),
                                  'cpus': {str(d): (None if c is None else '%d cpus: %d..%d' % (len(c), c[0], c[-1]))
                                           for d, _, c, _ in r},
                                  'pinned_pages_numa': {str(d): nd for d, _, _, nd in r}})
            print(json.dumps(out['points'][-1]), flush=True)
    out['nvidia_smi_topo'] = sh('nvidia-smi topo -m')
    out['numa_nodes'] = sh('ls -d /sys/devices/system/node/node* 2>/dev/null | wc -l')
    out['gpu_pci_numa'] = {os.path.basename(os.path.dirname(p)): open(p).read().strip()
                           for p in glob.glob('/sys/bus/pci/devices/*/numa_node')
                           if os.path.exists(os.path.join(os.path.dirname(p), 'vendor')) and
                           open(os.path.join(os.path.dirname(p), 'vendor')).read().strip() == '0x10de'}
    out['pcie_link'] = sh('nvidia-smi --query-gpu=index,pcie.link.gen.current,pcie.link.width.current,pci.bus_id --format=csv')
    out['lscpu'] = sh("lscpu | egrep 'Model name|Socket|NUMA|^CPU\\(s\\)'")
    text = json.dumps(out, indent=1)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        open(a.out, 'w').write(text + '\n')
    print(text)


if __name__ == '__main__':
    main()
