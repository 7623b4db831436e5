"""Host-side probe: where does the GPU hang (NUMA node), where may this process run and allocate, and what pinned
H2D/D2H bandwidth does each memory node give.  Diagnostic for the end-to-end figure of bench.py."""
import ctypes, glob, os, subprocess, time
import torch

def sh(c):
    try:
        return subprocess.run(c, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:
        return repr(e)

print("affinity", sorted(os.sched_getaffinity(0)))
print(sh("lscpu | grep -i -E 'numa|model name|socket|^cpu\\(s\\)'"))
print("nodes", [os.path.basename(p) for p in glob.glob('/sys/devices/system/node/node*')])
print(sh("grep -E 'Mems_allowed_list|Cpus_allowed_list' /proc/self/status"))
bus = torch.cuda.get_device_properties(0).pci_bus_id if hasattr(torch.cuda.get_device_properties(0), 'pci_bus_id') else None
print(sh("nvidia-smi --query-gpu=pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv"))
for p in glob.glob('/sys/bus/pci/devices/*/numa_node'):
    dev = os.path.dirname(p)
    try:
        if open(dev + '/vendor').read().strip() == '0x10de' and open(dev + '/class').read().startswith('0x0302'):
            print(dev, 'numa_node', open(p).read().strip(), 'local_cpulist', open(dev + '/local_cpulist').read().strip())
    except OSError:
        pass
print(sh("nvidia-smi topo -m | head -20"))

libc = ctypes.CDLL(None, use_errno=True)
SYS_set_mempolicy = 238
def set_policy(node):
    if node is None:
        r = libc.syscall(SYS_set_mempolicy, 0, None, 0)
    else:
        mask = ctypes.c_ulong(1 << node)
        r = libc.syscall(SYS_set_mempolicy, 2, ctypes.byref(mask), 65)  # MPOL_BIND
    return r, ctypes.get_errno()

nbytes = 2 << 30
d = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
nodes = sorted(int(os.path.basename(p)[4:]) for p in glob.glob('/sys/devices/system/node/node*'))
for node in [None] + nodes:
    r = set_policy(node)
    if r[0] != 0:
        print('node', node, 'set_mempolicy failed', r); continue
    try:
        h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    except Exception as e:
        print('node', node, 'alloc failed', e); continue
    h.fill_(1)
    set_policy(None)
    for name, fn in (('h2d', lambda: d.copy_(h, non_blocking=True)), ('d2h', lambda: h.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(3): fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) / 3
        print('node', node, name, round(nbytes / dt / 1e9, 1), 'GB/s')
    # both directions at once
    d2 = torch.empty(nbytes // 2, dtype=torch.uint8, device='cuda'); h2 = torch.empty(nbytes // 2, dtype=torch.uint8).pin_memory()
    s2 = torch.cuda.Stream()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(3):
        d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 3
    print('node', node, 'h2d 2GiB + d2h 1GiB concurrently', round(dt * 1e3, 1), 'ms')
    del h, h2, d2
