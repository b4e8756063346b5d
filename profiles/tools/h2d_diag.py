import os, time, torch, subprocess
print(subprocess.run("nvidia-smi topo -m | head -8; lscpu | grep -i 'numa\|socket\|model name' ; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c", shell=True, capture_output=True, text=True).stdout)
print("affinity", sorted(os.sched_getaffinity(0)))
torch.cuda.init()
d = torch.empty(2_400_000, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
def bw(h, label):
    with torch.cuda.stream(s):
        for _ in range(5): d.copy_(h, non_blocking=True)
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(20):
            e0.record(s); d.copy_(h, non_blocking=True); e1.record(s); s.synchronize()
            ts.append(e0.elapsed_time(e1))
        t0 = time.perf_counter()
        for _ in range(50):
            d.copy_(h, non_blocking=True); s.synchronize()
        wall = (time.perf_counter() - t0) / 50
    ts.sort()
    print(f"{label}: device-timed median {ts[10]*1e3:.1f} us ({2.4/ts[10]:.1f} GB/s), min {ts[0]*1e3:.1f} us; wall per copy+sync {wall*1e6:.1f} us")
h = torch.empty(2_400_000, dtype=torch.uint8).pin_memory()
bw(h, "default affinity")
cpus = sorted(os.sched_getaffinity(0))
for part in (cpus[:len(cpus)//2], cpus[len(cpus)//2:]):
    os.sched_setaffinity(0, part)
    h2 = torch.empty(2_400_000, dtype=torch.uint8).pin_memory(); h2.fill_(1)
    bw(h2, f"affinity {part[0]}-{part[-1]}")
os.sched_setaffinity(0, cpus)
# small D2H
h3 = torch.empty(100_000, dtype=torch.uint8).pin_memory()
with torch.cuda.stream(s):
    t0 = time.perf_counter()
    for _ in range(50):
        h3.copy_(d[:100000], non_blocking=True); s.synchronize()
    print("D2H 100 KB copy+sync wall", (time.perf_counter() - t0) / 50 * 1e6, "us")
    t0 = time.perf_counter()
    for _ in range(200):
        s.synchronize()
    print("empty sync", (time.perf_counter() - t0) / 200 * 1e6, "us")
