"""PCIe probe: H2D / D2H / bidirectional bandwidth from pinned memory, and the NUMA picture."""
import os, subprocess, time
import torch
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
print("cpus allowed:", len(os.sched_getaffinity(0)), "nproc", os.cpu_count())
try:
    print(subprocess.run(["numactl", "-H"], capture_output=True, text=True).stdout[:800])
except Exception as e:
    print("numactl:", e)
dev = torch.device("cuda", 0)
N = 4 << 30
d1 = torch.empty(N, dtype=torch.uint8, device=dev)
d2 = torch.empty(N, dtype=torch.uint8, device=dev)
h1 = torch.empty(N, dtype=torch.uint8).pin_memory()
h2 = torch.empty(N, dtype=torch.uint8).pin_memory()
h1.fill_(1); h2.fill_(2)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return N / dt / 1e9
run(True, True, 1)
print("H2D only   GB/s", round(run(True, False), 1))
print("D2H only   GB/s", round(run(False, True), 1))
print("both, each GB/s", round(run(True, True), 1))
# many 33.5 MB copies instead of one 4 GB copy
chunk = 256 * 256 * 128 * 4
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s2):
    for o in range(0, N - chunk, chunk):
        h2[o:o + chunk].copy_(d2[o:o + chunk], non_blocking=True)
torch.cuda.synchronize()
print("D2H in 33.5 MB pieces GB/s", round((N - chunk) / (time.perf_counter() - t0) / 1e9, 1))
