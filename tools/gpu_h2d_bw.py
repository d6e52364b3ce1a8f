"""Pinned / pageable host->device bandwidth of the box (context for the upload path)."""
import time, torch, subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current", "--format=csv"], capture_output=True, text=True).stdout)
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
p = torch.empty(n, dtype=torch.uint8).pin_memory()
q = torch.empty(n, dtype=torch.uint8)
for name, src in (("pinned", p), ("pageable", q)):
    for _ in range(2):
        torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(src, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"H2D {name}: {n / dt / 1e9:.1f} GB/s")
torch.cuda.synchronize(); t = time.perf_counter(); p.copy_(d, non_blocking=True); torch.cuda.synchronize(); print(f"D2H pinned: {n / (time.perf_counter() - t) / 1e9:.1f} GB/s")
t = time.perf_counter(); p.copy_(q); print(f"host pageable->pinned memcpy (1 thread, torch): {n / (time.perf_counter() - t) / 1e9:.1f} GB/s")
