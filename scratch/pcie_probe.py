"""H2D / D2H / bidirectional PCIe bandwidth with pinned memory (what bounds bench.py's e2e number)."""
import torch, time
dev = torch.device("cuda", 0)
for mb in (256, 2048):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(h2d, d2h, reps=5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()
        return reps * n / (time.perf_counter() - t0) / 1e9
    run(True, True, 1)
    print("piece %5d MB: H2D %.1f GB/s  D2H %.1f GB/s  both: %.1f GB/s each direction" % (mb, run(True, False), run(False, True), run(True, True)))
