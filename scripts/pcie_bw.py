"""Development aid: pinned host<->device copy bandwidth of the box (uni- and bidirectional), the ceiling of the
e2e arm of bench.py (host arrays in, host arrays out every step)."""
import time
import torch

n = 1 << 30                                    # 1 GiB per buffer
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return reps * n / (time.time() - t0) / 1e9


run(True, True, 1)
print(f"H2D alone        {run(True, False):6.1f} GB/s")
print(f"D2H alone        {run(False, True):6.1f} GB/s")
print(f"both directions  {run(True, True):6.1f} GB/s each")
