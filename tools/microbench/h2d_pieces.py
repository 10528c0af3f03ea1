"""H2D bandwidth out of pinned host memory vs piece size and vs who wrote the bytes last (cold in DRAM / just written by many host threads):
explains the upload rate of the Parquet scan (tools/bench_shapes.py M6)."""
import threading, time
import numpy as np, torch

N = 256 << 20
h = torch.empty(N, dtype=torch.uint8).pin_memory(); d = torch.empty(N, dtype=torch.uint8, device="cuda")
h.fill_(1)
def h2d(piece):
    torch.cuda.synchronize(); t = time.perf_counter()
    for o in range(0, N, piece): d[o:o + piece].copy_(h[o:o + piece], non_blocking=True)
    torch.cuda.synchronize(); return N / (time.perf_counter() - t) / 1e9
h2d(N)
for piece in (256 << 20, 8 << 20, 2 << 20, 1 << 20, 256 << 10):
    print(f"pieces of {piece >> 10:7d} KiB, bytes cold in DRAM: {max(h2d(piece) for _ in range(3)):6.1f} GB/s", flush=True)
src = np.random.randint(0, 255, N, dtype=np.uint8); hn = h.numpy()
def rewrite(nth):
    def w(i): s = i * (N // nth); hn[s:s + N // nth] = src[s:s + N // nth]
    ths = [threading.Thread(target=w, args=(i,)) for i in range(nth)]; [t.start() for t in ths]; [t.join() for t in ths]
for nth in (1, 8, 32):
    r = []
    for _ in range(3): rewrite(nth); r.append(h2d(8 << 20))
    print(f"pieces of    8192 KiB, just rewritten by {nth:2d} host threads: {max(r):6.1f} GB/s (min {min(r):.1f})", flush=True)
# copies racing with host threads that keep writing OTHER pinned memory (the scan's workers run ahead of the uploads)
h2 = torch.empty(N, dtype=torch.uint8).pin_memory(); hn2 = h2.numpy(); stop = False
def churn(i):
    s = i * (N // 32)
    while not stop: hn2[s:s + N // 32] = src[s:s + N // 32]
ths = [threading.Thread(target=churn, args=(i,)) for i in range(32)]; [t.start() for t in ths]
time.sleep(0.2); r = [h2d(8 << 20) for _ in range(5)]; stop = True; [t.join() for t in ths]
print(f"pieces of    8192 KiB, while 32 host threads write other pinned memory: {max(r):6.1f} GB/s (min {min(r):.1f})", flush=True)
