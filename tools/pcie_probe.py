"""PCIe copy characteristics on this box: one direction, both directions at once, chunked copies (CUDA events)."""
import torch
dev = torch.device("cuda", 0)
MB = 5.25
n = int(MB * 2**20 // 8)
h_in = torch.randn(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_a = torch.empty(n, dtype=torch.float64, device=dev)
d_b = torch.randn(n, dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

def h2d():
    d_a.copy_(h_in, non_blocking=True)
def d2h():
    h_out.copy_(d_b, non_blocking=True)
def both():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)
def chunked(k, fn_dir):
    c = n // k
    def f():
        for i in range(k):
            if fn_dir == "h2d":
                d_a[i * c:(i + 1) * c].copy_(h_in[i * c:(i + 1) * c], non_blocking=True)
            else:
                h_out[i * c:(i + 1) * c].copy_(d_b[i * c:(i + 1) * c], non_blocking=True)
    return f
t = timed(h2d); print(f"H2D {MB} MB: {t:.1f} us = {MB * 2**20 / t / 1e3:.1f} GB/s")
t = timed(d2h); print(f"D2H {MB} MB: {t:.1f} us = {MB * 2**20 / t / 1e3:.1f} GB/s")
t = timed(both); print(f"H2D + D2H concurrently: {t:.1f} us")
for k in (2, 4, 8, 16, 48):
    print(f"H2D in {k} chunks: {timed(chunked(k, 'h2d')):.1f} us   D2H in {k} chunks: {timed(chunked(k, 'd2h')):.1f} us")
# graph of chunked copies
for k in (16, 48):
    g = torch.cuda.CUDAGraph()
    f = chunked(k, "h2d")
    with torch.cuda.graph(g):
        f()
    print(f"H2D in {k} chunks, graph: {timed(g.replay):.1f} us")
# zero-copy: a kernel reading pinned host memory / writing pinned host memory (torch elementwise copy on a mapped tensor is not available; skip)
