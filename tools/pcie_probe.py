import torch, time
n = 507_000_000
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, chunks=8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c = n // chunks
    for k in range(chunks):
        if h2d:
            with torch.cuda.stream(s1): d_in[k*c:(k+1)*c].copy_(h_in[k*c:(k+1)*c], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out[k*c:(k+1)*c].copy_(d_out[k*c:(k+1)*c], non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
for _ in range(2): run(True, True)
print("h2d only ms", min(run(True, False) for _ in range(3)))
print("d2h only ms", min(run(False, True) for _ in range(3)))
print("both ms", min(run(True, True) for _ in range(3)))
