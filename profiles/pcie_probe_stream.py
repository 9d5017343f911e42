# PCIe / host-memory ceiling for the access pattern of the host-buffer trace path: multi-GB pinned buffers streamed once, in 8 MB
# pieces on several streams, both directions at the same time (profiles/pcie_probe.py re-copies one 256 MiB buffer instead).
import torch, time
GB = 1 << 30
nbytes = 6 * GB
piece = 8 << 20
h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
d = [torch.empty(piece, dtype=torch.uint8, device='cuda') for _ in range(12)]
streams = [torch.cuda.Stream() for _ in range(6)]
def run(h2d, d2h):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(nbytes // piece):
        s = streams[i % 6]
        with torch.cuda.stream(s):
            if h2d: d[i % 6].copy_(h_in[i * piece:(i + 1) * piece], non_blocking=True)
            if d2h: h_out[i * piece:(i + 1) * piece].copy_(d[6 + i % 6], non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter() - t0
for _ in range(2):
    a, b, c = run(True, False), run(False, True), run(True, True)
    print("6 GB in 8 MB pieces on 6 streams: H2D %.1f GB/s, D2H %.1f GB/s, both at once %.1f GB/s each way (%.1f GB/s together)" % (nbytes / a / 1e9, nbytes / b / 1e9, nbytes / c / 1e9, 2 * nbytes / c / 1e9), flush=True)
