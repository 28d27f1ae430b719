"""ETC1 exhaustive-search time on different image content (the no-clamp shortcut is chosen per warp, so speed depends on the
data; output is bit-exact either way).  Run on the GPU box: python tools/bench_etc1_content.py"""
import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import imagegen
import image_compression_b200 as icb
n = 4096
for kind in ("random", "smooth_noise", "gradient", "dark", "checker"):
    img = imagegen.make(kind, 1024, 1024, 3, seed=3)
    img = np.tile(img, (4, 4, 1))
    d = torch.from_numpy(np.ascontiguousarray(img).ravel()).cuda()
    out = torch.empty(n * n // 2, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        icb.encode_device(2, icb.RGB, d, n, n, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        icb.encode_device(2, icb.RGB, d, n, n, out=out)
    e1.record(); torch.cuda.synchronize()
    print("etc1 4096^2 %-14s %.1f us" % (kind, e0.elapsed_time(e1) * 100))
