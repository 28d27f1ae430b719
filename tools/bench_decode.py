#!/usr/bin/env python3
"""Times the block decoders (device-resident) at the BASELINE sizes: encode a synthetic image, decode it K times
with CUDA events, print one JSON line per codec with the algorithmic bandwidth (blocks read + pixels written)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import image_compression_b200 as icb  # noqa: E402


def main():
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    for name, codec, fmt, nc_in, n in (("dxt1", 0, icb.RGB, 3, 8192), ("dxt5", 1, icb.RGBA, 4, 8192), ("etc1", 2, icb.RGB, 3, 4096)):
        src = torch.empty(n * n * nc_in, dtype=torch.uint8, device="cuda")
        icb.fill_synthetic(src, 1 if nc_in == 3 else 2)
        blocks = [icb.encode_device(codec, fmt, src, n, n) for _ in range(2)]
        nc_out = 4 if codec == 1 else 3
        outs = [torch.empty(n * n * nc_out, dtype=torch.uint8, device="cuda") for _ in range(2)]
        for i in range(5):
            icb.decode_device(codec, blocks[i % 2], n, n, out=outs[i % 2])
        torch.cuda.synchronize()
        steps = 30
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            icb.decode_device(codec, blocks[i % 2], n, n, out=outs[i % 2])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        nbytes = blocks[0].numel() + outs[0].numel()
        print(json.dumps({"workload": name + "_decode_%dx%d" % (n, n), "ms": ms, "mpix_s": n * n / ms / 1e3,
                          "algorithmic_bytes": nbytes, "gb_s": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak}))


if __name__ == "__main__":
    main()
