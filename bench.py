#!/usr/bin/env python3
"""bench.py -- Mpixels/s of the B200 texture-block encoder on BASELINE.json's headline configuration.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload dxt1_rgba8] [--impl reference]

Default workload (config.workload): DXT1 encode of ONE 8192x8192 synthetic RGBA8 image (BASELINE.json configs[1]),
input resident in HBM.  One "step" = one pass of the encoder over the image.

N = 1   one kernel launch per step; `value` = 8192^2 / step time.
N > 1   (torchrun, one rank per GPU) the SAME single image, sharded by block-row stripes: rank r holds and encodes the
        block rows icb_stripe_partition gives it and its encode kernel stores the blocks straight into rank 0's buffer
        over NVLink (peer-mapped output, sharding.PeerStream) -- the north star's "gather of the packed block stream"
        fused into the kernels.  The delivery is INSIDE the timed region: `value` = 8192^2 / (time until the whole
        stream sits on rank 0), "scaling": "strong".  The job is bounded by rank 0's NVLink ingress from N = 4 up
        ("delivery" key: bytes into the root / time against 900 GB/s); kernel-only, NCCL-gather and weak-scaling
        figures are reported beside it under their own keys, not in `value`.
        PVRTC does not shard through its public shape rules (square power-of-two images): N replicas, "weak".

Keys beyond the base contract: "roofline" (dominant kernel vs MEASURED_PEAKS.json HBM copy bandwidth), "parity" (the
run's own device output byte-compared with the reference CPU encoder's output of the same input), "cpu_baseline"
(N = 1: the unmodified reference, or the oracle port, timed on this box's host cores on a bounded sample), "e2e" (ONE
image through the host-buffer API with pinned buffers: H2D + kernels + D2H inside the timed region; at N > 1 rank 0
makes one icb_ctx_compress_host call that spreads the image over all N GPUs), "other_workloads" (N = 1: compact records
for the other BASELINE configurations and for structured image content), "clocks", "gpu_launches".

--impl reference times the reference's own CPU implementation (oracle/_ref when it was built, else the oracle
port) with all host threads on the same metric; rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: codec, format, components, image side, bytes out per pixel, synthetic seed, description
    "dxt1_rgba8": dict(codec=0, fmt=2, nc=4, n=8192, out_bpp=0.5, seed=2, desc="DXT1 8192x8192 RGBA8 (alpha ignored)"),
    "dxt1_rgb8": dict(codec=0, fmt=0, nc=3, n=8192, out_bpp=0.5, seed=1, desc="DXT1 8192x8192 RGB888"),
    "dxt5_rgba8": dict(codec=1, fmt=2, nc=4, n=8192, out_bpp=1.0, seed=2, desc="DXT5 8192x8192 RGBA8"),
    "etc1_rgb8": dict(codec=2, fmt=0, nc=3, n=4096, out_bpp=0.5, seed=1, desc="ETC1 4096x4096 RGB888, kSmallerError"),
    "pvrtc2_rgba8": dict(codec=3, fmt=2, nc=4, n=4096, out_bpp=0.25, seed=2, desc="PVRTC1 2bpp 4096x4096 RGBA8"),
}
L2_BYTES = 126 << 20
NVLINK_INGRESS_GBS = 900.0  # NVLink 5, per direction per GPU (B200_PROFILING.md)

_RESULT_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries underneath (NCCL prints its version banner to fd 1 at some
    debug levels) do not know that: point fd 1 at stderr for the duration of the run and keep the real stdout for the
    result line."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_RESULT_FD, data)


def metric_name(workload):
    return "Mpixels/sec DXT1 encode, 8K x 8K RGBA8" if workload == "dxt1_rgba8" else "Mpixels/sec " + WORKLOADS[workload]["desc"]


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def profile_constant(filename, workload):
    """Per-launch figures of the dominant kernel taken from committed ncu captures (profiles/<filename>): DRAM traffic
    (traffic.json) and integer-pipe instruction counts (pipe_counts.json)."""
    path = os.path.join(ROOT, "profiles", filename)
    if os.path.exists(path):
        try:
            return json.load(open(path)).get(workload)
        except Exception:
            return None
    return None


class ClockSampler:
    """Samples nvidia-smi while the timed region runs (B200_PROFILING.md clocks line)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], None, set(), []
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None,
                "sampled_over": "timed region + 0.5 s of the same launches back to back (100 ms nvidia-smi period)"}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation on the host cores
# ---------------------------------------------------------------------------------------------------------------

def cpu_reference_throughput(workload, steps, warmup):
    """Times the reference CPU encoder (oracle/_ref, unmodified, -O2) or, when that was not built, the oracle port,
    on T row stripes from T threads (zero-copy external-storage outputs).  The sample is a
    WIDTH x (256*T) stripe stack of the workload's synthetic image; Mpix/s does not depend on the sample height."""
    import numpy as np

    import checkers as ck
    wl = WORKLOADS[workload]
    n = wl["n"]
    threads = os.cpu_count() or 1
    use_ref = ck.have_ref()
    if workload == "pvrtc2_rgba8":
        threads = 1  # wrap-around + Z-order: cannot be stripe-split through the public API
        side = 1024
    rows_per_thread = 1024
    # The reference's DXT1 takes RGB888: feed it the alpha-stripped image (what a caller must do today).
    ref_nc = 3 if workload in ("dxt1_rgba8", "dxt1_rgb8", "etc1_rgb8") else 4
    if workload == "etc1_rgb8":
        rows_per_thread = 128  # exhaustive search: ~2 Mpix/s per core
    if workload == "pvrtc2_rgba8":
        src = ck.synthetic(side * side * 4, wl["seed"])
        sample_px = side * side
    else:
        # one stripe of input shared by all threads (each writes its own output slice): same work per thread as
        # a T-stripe image without generating T stripes of input
        raw = ck.synthetic(n * rows_per_thread * wl["nc"], wl["seed"])
        if wl["nc"] == 4 and ref_nc == 3:
            raw = np.ascontiguousarray(raw.reshape(-1, 4)[:, :3]).reshape(-1)
        src = raw
        sample_px = n * rows_per_thread * threads
    block_bytes = 16 if workload == "dxt5_rgba8" else 8
    out = np.zeros(sample_px // 16 * block_bytes if workload != "pvrtc2_rgba8" else sample_px // 4, np.uint8)

    def stripe_job(t):
        h = rows_per_thread
        s = src
        o = out[t * (h // 4) * (n // 4) * block_bytes:(t + 1) * (h // 4) * (n // 4) * block_bytes]
        if use_ref:
            L = ck.ref()
            if workload == "etc1_rgb8":
                ok = L.icref_etc_external(2, h, n, 0, ck._ptr(s), ck._ptr(o), o.size)
            else:
                fmt = ck.RGB if ref_nc == 3 else ck.RGBA
                ok = L.icref_dxt_external(fmt, h, n, 0, ck._ptr(s), ck._ptr(o), o.size)
            assert ok == 1
        else:
            L = ck.oracle()
            if workload == "etc1_rgb8":
                L.orc_etc1_compress(2, h, n, h, n, 0, ck._ptr(s), ck._ptr(o))
            else:
                L.orc_dxt_compress(ck.RGB if ref_nc == 3 else ck.RGBA, h, n, h, n, 0, ck._ptr(s), ck._ptr(o))

    def one_step():
        if workload == "pvrtc2_rgba8":
            if use_ref:
                res = ck.ref_pvrtc(src, side, side)
                assert res is not None
            else:
                ck.oracle_pvrtc(src, side, side)
            return
        ts = [threading.Thread(target=stripe_job, args=(t,)) for t in range(threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    for _ in range(max(1, warmup)):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / steps
    if workload == "pvrtc2_rgba8":
        sample = "%dx%d image, 1 thread (PVRTC cannot be stripe-split)" % (side, side)
    else:
        sample = "%d x %d px per step: %d threads x one %d-row stripe each%s" % (
            n, rows_per_thread * threads, threads, rows_per_thread, ", alpha stripped to RGB888" if wl["nc"] == 4 and ref_nc == 3 else "")
    return dict(value=sample_px / dt / 1e6, unit="Mpixels/s", cores=threads, kind="reference" if use_ref else "port",
                sample=sample, ms_per_step=dt * 1e3, flags="-O2 -std=c++14 -DIS_LITTLE_ENDIAN" if use_ref else "-O2 -std=c99")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    cpu = cpu_reference_throughput(args.workload, steps, warmup)
    line = {
        "impl": "reference", "metric": metric_name(args.workload),
        "value": cpu["value"], "unit": "Mpixels/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": cpu["ms_per_step"], "higher_is_better": True, "scaling": "strong" if args.workload != "pvrtc2_rgba8" else "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": args.workload, "image": "%dx%d" % (wl["n"], wl["n"]), "where": "host CPU, %d threads" % cpu["cores"]},
        "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "flags")},
        "e2e": {"value": cpu["value"], "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------

class Env:
    """torch / distributed plumbing shared by the legs of the GPU arm."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        import image_compression_b200 as icb
        self.torch, self.dist, self.icb = torch, dist, icb
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the encoder has no CPU path (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
                os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its version there)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            # A host-side meeting point: while rank 0 drives ALL GPUs through the one-call host API (end-to-end leg) the
            # other ranks must not sit in an NCCL barrier -- its kernel would spin on their GPUs and time-slice with
            # rank 0's work there.
            self.cpu_group = dist.new_group(backend="gloo")
        self.stream = torch.cuda.current_stream()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def cpu_barrier(self):
        """All GPUs idle, every rank met on the host (gloo): nothing of this job runs on any GPU afterwards."""
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.cpu_group)

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.world == 1:
            return int(x)
        t = self.torch.tensor([int(x)], device="cuda", dtype=self.torch.int64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return int(t.item())

    def timed(self, step, steps):
        """K steps back to back on the current stream between two CUDA events, bracketed by barrier + synchronize on
        both sides; returns (max over ranks of the elapsed ms, this rank's own ms)."""
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(self.stream)
        for i in range(steps):
            step(i)
        e1.record(self.stream)
        self.barrier()
        mine = e0.elapsed_time(e1)
        return self.max_over_ranks(mine), mine

    def warm(self, step, warmup, spin_s=0.05):
        """W untimed steps, then keep stepping until ~50 ms have passed so that the timed region starts with the clocks
        already up instead of inside the GPU's ramp from idle (the timed region itself is only a few ms long)."""
        for i in range(warmup):
            step(i)
        self.torch.cuda.synchronize()
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < spin_s:
            for i in range(16):
                step(i)
            self.torch.cuda.synchronize()


class StripeJob:
    """One image of workload `wl`; this rank holds pixel rows of block rows [r0, r1) in `nbuf` rotating buffers (a buffer is
    reused only after >= 1.5 x the 126 MB L2 of other traffic has passed) and encodes them per step."""

    def __init__(self, env, wl, r0, r1, content=None, nbuf=None):
        torch, icb = env.torch, env.icb
        self.env, self.wl, self.r0, self.r1 = env, wl, r0, r1
        n, nc = wl["n"], wl["nc"]
        self.pitch = n * nc
        self.in_bytes = (r1 - r0) * 4 * self.pitch
        self.out_bytes = int((r1 - r0) * 4 * n * wl["out_bpp"])
        per_step = max(1, self.in_bytes + self.out_bytes)
        self.nbuf = nbuf or max(2, 1 + -(-3 * L2_BYTES // (2 * per_step)))
        self.srcs = [torch.empty(max(16, self.in_bytes), dtype=torch.uint8, device="cuda") for _ in range(self.nbuf)]
        self.dsts = [torch.empty(max(16, self.out_bytes), dtype=torch.uint8, device="cuda") for _ in range(self.nbuf)]
        for i, s in enumerate(self.srcs):
            if content is None:
                if self.in_bytes:
                    icb.fill_synthetic(s[:self.in_bytes], wl["seed"] + 16 * i, byte_offset=r0 * 4 * self.pitch)
            else:
                s[:self.in_bytes].copy_(content(i, r0 * 4, (r1 - r0) * 4))
        self.scratch = torch.empty(icb.lib().icb_pvrtc2_scratch_size(n, n), dtype=torch.uint8, device="cuda") if wl["codec"] == 3 else None
        self.peer_out = None  # raw address of this stripe inside the root's peer-mapped stream

    def step(self, i):
        icb, wl, n = self.env.icb, self.wl, self.wl["n"]
        if self.r1 == self.r0:
            return
        s = self.srcs[i % self.nbuf]
        if wl["codec"] == 3:
            icb.pvrtc_encode_device(s, n, n, out=self.dsts[i % self.nbuf], scratch=self.scratch, stream=self.env.stream)
            return
        out = self.peer_out if self.peer_out is not None else self.dsts[i % self.nbuf]
        # virtual address of pixel (0,0) of the whole image; rows outside the stripe are never touched
        icb.encode_stripe_device(wl["codec"], wl["fmt"], s.data_ptr() - self.r0 * 4 * self.pitch, n, n, self.pitch, n, n,
                                 self.r0, self.r1, out, stream=self.env.stream)


def host_image(workload, buffer_index=0):
    """The whole synthetic image of buffer 0 on the host (what the CPU reference is fed for the parity flag)."""
    import checkers as ck
    wl = WORKLOADS[workload]
    return ck.synthetic(wl["n"] * wl["n"] * wl["nc"], wl["seed"] + 16 * buffer_index)


def parity_record(workload, got_blocks, host_pixels=None):
    """Byte-compares `got_blocks` (numpy) with the reference CPU encoder's output for the same input."""
    import numpy as np

    import checkers as ck
    wl = WORKLOADS[workload]
    t0 = time.perf_counter()
    src = host_image(workload) if host_pixels is None else host_pixels
    want, kind = ck.cpu_encode_full(workload, src, wl["n"], wl["n"])
    same = got_blocks.size == want.size and bool(np.array_equal(got_blocks, want))
    return {"equal": same, "bytes_compared": int(want.size), "against": kind, "fnv1a64": "%016x" % ck.fnv1a64(want),
            "cpu_seconds": round(time.perf_counter() - t0, 2)}


def hbm_roofline(workload, algo_bytes, in_bytes, kernel_ms, extra=None):
    peak, peak_src = measured_peak_gbs()
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
         "traffic": profile_constant("traffic.json", workload), "traffic_source": "ncu --set full capture of this kernel, profiles/traffic.json (not re-measured in the run)",
         "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms_avg": kernel_ms,
         "read_only_frac": (in_bytes / (kernel_ms * 1e-3) / 1e9) / peak}
    if extra:
        r.update(extra)
    return r


def int_alu_roofline(workload, kernel_ms, sm_mhz):
    """ETC1's real bound: warp-instructions per launch on the integer pipe and on the
    FMA-heavy pipe (IDP.4A / IMAD) and in all (ncu, profiles/pipe_counts.json) / kernel time against the measured pipe
    rates (profiles/*_b200_pipe_rates.txt: 2.0 warp-instr/clk/SM for LOP3/PRMT/VIMNMX/VABSDIFF, 2.0 for IDP.4A/IMAD,
    4.0 issue slots) x 148 SMs x the SM clock sampled in this run.  `frac` is the integer pipe's, as in round 1."""
    counts = profile_constant("pipe_counts.json", workload)
    if not counts or not sm_mhz:
        return None
    peak = 2.0 * 148 * sm_mhz * 1e6 / 1e9  # G warp-instr/s
    achieved = counts["pipe_alu_warp_inst"] / (kernel_ms * 1e-3) / 1e9
    r = {"bound": "int_alu", "achieved": achieved, "peak": peak, "unit": "G warp-instr/s", "frac": achieved / peak,
         "pipe_alu_warp_inst_per_launch": counts["pipe_alu_warp_inst"], "inst_source": counts.get("source"),
         "peak_source": "tools/microbench/pipe_rates.cu on B200 (2.0 warp-instr/clk/SM) x 148 SMs x %.0f MHz sampled under load" % sm_mhz}
    if "pipe_fmaheavy_warp_inst" in counts:
        r["fmaheavy_frac"] = counts["pipe_fmaheavy_warp_inst"] / (kernel_ms * 1e-3) / 1e9 / peak
        r["issue_frac"] = counts["warp_inst"] / (kernel_ms * 1e-3) / 1e9 / (2.0 * peak)
        r["pipe_fmaheavy_warp_inst_per_launch"] = counts["pipe_fmaheavy_warp_inst"]
        r["warp_inst_per_launch"] = counts["warp_inst"]
    return r


def structured_content(torch, kind, n, nc):
    """Structured 8192^2 images generated on the device (flat regions, dark gradients, 2-colour cells: what the
    warp-uniform fast path of the DXT index search does NOT cover).  Returns content(buffer, row0, rows) -> uint8 rows."""
    def make(i, row0, rows):
        yy = (torch.arange(rows, device="cuda", dtype=torch.int32) + row0)[:, None]
        xx = torch.arange(n, device="cuda", dtype=torch.int32)[None, :]
        if kind == "flat":  # 64-px constant tiles with a little dither in one channel
            r = ((xx // 64) * 37 + (yy // 64) * 91 + i) % 256 + 0 * yy
            g = ((xx // 64) * 53 + (yy // 64) * 17) % 256 + 0 * yy
            b = ((xx // 64) * 11 + (yy // 64) * 7) % 256 + ((xx ^ yy) & 1)
        elif kind == "dark":  # slow gradients near black
            r = (xx // 256 + i) % 24 + 0 * yy
            g = (yy // 256) % 24 + 0 * xx
            b = ((xx + yy) // 512) % 16
        elif kind == "checker":  # two colours, cell 8 px
            c = ((xx // 8) ^ (yy // 8)) & 1
            r, g, b = c * 255, c * 255, c * 255
        else:  # "gradient": smooth ramps, every block regular
            r = (xx // 32 + i) % 256 + 0 * yy
            g = (yy // 32) % 256 + 0 * xx
            b = ((xx + yy) // 64) % 256
        chans = [r, g, b] + ([255 - (r % 256)] if nc == 4 else [])
        return torch.stack([c.expand(rows, n) for c in chans], -1).to(torch.uint8).contiguous().view(-1)
    return make


def measure_single(env, workload, steps, warmup, content_kind=None, want_parity=True, isolated=0):
    """One GPU, whole image: kernel ms (K launches back to back between two events), optional isolated-launch pass,
    parity of buffer 0's output against the CPU reference."""
    torch, icb = env.torch, env.icb
    wl = WORKLOADS[workload]
    n = wl["n"]
    content = structured_content(torch, content_kind, n, wl["nc"]) if content_kind else None
    job = StripeJob(env, wl, 0, n // 4, content=content)
    env.warm(job.step, warmup)
    before = icb.launch_count()
    _, ms = env.timed(job.step, steps)
    launches = icb.launch_count() - before
    rec = {"kernel_ms": ms / steps, "mpix_s": n * n / (ms / steps * 1e-3) / 1e6, "steps": steps, "launches_per_step": launches // max(1, steps),
           "buffers": job.nbuf}
    iso = []
    if isolated:
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(isolated)]
        for i in range(isolated):
            ev[i][0].record(env.stream)
            job.step(i)
            ev[i][1].record(env.stream)
        torch.cuda.synchronize()
        iso = sorted(a.elapsed_time(b) for a, b in ev)
    if want_parity:
        job.step(0)
        torch.cuda.synchronize()
        host_px = job.srcs[0][:job.in_bytes].cpu().numpy() if content_kind else None
        rec["parity"] = parity_record(workload, job.dsts[0][:job.out_bytes].cpu().numpy(), host_px)
    return job, rec, iso, launches


def unaligned_source_record(env, workload):
    """The same image at a device address that is NOT 16-byte aligned (base + 4 bytes), which TMA cannot describe: the
    whole image goes through the generic per-block kernel (clamped byte loads).  Reported so that the cost of that
    path is visible next to the TMA path's."""
    torch, icb = env.torch, env.icb
    wl = WORKLOADS[workload]
    if wl["codec"] == 3:
        return None
    n, nc = wl["n"], wl["nc"]
    nbytes = n * n * nc
    bufs = [torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda") for _ in range(2)]
    outs = [torch.empty(int(n * n * wl["out_bpp"]), dtype=torch.uint8, device="cuda") for _ in range(2)]
    srcs = [b[4:4 + nbytes] for b in bufs]
    for i, s_ in enumerate(srcs):
        icb.fill_synthetic(s_, wl["seed"] + 16 * i)
    rec = {"what": "%s, source base = 16-byte boundary + 4 (generic kernel)" % wl["desc"]}
    step = lambda i: icb.encode_device(wl["codec"], wl["fmt"], srcs[i % 2], n, n, out=outs[i % 2], stream=env.stream)
    env.warm(step, 3, spin_s=0.02)
    _, ms = env.timed(step, 10)
    rec["kernel_ms"] = ms / 10
    step(0)
    torch.cuda.synchronize()
    rec["parity"] = parity_record(workload, outs[0].cpu().numpy())
    return rec


def other_workloads(env, headline, clocks_mhz):
    """Compact records for the BASELINE configurations the headline does not cover, and for structured content on the
    headline workload (the DXT index search is data dependent).  N = 1 only; ~20 steps each."""
    out = {}
    for name in ("dxt5_rgba8", "dxt1_rgb8", "etc1_rgb8", "pvrtc2_rgba8", "dxt1_rgba8"):
        if name == headline:
            continue
        wl = WORKLOADS[name]
        job, rec, _, _ = measure_single(env, name, 20, 3)
        algo = job.in_bytes + job.out_bytes
        rec["roofline"] = hbm_roofline(name, algo, job.in_bytes, rec["kernel_ms"])
        for k in ("traffic_source", "peak_source"):
            rec["roofline"].pop(k, None)
        if name == "etc1_rgb8":
            rec["roofline_int_alu"] = int_alu_roofline(name, rec["kernel_ms"], clocks_mhz)
            rec["bound"] = ("instruction issue (exhaustive search: ~1024 colour-distance evaluations per block, split evenly between the integer pipe "
                            "and the IDP.4A/IMAD pipe: roofline_int_alu.frac / fmaheavy_frac / issue_frac); the HBM fraction is reported as required, not attainable")
        elif name == "pvrtc2_rgba8":
            rec["bound"] = "hbm (fused-ideal bytes 4.25 B/px); three kernels per step, integer-pipe / latency limited today"
        rec["desc"] = wl["desc"]
        del job
        env.torch.cuda.empty_cache()
        out[name] = rec
    out["unaligned_device_source"] = unaligned_source_record(env, headline)
    if headline in ("dxt1_rgba8", "dxt5_rgba8"):
        content = {}
        for kind in ("gradient", "flat", "dark", "checker"):
            job, rec, _, _ = measure_single(env, headline, 20, 3, content_kind=kind)
            rec["roofline_frac"] = hbm_roofline(headline, job.in_bytes + job.out_bytes, job.in_bytes, rec["kernel_ms"])["frac"]
            content[kind] = rec
            del job
            env.torch.cuda.empty_cache()
        out["content_dependence_" + headline] = content
    return out


def run_gpu_arm(args):
    import ctypes as C

    import numpy as np

    env = Env()
    torch, dist, icb = env.torch, env.dist, env.icb
    world, rank = env.world, env.rank
    wl = WORKLOADS[args.workload]
    n, nc, codec, fmt = wl["n"], wl["nc"], wl["codec"], wl["fmt"]
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    grid_rows, grid_cols = n // 4, n // 4
    block_bytes = 16 if codec == 1 else 8
    replicas = codec == 3 and world > 1  # PVRTC: square power-of-two images only -> N independent replicas
    L = icb.lib()

    # ---- partition of the ONE image over the ranks
    if world == 1 or replicas:
        splits, share = [0, grid_rows] if world == 1 else None, None
    else:
        share = {"even": -1, "auto": icb.root_share_permille(codec, fmt, world)}.get(args.partition)
        if share is None:
            share = int(args.partition)
        splits = icb.stripe_partition(world, grid_rows, share)
    r0, r1 = (0, grid_rows) if (world == 1 or replicas) else (splits[rank], splits[rank + 1])
    job = StripeJob(env, wl, r0, r1)
    total_out = grid_rows * grid_cols * block_bytes if codec != 3 else n * n // 4

    ps = None
    if world > 1 and not replicas:
        from image_compression_b200 import sharding
        ps = sharding.PeerStream(total_out, dst=0)  # collective; raises on every rank or on none
        job.peer_out = ps.stripe_ptr(r0 * grid_cols * block_bytes)

    sampler = ClockSampler(env.local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    env.barrier()
    env.warm(job.step, warmup)
    launches_before = icb.launch_count()
    # Timed region: exactly K steps back to back on one stream between two CUDA events, max over ranks.  (Consecutive
    # launches overlap their launch latency through programmatic dependent launch; nothing else sits between them.)  At
    # N > 1 a step ends when this rank's blocks have been stored into rank 0's buffer (peer stores complete with the kernel).
    total_ms, my_ms = env.timed(job.step, steps)
    launches = env.sum_over_ranks(icb.launch_count() - launches_before)
    ms_per_step = total_ms / steps
    job_px = n * n * (world if replicas else 1)
    value = job_px / (ms_per_step * 1e-3) / 1e6

    # isolated launches (each between its own two events; the events keep consecutive launches from overlapping)
    iso_n = min(steps, 20)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iso_n)]
    for i in range(iso_n):
        ev[i][0].record(env.stream)
        job.step(i)
        ev[i][1].record(env.stream)
    env.barrier()
    kernel_iso = sorted(a.elapsed_time(b) for a, b in ev)
    # The timed region lasts milliseconds, shorter than one nvidia-smi sample: keep the identical launches going for
    # ~0.5 s (untimed) so that the clock / throttle record is taken under this kernel's load, not at idle.
    t_hold = time.perf_counter()
    i = 0
    while time.perf_counter() - t_hold < 0.5:
        for _ in range(64):
            job.step(i)
            i += 1
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    # ---- parity of THIS run's output: the stream that arrived on rank 0 (buffer 0's input) vs the CPU reference
    job.step(0)
    parity = None
    if ps is not None:
        ps.complete()
        if rank == 0:
            parity = parity_record(args.workload, ps.tensor().cpu().numpy())
    else:
        torch.cuda.synchronize()
        if rank == 0:
            parity = parity_record(args.workload, job.dsts[0][:job.out_bytes].cpu().numpy())

    # ---- N > 1: the pieces that explain `value`
    multi = {}
    if ps is not None:
        from image_compression_b200 import sharding
        # (a) encode only: same stripes, blocks stored locally (no delivery)
        job.peer_out = None
        enc_ms, _ = env.timed(job.step, steps)
        multi["encode_only"] = {"ms_per_step": enc_ms / steps, "mpix_s": job_px / (enc_ms / steps * 1e-3) / 1e6,
                                "what": "same stripes, blocks stored to local HBM, nothing delivered (max over ranks)"}
        # (b) even split delivered the same way, when `value` used a root-heavy one
        even = icb.stripe_partition(world, grid_rows, -1)
        if even != splits:
            ej = StripeJob(env, wl, even[rank], even[rank + 1])
            ej.peer_out = ps.stripe_ptr(even[rank] * grid_cols * block_bytes)
            env.warm(ej.step, warmup, spin_s=0.02)
            e_ms, _ = env.timed(ej.step, steps)
            multi["even_split_delivered"] = {"ms_per_step": e_ms / steps, "mpix_s": job_px / (e_ms / steps * 1e-3) / 1e6,
                                             "bytes_into_root": total_out - (even[1] - even[0]) * grid_cols * block_bytes}
        else:
            ej = job
        # (c) the north star's literal form: encode locally, then ONE NCCL gather of the packed stream (even stripes)
        ej.peer_out = None
        local = ej.dsts[0][:ej.out_bytes]
        for _ in range(3):
            ej.step(0)
            sharding.gather_blocks(local, grid_rows, grid_cols, block_bytes, dst=0, splits=even)

        def gather_step(i):
            ej.step(0)
            sharding.gather_blocks(local, grid_rows, grid_cols, block_bytes, dst=0, splits=even)
        reps = min(steps, 10)
        g_ms, _ = env.timed(gather_step, reps)
        got = sharding.gather_blocks(local, grid_rows, grid_cols, block_bytes, dst=0, splits=even)
        multi["nccl_gather"] = {"ms_per_step": g_ms / reps, "mpix_s": job_px / (g_ms / reps * 1e-3) / 1e6,
                                "what": "encode to local HBM + dist.gather of the stripes to rank 0 (reference implementation of the delivery)"}
        if rank == 0:  # the gathered stream is buffer 0's, like the one the parity flag checked
            import checkers as ck
            multi["nccl_gather"]["equals_reference"] = ("%016x" % ck.fnv1a64(got.cpu().numpy())) == (parity or {}).get("fnv1a64")
        # (d) round 1's figure: every rank a full image of its own, kernel only -- weak scaling of independent kernels
        del ej
        wj = StripeJob(env, wl, 0, grid_rows)
        env.warm(wj.step, warmup, spin_s=0.02)
        w_ms, _ = env.timed(wj.step, min(steps, 20))
        multi["weak_scaling_kernel_only"] = {"ms_per_step": w_ms / min(steps, 20), "mpix_s": n * n * world / (w_ms / min(steps, 20) * 1e-3) / 1e6,
                                             "what": "%d independent %dx%d images, one per GPU, no delivery (round 1's `value`)" % (world, n, n)}
        del wj
        ps.close()
        torch.cuda.empty_cache()

    # ---- end to end: ONE image through the host-buffer API (pinned buffers; H2D + kernels + D2H per step)
    e2e = None
    in_total = n * n * nc
    env.barrier()
    env.cpu_barrier()
    if rank == 0:
        h_in_ptr, h_out_ptr = L.icb_host_alloc(in_total), L.icb_host_alloc(total_out)
        if not h_in_ptr or not h_out_ptr:
            raise SystemExit("pinned allocation failed")
        h_in = np.ctypeslib.as_array(C.cast(h_in_ptr, C.POINTER(C.c_uint8)), shape=(in_total,))
        h_out = np.ctypeslib.as_array(C.cast(h_out_ptr, C.POINTER(C.c_uint8)), shape=(total_out,))
        h_in[:] = host_image(args.workload)
        ctx = icb.ShardContext(list(range(world))) if (world > 1 and not replicas) else None
        call = (lambda: ctx.compress_host(codec, fmt, h_in, n, n, out=h_out)) if ctx else (lambda: icb.compress_host(codec, fmt, h_in, n, n, out=h_out))
        e2e_steps = max(3, min(steps, 10))
        for _ in range(2):
            call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            call()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        import checkers as ck
        e2e = {"value": n * n / (e2e_ms * 1e-3) / 1e6, "unit": "Mpixels/s", "h2d_bytes_per_step": in_total, "d2h_bytes_per_step": total_out,
               "ms_per_step": e2e_ms, "steps": e2e_steps,
               "api": "icb_ctx_compress_host: one call from rank 0 spreads the image over %d GPUs, each over its own PCIe link (pinned host buffers)" % world if ctx
               else "icb_compress_host (pinned host buffers)",
               "output_fnv1a64": "%016x" % ck.fnv1a64(h_out), "output_equals_reference": (parity or {}).get("fnv1a64") == "%016x" % ck.fnv1a64(h_out)}
        if ctx:
            ctx.close()
        L.icb_host_free(h_in_ptr)
        L.icb_host_free(h_out_ptr)
    env.cpu_barrier()
    env.barrier()

    if rank == 0:
        k_avg = my_ms / steps if world == 1 else multi.get("encode_only", {}).get("ms_per_step", ms_per_step)
        algo_bytes = job.in_bytes + job.out_bytes  # this rank's stripe: read every source byte once, write every block once
        roof = hbm_roofline(args.workload, algo_bytes, job.in_bytes, k_avg, {
            "kernel_ms_avg_source": "timed region: (end event - begin event) / steps, launches back to back" if world == 1
            else "rank 0's stripe kernel in the encode-only leg (the timed `value` also contains the delivery)",
            "kernel_ms_isolated_avg": sum(kernel_iso) / len(kernel_iso), "kernel_ms_isolated_median": kernel_iso[len(kernel_iso) // 2],
            "kernel_ms_min": kernel_iso[0], "kernel_ms_isolated_source": "separate pass, each launch between its own two events"})
        if world > 1:
            roof["traffic"] = None
        line = {
            "metric": metric_name(args.workload),
            "value": value, "unit": "Mpixels/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if codec == 3 else "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": args.workload, "image": "%dx%d" % (n, n),
                       "job": ("%d replicas of one %dx%d image (PVRTC does not shard)" % (world, n, n)) if replicas
                       else ("one %dx%d image" % (n, n)) + ("" if world == 1 else ", block-row stripes %s over %d GPUs, packed stream delivered to rank 0 inside the timed region" % (splits, world)),
                       "input": "splitmix64 byte stream, resident in HBM",
                       "l2": "%d rotating buffer pairs on rank 0, %d MB of other traffic between two uses of a buffer (126 MB L2)" % (job.nbuf, ((job.nbuf - 1) * (job.in_bytes + job.out_bytes)) >> 20),
                       "parallelism": "stripe%d" % world if not replicas else "replica%d" % world},
            "roofline": roof,
            "parity": parity,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if args.workload == "etc1_rgb8" and world == 1:
            line["roofline_int_alu"] = int_alu_roofline(args.workload, k_avg, (clocks or {}).get("sm_mhz"))
        if world > 1 and not replicas:
            into_root = total_out - (r1 - r0) * grid_cols * block_bytes
            gbs = into_root / (ms_per_step * 1e-3) / 1e9
            line["delivery"] = {
                "bound": "nvlink_ingress", "bytes_into_root": into_root, "achieved": gbs, "peak": NVLINK_INGRESS_GBS, "unit": "GB/s",
                "frac": gbs / NVLINK_INGRESS_GBS, "partition": "root share %s permille" % share if share is not None and share >= 0 else "even",
                "how": "encoder block stores go to rank 0's buffer through a CUDA-IPC peer mapping (NVLink); no separate gather pass",
                "limiter": "rank 0's NVLink ingress: every block of the other ranks' stripes crosses one GPU's inbound links; "
                           "the root encodes its own (larger) stripe while they arrive" if world > 2 else
                           "encode kernels (two ranks: the transfer of one half hides behind the encode of the other)"}
            line.update(multi)
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_reference_throughput(args.workload, 2, 1)
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "flags")}
        if world == 1 and not args.no_others:
            del job
            torch.cuda.empty_cache()
            line["other_workloads"] = other_workloads(env, args.workload, (clocks or {}).get("sm_mhz"))
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dxt1_rgba8", choices=sorted(WORKLOADS))
    ap.add_argument("--partition", default="auto", help="N > 1: auto (icb_root_share_permille), even, or the root's share in permille")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the other_workloads records (N = 1)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
