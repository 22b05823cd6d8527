#!/usr/bin/env python
"""bench.py — encode MB/s (bit-exact output) on enwik8-shaped input, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size-mb 100] [--level 0]

N = 1 workload = BASELINE.json configs[1]: "enwik8-shaped 100 MB, level e0, 1xB200 encode, bit-exact vs CPU".
A step = one encode of the whole 100 000 000-byte stream (6 blocks of 16 MiB, one parse chain each).
  value   MB/s (1 MB = 1e6 input bytes) with the input already resident in HBM and the framed output left in HBM
          (zlb_encode_blocks_device), timed with the engine's CUDA events on the stream the kernels run on.
  e2e     same metric through the host-buffer entry point the C++ drop-in API uses (zlb_encode_blocks):
          pinned host input -> H2D -> kernels -> D2H of the framed stream, all inside the timed region.
  N > 1   one process per GPU (torchrun); the path shards by STREAM (blocks of one stream are
          coupled by the MTF carry), so every rank encodes its own 100 MB stream (weak scaling, no data-path
          collective) and the packed outputs are gathered to rank 0 with one NCCL all_gather per step in the e2e leg.
  --impl reference   times the reference's own CPU encoder (oracle/_ref when built, else the C restatement) on the
          host cores, same workload, same JSON shape.
Before any timing the GPU output of the workload is compared byte-for-byte with the reference CPU encoder.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "encode MB/s (bit-exact output) on enwik8-shaped input"
PEAK_FALLBACK_GBS = 6650.0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return PEAK_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def load_traffic(nbytes, level):
    """dram__bytes_read.sum + dram__bytes_write.sum of the parse kernel from the committed ncu --set full capture
    (profiles/r1_parse_traffic.json), valid for the workload it was captured on only; None otherwise"""
    p = os.path.join(ROOT, "profiles", "r1_parse_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            t = json.load(f)
        if int(t.get("workload_bytes", -1)) == int(nbytes) and int(t.get("level", -1)) == int(level):
            return int(t["dram_bytes_read"] + t["dram_bytes_write"]), t.get("source")
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference(data, level, want_bytes=True):
    """the reference's CPU encoder on this box: (seconds, compressed bytes, kind)"""
    from _libs import Oracle, Ref, have_ref
    if have_ref():
        lib, kind = Ref(), "reference"
    else:
        lib, kind = Oracle(), "port"
    t = time.perf_counter()
    z = lib.encode(data, level)
    return time.perf_counter() - t, z, kind


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU encoder on the host cores, same workload as our arm at N GPUs —
    N independent streams, one thread each (the reference codec has no threading inside a stream)."""
    if rank != 0:
        return
    from libzling_b200 import corpus
    nbytes = int(args.size_mb * 1e6)
    streams = [corpus.enwik8_shaped(nbytes, seed=8 + r) if args.corpus == "enwik8" else corpus.mixed(nbytes, seed=4 + r) for r in range(world)]
    threads_used = min(world, os.cpu_count() or 1)
    out = [None] * world
    kinds = []

    def work(i):
        dt, z, kind = cpu_reference(streams[i], args.level)
        out[i] = len(z)
        kinds.append(kind)

    def step():
        t0 = time.perf_counter()
        for lo in range(0, world, threads_used):
            th = [threading.Thread(target=work, args=(i,)) for i in range(lo, min(world, lo + threads_used))]
            for t in th:
                t.start()
            for t in th:
                t.join()
        return time.perf_counter() - t0

    cpu_reference(streams[0][: min(nbytes, 8 << 20)], args.level)          # page the library in
    for _ in range(min(args.warmup, 1)):
        step()
    times = [step() for _ in range(args.steps)]
    dt = float(np.mean(times))
    mbs = world * nbytes / 1e6 / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": round(mbs, 3), "unit": "MB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "enwik8-shaped %d B, level e%d, encode, one stream per GPU (BASELINE.json configs[1])" % (nbytes, args.level),
                   "bytes_per_gpu": int(nbytes), "level": args.level, "streams": world, "compressed_bytes": int(out[0])},
        "cpu_baseline": {"value": round(mbs, 3), "unit": "MB/s", "cores": threads_used, "kind": kinds[0] if kinds else "reference",
                         "sample": "the whole workload (%d stream(s) of %d B), %d step(s), one thread per stream; %d host cores present"
                                   % (world, nbytes, args.steps, os.cpu_count())},
        "e2e": {"value": round(mbs, 3), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size-mb", type=float, default=100.0)
    ap.add_argument("--level", type=int, default=0)
    ap.add_argument("--skip-parity", action="store_true")
    ap.add_argument("--no-decode", action="store_true", help="skip the secondary decode leg")
    ap.add_argument("--corpus", default="enwik8", choices=["enwik8", "mixed"], help="mixed = BASELINE.json configs[3] (text + binary + random)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    nbytes = int(args.size_mb * 1e6)

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    from libzling_b200 import corpus
    data = corpus.enwik8_shaped(nbytes, seed=8 + rank) if args.corpus == "enwik8" else corpus.mixed(nbytes, seed=4 + rank)

    import torch
    import libzling_b200
    from libzling_b200 import build as zbuild
    if not os.path.exists(libzling_b200.lib_path()):
        zbuild.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this path has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    nblocks = (nbytes + libzling_b200.BLOCK - 1) // libzling_b200.BLOCK
    ctx = libzling_b200.Context(device=local, max_blocks=nblocks)
    L = libzling_b200.load()
    out_cap = L.zlb_encode_bound(nbytes)

    # ---- parity gate: GPU bytes == reference CPU bytes on this exact workload (rank 0 also times the CPU here)
    cpu_dt, want, cpu_kind = cpu_reference(data, args.level)
    got = ctx.encode(data, args.level)
    if not args.skip_parity and got != want:
        raise SystemExit("bench.py: GPU output differs from the CPU reference (%d vs %d bytes) — refusing to report a number" % (len(got), len(want)))
    ratio = len(want) / nbytes

    # ---- device-resident leg (value)
    d_in = torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda")
    d_in[:nbytes].copy_(torch.from_numpy(data))
    d_in[nbytes:].zero_()
    d_out = torch.empty(out_cap, dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2: written between timed steps
    pin_in = libzling_b200.PinnedBuffer(nbytes)
    pin_in.array[:] = data
    pin_out = libzling_b200.PinnedBuffer(out_cap)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def step_device():
        enc = libzling_b200.Encoder(ctx, args.level)
        n = enc.encode_blocks_device(d_in.data_ptr(), nbytes, d_out.data_ptr(), out_cap)
        enc.close()
        return n, ctx.stats()

    def step_host():
        enc = libzling_b200.Encoder(ctx, args.level)
        n = enc.encode_blocks_into(pin_in.array, pin_out.array)
        enc.close()
        return n, ctx.stats()

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    sampler.start()
    dev_ms, parse_ms, mtf_ms, build_ms, pack_ms, launches = [], [], [], [], [], 0
    barrier()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        n, st = step_device()
        dev_ms.append(st["ms_total"]); parse_ms.append(st["ms_parse"]); mtf_ms.append(st["ms_mtf"])
        build_ms.append(st["ms_huff_build"]); pack_ms.append(st["ms_pack"]); launches += st["launches"]
        assert n == len(want)
    barrier()
    wall_dev = time.perf_counter() - wall0
    last = st

    # ---- end-to-end leg: host buffers through the C-ABI call the C++ drop-in API makes
    for _ in range(max(1, args.warmup - 2)):
        step_host()
    e2e_s = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        n, st_h = step_host()
        ms = st_h["ms_total"]            # engine events on its own stream: H2D -> kernels -> D2H of the framed stream
        if dist is not None:             # single gather of the packed outputs over NCCL (sizes, then padded payloads)
            ev0.record()
            sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
            dist.all_gather(sizes, torch.tensor([n], dtype=torch.int64, device="cuda"))
            mx = int(max(int(x.item()) for x in sizes))
            mine = torch.zeros(mx, dtype=torch.uint8, device="cuda")
            mine[:n].copy_(torch.from_numpy(pin_out.array[:n]))
            bufs = [torch.empty(mx, dtype=torch.uint8, device="cuda") for _ in range(world)] if rank == 0 else None
            dist.gather(mine, bufs, dst=0)
            ev1.record()
            torch.cuda.synchronize()
            ms += ev0.elapsed_time(ev1)
        e2e_s.append(ms / 1e3)
        assert n == len(want)
    barrier()
    clocks = sampler.stop()

    # ---- secondary: decode (BASELINE.json configs[4] shape) on a bounded sample — the first 16 MiB block of the stream.
    # One chain per stream (MTF state + context dependence): reported, not optimised for, next to the CPU decoder.
    decode = None
    if rank == 0 and not args.no_decode:
        from _libs import Ref, Oracle, have_ref, bound  # noqa: F401
        sample = data[: min(nbytes, libzling_b200.BLOCK)]
        zs = ctx.encode(sample, args.level)
        back = ctx.decode(zs)                                             # warm-up + round-trip check
        if back != sample.tobytes():
            raise SystemExit("bench.py: GPU decode round trip failed")
        dts = []
        for _ in range(2):
            ctx.decode(zs)
            dts.append(ctx.stats()["ms_total"])
        lib = Ref() if have_ref() else Oracle()
        t0 = time.perf_counter()
        lib.decode(zs, sample.size)
        cpu_dec = time.perf_counter() - t0
        decode = {"value": round(sample.size / 1e6 / (min(dts) / 1e3), 3), "unit": "MB/s (decoded bytes)", "sample": "first %d B of the stream, 1 block" % sample.size,
                  "ms": round(min(dts), 3), "cpu_reference_mbs": round(sample.size / 1e6 / cpu_dec, 3),
                  "note": "single-stream decode is one serial chain (DESIGN.md 4.5)"}
    assert bytes(pin_out.array[:n]) == want or args.skip_parity

    ms_step = float(np.mean(dev_ms))
    e2e_step = float(np.mean(e2e_s))
    if dist is not None:
        t = torch.tensor([ms_step, e2e_step], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_step = float(t[0]), float(t[1])
    value = world * nbytes / 1e6 / (ms_step / 1e3)
    e2e_value = world * nbytes / 1e6 / e2e_step
    peak, peak_src = load_peaks()
    traffic, traffic_src = load_traffic(nbytes, args.level)
    pms = float(np.mean(parse_ms))
    ach = nbytes / 1e9 / (pms / 1e3) if pms > 0 else 0.0      # algorithmic bytes of the parse launch: every input byte read once
    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": "MB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "%s %d B, level e%d, encode, one stream per GPU%s" % ("enwik8-shaped" if args.corpus == "enwik8" else "mixed text+binary+random", nbytes, args.level,
                                                                                          " (BASELINE.json configs[1])" if (args.corpus, nbytes, args.level) == ("enwik8", 100000000, 0) else ""),
                       "bytes_per_gpu": int(nbytes), "level": args.level, "blocks_per_gpu": int(nblocks), "compressed_bytes": len(want),
                       "ratio": round(ratio, 4), "bit_exact_vs_cpu_reference": not args.skip_parity,
                       "l2": "256 MB buffer written between timed steps (L2 flush); working set (input + 12 MB bucket state/block + tokens) exceeds L2",
                       "parse_kernel": "zl_rolz_parse_v%s" % os.environ.get("ZLB_PARSE", "4")},
            "e2e": {"value": round(e2e_value, 3), "unit": "MB/s", "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(len(want)),
                    "ms_per_step": round(e2e_step * 1e3, 3)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "zl_rolz_parse_v%s (one launch per step, 1 CTA per 16 MiB block)" % os.environ.get("ZLB_PARSE", "4"),
                         "achieved": round(ach, 4), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 6), "traffic": traffic, "traffic_unit": "bytes per launch (dram read + write)",
                         "traffic_source": traffic_src, "algorithmic_bytes_per_launch": int(nbytes),
                         "peak_source": peak_src,
                         "note": "algorithmic bytes = input bytes (each read once); the kernel is bound by the serial token chain, not HBM "
                                 "(DESIGN.md §4): %d tokens in %d chains" % (last["tokens"], nblocks)},
            "kernel_ms": {"parse": round(pms, 3), "mtf": round(float(np.mean(mtf_ms)), 3), "huff_build": round(float(np.mean(build_ms)), 3),
                          "pack": round(float(np.mean(pack_ms)), 3), "wall_ms_per_step_incl_flush": round(wall_dev / args.steps * 1e3, 3)},
            "parse_counters": {k: int(last[k]) for k in ("tokens", "subblocks", "slow_main", "slow_lazy", "general_path", "window_hits", "windows", "reparsed_blocks",
                                                          "cyc_spec", "cyc_resolve", "cyc_total", "flagged", "rounds", "cyc_final", "cyc_orbit", "cyc_rank", "cyc_decide")},
            "decode": decode,
            "cpu_baseline": {"value": round(nbytes / 1e6 / cpu_dt, 3), "unit": "MB/s", "cores": 1, "kind": cpu_kind,
                             "sample": "the whole %d-byte workload once, single thread (the reference codec has no threading); %d host cores present" % (nbytes, os.cpu_count())},
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
