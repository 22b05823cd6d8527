#!/usr/bin/env python
"""bench.py — encode MB/s (bit-exact output) on enwik8-shaped input, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--mode encode|decode] [--shard streams|stream] [--size-mb 100] [--level 0] [--corpus enwik8|mixed]

Default (N = 1) workload = BASELINE.json configs[1]: "enwik8-shaped 100 MB, level e0, 1xB200 encode, bit-exact vs CPU".
A step = one encode of the whole 100 000 000-byte stream (6 blocks of 16 MiB).
  value   MB/s (1 MB = 1e6 input bytes) with the input already resident in HBM and the framed output left in HBM
          (zlb_encode_blocks_device), timed with the engine's CUDA events on the stream the kernels run on.
  e2e     the same metric through the host-buffer calls the C++ drop-in API makes, HOST WALL CLOCK around
          zlb_encoder_begin -> zlb_encode_blocks (pinned host input -> H2D -> kernels -> D2H of the framed stream) ->
          zlb_encoder_end; at N > 1 the call is zlb_encode_blocks_gathered: the framed outputs stay in HBM and reach
          rank 0 through ONE NCCL gather (sizes by all-gather of u64, payloads by one send/recv group), rank 0 copies
          all of them to host memory — all inside the timed region, max over ranks.
  --shard streams (default)  one process per GPU (torchrun), every rank encodes its own stream of --size-mb (weak scaling).
  --shard stream             ONE stream of --size-mb over the N GPUs (BASELINE.json configs[3] with --corpus mixed
          --size-mb 1000 --level 2): contiguous 16 MiB block ranges per rank, the 65 540-byte carried state (MTF tables +
          level) GPU -> GPU by ncclSend/ncclRecv in block order, ONE gather of the framed ranges
          (zlb_encode_stream_sharded).  Strong scaling; the MTF chain of the stream is serial.
  --mode decode              BASELINE.json configs[4]: decode MB/s (decoded bytes) of the reference-encoded stream,
          next to the reference's CPU decoder; N > 1 = one stream per GPU (replicas: one stream is one chain).
  --impl reference           times the reference's own CPU implementation (oracle/_ref when built, else the C restatement) on
          the host cores, same workload, same JSON shape.
Before any timing the GPU output of the workload is compared byte-for-byte with the reference CPU implementation.
"""
import argparse
import hashlib
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
METRIC_DEC = "decode MB/s (decoded bytes, bit-exact) of the reference-encoded stream"
PEAK_FALLBACK_GBS = 6650.0
BLOCK = 16777216
CPU_SAMPLE = 200_000_000        # at most this many bytes of CPU work per stream and step (bounded sample)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return PEAK_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def lib_sha16():
    import libzling_b200
    with open(libzling_b200.lib_path(), "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()[:16]


def src_sha16():
    """the library's SOURCES (the nvcc build is not bit-reproducible, the sources are).  None when a source is newer than the library
    that is loaded (then the hash would not describe what runs)"""
    from libzling_b200 import build as zbuild
    return None if zbuild.needs_build() else zbuild.source_sha16()


def load_traffic(kernel, nbytes, level):
    """dram__bytes_read.sum + dram__bytes_write.sum of `kernel` from a committed `ncu --set full` capture
    (profiles/traffic_<kernel>.json, written by scripts/ncu_traffic.py); accepted only for the workload it was captured
    on AND when the capture was taken from the libzling.so that is loaded now (sha256 recorded beside it)"""
    p = os.path.join(ROOT, "profiles", "traffic_%s.json" % kernel)
    if os.path.exists(p):
        with open(p) as f:
            t = json.load(f)
        same_code = t.get("lib_sha16") == lib_sha16() or (t.get("src_sha16") is not None and t.get("src_sha16") == src_sha16())
        if int(t.get("workload_bytes", -1)) == int(nbytes) and int(t.get("level", -1)) == int(level) and same_code:
            return int(t["dram_bytes_read"] + t["dram_bytes_write"]), t.get("source")
        return None, "the capture in profiles/ is for another build or workload (sources %s, now %s): not used" % (t.get("src_sha16"), src_sha16())
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
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


def cpu_lib():
    from _libs import Oracle, Ref, have_ref
    return (Ref(), "reference") if have_ref() else (Oracle(), "port")


def make_stream(args, seed_off):
    from libzling_b200 import corpus
    nbytes = int(args.size_mb * 1e6)
    return corpus.enwik8_shaped(nbytes, seed=8 + seed_off) if args.corpus == "enwik8" else corpus.mixed(nbytes, seed=4 + seed_off)


def workload_name(args, world):
    nbytes = int(args.size_mb * 1e6)
    kind = "enwik8-shaped" if args.corpus == "enwik8" else "mixed text+binary+random"
    if args.mode == "decode":
        return "%s %d B encoded at level e%d by the reference, decode, one stream per GPU%s" % (
            kind, nbytes, args.level, " (BASELINE.json configs[4])" if nbytes == 1000000000 else "")
    tag = ""
    if (args.corpus, nbytes, args.level, args.shard) == ("enwik8", 100000000, 0, "streams"):
        tag = " (BASELINE.json configs[1])"
    elif (args.corpus, nbytes, args.level, args.shard) == ("enwik8", 100000000, 4, "streams"):
        tag = " (BASELINE.json configs[2])"
    elif (args.corpus, nbytes, args.level, args.shard) == ("mixed", 1000000000, 2, "stream"):
        tag = " (BASELINE.json configs[3])"
    if args.shard == "stream":
        return "%s %d B, level e%d, encode, ONE stream over the GPUs in contiguous 16 MiB block ranges%s" % (kind, nbytes, args.level, tag)
    return "%s %d B, level e%d, encode, one stream per GPU%s" % (kind, nbytes, args.level, tag)


def block_ranges(nbytes, world):
    nblocks = (nbytes + BLOCK - 1) // BLOCK
    base, extra = divmod(nblocks, world)
    out, b = [], 0
    for r in range(world):
        nb = base + (1 if r < extra else 0)
        out.append((min(b * BLOCK, nbytes), min((b + nb) * BLOCK, nbytes)))
        b += nb
    return out


# ----------------------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation on the host cores, same workload as our arm at N GPUs.
    shard=streams / decode: N independent streams, one thread each (the codec has no threading inside a stream);
    shard=stream: the one stream, one thread.  Each step works on a bounded sample (<= 200 MB per stream)."""
    if rank != 0:
        return
    nbytes = int(args.size_mb * 1e6)
    lib, kind = cpu_lib()
    nstreams = 1 if args.shard == "stream" else world
    sample = min(nbytes, CPU_SAMPLE)
    streams = [make_stream(args, r)[:sample] for r in range(nstreams)]
    comp = [lib.encode(s, args.level) for s in streams] if args.mode == "decode" else None
    threads_used = min(nstreams, os.cpu_count() or 1)
    sizes = [0] * nstreams

    def work(i):
        if args.mode == "decode":
            lib.decode(comp[i], streams[i].size)
            sizes[i] = len(comp[i])
        else:
            sizes[i] = len(lib.encode(streams[i], args.level))

    def step():
        t0 = time.perf_counter()
        for lo in range(0, nstreams, threads_used):
            th = [threading.Thread(target=work, args=(i,)) for i in range(lo, min(nstreams, lo + threads_used))]
            for t in th:
                t.start()
            for t in th:
                t.join()
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        step()
    times = [step() for _ in range(args.steps)]
    dt = float(np.mean(times))
    mbs = nstreams * sample / 1e6 / dt
    line = {
        "impl": "reference", "metric": METRIC if args.mode == "encode" else METRIC_DEC, "value": round(mbs, 3), "unit": "MB/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True,
        "scaling": "strong" if args.shard == "stream" else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args, world), "bytes_per_gpu": int(nbytes if args.shard != "stream" else nbytes // world), "level": args.level,
                   "streams": nstreams, "compressed_bytes": int(sizes[0]), "sample_bytes_per_stream": int(sample)},
        "cpu_baseline": {"value": round(mbs, 3), "unit": "MB/s", "cores": threads_used, "kind": kind,
                         "sample": "%d stream(s), the first %d B of each (%s), %d warm-up + %d timed step(s), one thread per stream; %d host cores present"
                                   % (nstreams, sample, "the whole workload" if sample == nbytes else "bounded sample of the %d-byte workload" % nbytes,
                                      args.warmup, args.steps, os.cpu_count())},
        "e2e": {"value": round(mbs, 3), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="encode", choices=["encode", "decode"])
    ap.add_argument("--shard", default=None, choices=["streams", "stream"],
                    help="streams: one stream per GPU (default); stream: ONE stream over all GPUs (default for BASELINE.json configs[3]: --corpus mixed --size-mb 1000)")
    ap.add_argument("--size-mb", type=float, default=100.0)
    ap.add_argument("--level", type=int, default=0)
    ap.add_argument("--streams", type=int, default=1, help="> 1: that many independent streams of --size-mb each per GPU in ONE batch call (zlb_encode_batch / zlb_decode_batch)")
    ap.add_argument("--skip-parity", action="store_true")
    ap.add_argument("--no-decode", action="store_true", help="skip the secondary decode leg of the encode mode")
    ap.add_argument("--corpus", default="enwik8", choices=["enwik8", "mixed"], help="mixed = BASELINE.json configs[3] (text + binary + random)")
    args = ap.parse_args()
    if args.shard is None:
        args.shard = "stream" if (args.corpus == "mixed" and args.size_mb >= 1000 and args.mode == "encode" and args.streams == 1) else "streams"

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    nbytes = int(args.size_mb * 1e6)

    if args.impl == "reference":
        # launched like our arm (torchrun at N > 1: rank 0 works, the others exit); a plain `python bench.py --impl reference
        # --gpus N` gives the same N-stream workload
        run_reference(args, rank, world if "WORLD_SIZE" in os.environ else max(1, args.gpus))
        return

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
        if "NCCL_DEBUG_FILE" not in os.environ:      # keep NCCL's INFO log (communicator sizes, transports) but off stdout:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)     # rank 0 prints exactly one JSON line there
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT,ENV")
            os.environ["NCCL_DEBUG_FILE"] = os.path.join(ROOT, "gpurun_out", "nccl_n%d_rank%d.log" % (world, rank))
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    if args.streams > 1:
        run_batch(args, rank, world, local, dist, torch, libzling_b200)
        return
    if args.mode == "decode":
        run_decode(args, rank, world, local, dist, torch, libzling_b200)
        return

    one_stream = args.shard == "stream"
    whole = make_stream(args, 0 if one_stream else rank)
    if one_stream:
        lo, hi = block_ranges(nbytes, world)[rank]
        data = whole[lo:hi]
    else:
        data = whole
    nlocal = int(data.size)
    nblocks = max(1, (nlocal + BLOCK - 1) // BLOCK)
    ctx = libzling_b200.Context(device=local, max_blocks=nblocks)
    L = libzling_b200.load()
    out_cap = L.zlb_encode_bound(nlocal)
    all_cap = L.zlb_encode_bound(nbytes) if one_stream else world * L.zlb_encode_bound(nbytes)

    comm = None
    if world > 1:
        def bcast(idb):
            t = torch.from_numpy(idb.copy()).cuda()
            dist.broadcast(t, src=0)
            return t.cpu().numpy()
        comm = libzling_b200.Comm(ctx, rank, world, bcast)

    pin_in = libzling_b200.PinnedBuffer(max(nlocal, 1))
    pin_in.array[:nlocal] = data
    pin_out = libzling_b200.PinnedBuffer(all_cap if (rank == 0 and world > 1) else max(out_cap, 1))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def step_host():
        """the user's call with host buffers; returns (bytes on this rank, seconds of host wall clock)"""
        t0 = time.perf_counter()
        enc = libzling_b200.Encoder(ctx, args.level)
        if comm is None:
            n = enc.encode_blocks_into(pin_in.array[:nlocal], pin_out.array)
        elif one_stream:
            n = comm.encode_stream(enc, pin_in.array[:nlocal], out=pin_out.array if rank == 0 else None)
        else:
            n, _ = comm.encode_gathered(enc, pin_in.array[:nlocal], out=pin_out.array if rank == 0 else None)
        enc.close()
        return n, time.perf_counter() - t0

    # ---- parity gate: GPU bytes == reference CPU bytes on this exact workload; rank 0 also times the CPU here (on the
    # whole stream when parity is checked, else on a bounded sample)
    lib, cpu_kind = cpu_lib()
    want, cpu_dt, cpu_bytes = None, None, 0
    if rank == 0:
        src = whole if (not args.skip_parity or nbytes <= CPU_SAMPLE) else whole[:CPU_SAMPLE]
        t0 = time.perf_counter()
        z = lib.encode(src, args.level)
        cpu_dt, cpu_bytes = time.perf_counter() - t0, int(src.size)
        want = z if src.size == whole.size else None
    n, _ = step_host()
    if not args.skip_parity and rank == 0:
        got = bytes(pin_out.array[:len(want)]) if (comm is not None and not one_stream) else bytes(pin_out.array[:n])
        if got != want:                  # (streams mode at N > 1: rank 0's own stream is the first one in the gathered buffer)
            raise SystemExit("bench.py: GPU output differs from the CPU reference (%d vs %d bytes) — refusing to report a number" % (len(got), len(want)))
    comp_len = len(want) if want is not None else int(n)

    # ---- device-resident leg (value): input in HBM, output left in HBM
    d_in = torch.empty(nlocal + 64, dtype=torch.uint8, device="cuda")
    d_in[:nlocal].copy_(torch.from_numpy(np.ascontiguousarray(data)))
    d_in[nlocal:].zero_()
    d_out = torch.empty(max(out_cap, 1), dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2: written between timed steps

    def step_device():
        enc = libzling_b200.Encoder(ctx, args.level)
        if comm is not None and one_stream:
            nn = comm.encode_stream(enc, nlocal, out=None, device_ptr=d_in.data_ptr())
        else:
            nn = enc.encode_blocks_device(d_in.data_ptr(), nlocal, d_out.data_ptr(), out_cap) if nlocal else 0
        enc.close()
        return nn, ctx.stats()

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    sampler.start()
    dev_ms, parse_ms, mtf_ms, build_ms, pack_ms, launches = [], [], [], [], [], 0
    barrier()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        t0 = time.perf_counter()
        nn, st = step_device()
        dt = time.perf_counter() - t0
        # ONE stream over N GPUs: a step ends when the last rank is done (host wall clock, max over ranks below);
        # otherwise the engine's CUDA events from the first to the last kernel of the call
        dev_ms.append(dt * 1e3 if (comm is not None and one_stream) else st["ms_total"])
        parse_ms.append(st["ms_parse"]); mtf_ms.append(st["ms_mtf"]); build_ms.append(st["ms_huff_build"]); pack_ms.append(st["ms_pack"])
        launches += st["launches"]
    barrier()
    wall_dev = time.perf_counter() - wall0
    last = st

    # ---- end-to-end leg: host buffers through the C-ABI calls the C++ drop-in API makes, host wall clock
    for _ in range(max(1, args.warmup - 2)):
        step_host()
    e2e_s = []
    barrier()
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        n, dt = step_host()
        e2e_s.append(dt)
    barrier()
    clocks = sampler.stop()
    shard_stats = comm.stats() if comm is not None else None

    # ---- secondary: decode of the first 16 MiB block of the stream next to the CPU decoder (full decode: --mode decode)
    decode = None
    if rank == 0 and not args.no_decode and nlocal:
        sample = np.ascontiguousarray(data[: min(nlocal, BLOCK)])
        dctx = libzling_b200.Context(device=local, max_blocks=1)
        zs = dctx.encode(sample, args.level)
        if dctx.decode(zs) != sample.tobytes():
            raise SystemExit("bench.py: GPU decode round trip failed")
        dts = []
        for _ in range(2):
            dctx.decode(zs)
            dts.append(dctx.stats()["ms_total"])
        t0 = time.perf_counter()
        lib.decode(zs, sample.size)
        cpu_dec = time.perf_counter() - t0
        decode = {"value": round(sample.size / 1e6 / (min(dts) / 1e3), 3), "unit": "MB/s (decoded bytes)", "sample": "first %d B of the stream, 1 block" % sample.size,
                  "ms": round(min(dts), 3), "cpu_reference_mbs": round(sample.size / 1e6 / cpu_dec, 3),
                  "note": "single-stream decode is one serial chain (DESIGN.md 4.5); full-size run: bench.py --mode decode"}
        dctx.close()

    ms_step = float(np.mean(dev_ms))
    e2e_step = float(np.mean(e2e_s))
    if dist is not None:
        t = torch.tensor([ms_step, e2e_step], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_step = float(t[0]), float(t[1])
    total_bytes = nbytes if one_stream else world * nbytes
    value = total_bytes / 1e6 / (ms_step / 1e3)
    e2e_value = total_bytes / 1e6 / e2e_step
    peak, peak_src = load_peaks()
    traffic, traffic_src = load_traffic("zl_rolz_parse_v4_kernel", nlocal, args.level)
    pms = float(np.mean(parse_ms))
    ach = nlocal / 1e9 / (pms / 1e3) if pms > 0 else 0.0      # algorithmic bytes of the parse launch: every input byte read once
    if rank == 0:
        pk = {k: int(last[k]) for k in ("tokens", "subblocks", "windows", "rounds", "reparsed_blocks", "parse_launches", "cyc_spec", "cyc_resolve", "cyc_final",
                                       "cyc_total", "cyc_orbit", "cyc_rank", "cyc_decide")}
        pk["rounds_per_window"] = round(pk["rounds"] / max(pk["windows"], 1), 3)
        pk["cyc_per_token"] = round(pk["cyc_total"] / max(pk["tokens"], 1), 1)
        pk["cyc_resolve_per_token"] = round(pk["cyc_resolve"] / max(pk["tokens"], 1), 1)
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": "MB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "strong" if one_stream else "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": workload_name(args, world), "bytes_per_gpu": int(nlocal), "level": args.level, "blocks_per_gpu": int(nblocks),
                       "compressed_bytes": int(comp_len), "ratio": round(comp_len / max(whole.size if want is not None else nlocal, 1), 4),
                       "bit_exact_vs_cpu_reference": not args.skip_parity,
                       "l2": "256 MB buffer written between timed steps (L2 flush); working set (input + 12 MB bucket state/block + tokens) exceeds L2",
                       "parse_kernel": "zl_rolz_parse_v%s" % os.environ.get("ZLB_PARSE", "4"), "shard": args.shard, "lib_sha16": lib_sha16(), "src_sha16": src_sha16()},
            "e2e": {"value": round(e2e_value, 3), "unit": "MB/s", "h2d_bytes_per_step": int(nlocal), "d2h_bytes_per_step": int(n),
                    "ms_per_step": round(e2e_step * 1e3, 3),
                    "timing": "host wall clock around zlb_encoder_begin -> %s -> zlb_encoder_end, pinned host buffers, max over ranks"
                              % ("zlb_encode_blocks" if comm is None else ("zlb_encode_stream_sharded" if one_stream else "zlb_encode_blocks_gathered"))},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "zl_rolz_parse_v4_kernel (one launch per step, 1 CTA of 1024 threads per 16 MiB block)",
                         "achieved": round(ach, 4), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 6), "traffic": traffic,
                         "traffic_unit": "bytes per launch (dram read + write)", "traffic_source": traffic_src, "algorithmic_bytes_per_launch": int(nlocal),
                         "peak_source": peak_src,
                         "note": "algorithmic bytes = input bytes (each read once); the kernel is bound by the per-block window chain "
                                 "(%d blocks = %d CTAs on 148 SMs), not by HBM (DESIGN.md §4)" % (nblocks, nblocks)},
            "kernel_ms": {"parse": round(pms, 3), "mtf": round(float(np.mean(mtf_ms)), 3), "huff_build": round(float(np.mean(build_ms)), 3),
                          "pack": round(float(np.mean(pack_ms)), 3), "wall_ms_per_step_incl_flush": round(wall_dev / args.steps * 1e3, 3)},
            "parse_counters": pk,
            "shard_stats": shard_stats,
            "decode": decode,
            "cpu_baseline": {"value": round(cpu_bytes / 1e6 / cpu_dt, 3), "unit": "MB/s", "cores": 1, "kind": cpu_kind,
                             "sample": "%d B of the workload once, single thread (the reference codec has no threading); %d host cores present" % (cpu_bytes, os.cpu_count())},
        }
        print(json.dumps(line), flush=True)
    if comm is not None:
        comm.close()
    if dist is not None:
        dist.destroy_process_group()


def run_decode(args, rank, world, local, dist, torch, libzling_b200):
    """--mode decode: every rank decodes its own reference-encoded stream (one stream = one chain: replicas)"""
    nbytes = int(args.size_mb * 1e6)
    data = make_stream(args, rank)
    lib, cpu_kind = cpu_lib()
    z = np.frombuffer(lib.encode(data, args.level), dtype=np.uint8)
    nblocks = (nbytes + BLOCK - 1) // BLOCK
    ctx = libzling_b200.Context(device=local, max_blocks=min(nblocks, 64))
    pin_z = libzling_b200.PinnedBuffer(z.size)
    pin_z.array[:] = z

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def step():
        t0 = time.perf_counter()
        out = ctx.decode(pin_z.array)
        return out, time.perf_counter() - t0, ctx.stats()

    out, _, _ = step()
    if out != data.tobytes():
        raise SystemExit("bench.py: GPU decode differs from the original data — refusing to report a number")
    for _ in range(max(0, args.warmup - 1)):
        step()
    sampler = ClockSampler(local)
    sampler.start()
    times, launches = [], 0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    barrier()
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        _, dt, st = step()
        times.append(dt)
        launches += st["launches"]
    barrier()
    clocks = sampler.stop()
    sample = min(nbytes, CPU_SAMPLE)
    zs = lib.encode(data[:sample], args.level) if sample < nbytes else z.tobytes()
    t0 = time.perf_counter()
    lib.decode(zs, sample)
    cpu_dt = time.perf_counter() - t0
    step_s = float(np.mean(times))
    if dist is not None:
        t = torch.tensor([step_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_s = float(t[0])
    value = world * nbytes / 1e6 / step_s
    peak, peak_src = load_peaks()
    ach = (nbytes + z.size) / 1e9 / step_s
    if rank == 0:
        line = {
            "metric": METRIC_DEC, "value": round(value, 3), "unit": "MB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(step_s * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(args, world), "bytes_per_gpu": int(nbytes), "level": args.level, "compressed_bytes": int(z.size),
                       "bit_exact": True, "l2": "256 MB buffer written between timed steps (L2 flush)", "lib_sha16": lib_sha16(),
                       "note": "one stream is one serial chain (MTF state + context dependence, DESIGN.md 4.5); N > 1 = independent replicas"},
            "e2e": {"value": round(value, 3), "unit": "MB/s", "h2d_bytes_per_step": int(z.size), "d2h_bytes_per_step": int(nbytes),
                    "ms_per_step": round(step_s * 1e3, 3), "timing": "host wall clock around zlb_decoder_begin -> zlb_decode_blocks (host buffers) -> zlb_decoder_end"},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "zl_rolz_decode_kernel", "achieved": round(ach, 4), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 7),
                         "traffic": None, "algorithmic_bytes_per_launch": int(nbytes + z.size), "peak_source": peak_src,
                         "note": "algorithmic bytes = compressed bytes read + decoded bytes written (r + 1 B per unit); the kernel is one serial chain per stream"},
            "cpu_baseline": {"value": round(sample / 1e6 / cpu_dt, 3), "unit": "MB/s", "cores": 1, "kind": cpu_kind,
                             "sample": "decode of the first %d B of the stream once, single thread; %d host cores present" % (sample, os.cpu_count())},
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_batch(args, rank, world, local, dist, torch, libzling_b200):
    """--streams S: S independent streams per GPU through ONE zlb_encode_batch / zlb_decode_batch call (host buffers in and out);
    every stream is checked against the CPU reference before anything is timed"""
    nbytes = int(args.size_mb * 1e6)
    S = args.streams
    rng_seed = 1000 * rank
    import libzling_b200.corpus as corpus
    streams = [corpus.enwik8_shaped(nbytes, seed=rng_seed + 8 + i) if args.corpus == "enwik8" else corpus.mixed(nbytes, seed=rng_seed + 4 + i) for i in range(S)]
    lib, cpu_kind = cpu_lib()
    t0 = time.perf_counter()
    want = [lib.encode(x, args.level) for x in streams]
    cpu_enc = time.perf_counter() - t0
    nblocks = S * ((nbytes + BLOCK - 1) // BLOCK)
    ctx = libzling_b200.Context(device=local, max_blocks=max(nblocks, S))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    if args.mode == "encode":
        def step():
            t0 = time.perf_counter()
            out = ctx.encode_batch(streams, args.level)
            return out, time.perf_counter() - t0
        ok = step()[0] == want
    else:
        def step():
            t0 = time.perf_counter()
            out = ctx.decode_batch(want, [nbytes] * S)
            return out, time.perf_counter() - t0
        ok = step()[0] == [x.tobytes() for x in streams]
    if not ok:
        raise SystemExit("bench.py: batch output differs from the CPU reference — refusing to report a number")
    cpu_dt = cpu_enc
    if args.mode == "decode":
        t0 = time.perf_counter()
        for zz in want:
            lib.decode(zz, nbytes)
        cpu_dt = time.perf_counter() - t0
    for _ in range(max(0, args.warmup - 1)):
        step()
    sampler = ClockSampler(local)
    sampler.start()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    times, launches = [], 0
    barrier()
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        times.append(step()[1])
        launches += ctx.stats()["launches"]
    barrier()
    clocks = sampler.stop()
    step_s = float(np.mean(times))
    if dist is not None:
        t = torch.tensor([step_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_s = float(t[0])
    value = world * S * nbytes / 1e6 / step_s
    zbytes = sum(len(z) for z in want)
    if rank == 0:
        line = {
            "metric": METRIC if args.mode == "encode" else METRIC_DEC, "value": round(value, 3), "unit": "MB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(step_s * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "%d independent %s streams of %d B per GPU, level e%d, %s, ONE batch call (zlb_%s_batch), host buffers" % (
                           S, "enwik8-shaped" if args.corpus == "enwik8" else "mixed", nbytes, args.level, args.mode, args.mode),
                       "streams_per_gpu": S, "bytes_per_gpu": int(S * nbytes), "level": args.level, "compressed_bytes": int(zbytes), "bit_exact_vs_cpu_reference": True,
                       "l2": "256 MB buffer written between timed steps (L2 flush)", "lib_sha16": lib_sha16()},
            "e2e": {"value": round(value, 3), "unit": "MB/s", "h2d_bytes_per_step": int(S * nbytes if args.mode == "encode" else zbytes),
                    "d2h_bytes_per_step": int(zbytes if args.mode == "encode" else S * nbytes), "ms_per_step": round(step_s * 1e3, 3),
                    "timing": "host wall clock around the batch call (pageable host buffers in and out), max over ranks"},
            "gpu_launches": int(launches), "clocks": clocks,
            "cpu_baseline": {"value": round(S * nbytes / 1e6 / cpu_dt, 3), "unit": "MB/s", "cores": 1, "kind": cpu_kind,
                             "sample": "the same %d streams once, one after the other, single thread; %d host cores present" % (S, os.cpu_count())},
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
