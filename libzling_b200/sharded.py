"""One stream over several GPUs: contiguous ranges of 16 MiB blocks per rank (one process per GPU).

What couples the blocks of a stream in the reference is only what `baidu::zling::Encode` keeps outside its block
loop: the MTF tables (m_mtf[256], src/libzling_lz.h:105 — never reset) and `current_level`
(src/libzling.cpp:185,261-266).  Everything else of a block — above all the ROLZ parse, the dominant cost — is
independent.  So every rank

  1. submits its range (H2D + parse launch; returns at once),
  2. receives the 65 540-byte carried state of the previous range from rank-1 (the one real exchange step of the
     path: a point-to-point send/recv in block order) and installs it,
  3. completes its range (MTF ranks, Huffman, framing; re-parses its first block only if the carried level differs
     from the requested one), and forwards its own final state to rank+1,
  4. takes part in ONE gather of the framed outputs to rank 0, which concatenates them in block order.

`enc` is anything with submit / set_state / complete / get_state (libzling_b200.Encoder on a GPU; the tests drive the
same code with a CPU stand-in over gloo).
"""
import numpy as np

BLOCK = 16777216
STATE_BYTES = 65540


def block_ranges(nbytes, world):
    """contiguous block ranges: rank r gets bytes [lo, hi) — whole blocks, earlier ranks get the extra ones"""
    nblocks = (nbytes + BLOCK - 1) // BLOCK
    base, extra = divmod(nblocks, world)
    out, b = [], 0
    for r in range(world):
        nb = base + (1 if r < extra else 0)
        out.append((min(b * BLOCK, nbytes), min((b + nb) * BLOCK, nbytes)))
        b += nb
    return out


def encode_range(enc, shard, rank, world, dist, device="cpu"):
    """steps 1-3 for this rank; returns the framed bytes of its range (b"" for an empty range)"""
    import torch
    shard = np.ascontiguousarray(shard, dtype=np.uint8)
    if shard.size:
        enc.submit(shard)
    if rank > 0:
        t = torch.empty(STATE_BYTES, dtype=torch.uint8, device=device)
        dist.recv(t, src=rank - 1)
        enc.set_state(t.cpu().numpy())
    out = enc.complete() if shard.size else b""
    if rank + 1 < world:
        t = torch.from_numpy(np.ascontiguousarray(enc.get_state())).to(device)
        dist.send(t, dst=rank + 1)
    return out


def gather_framed(out, rank, world, dist, device="cpu"):
    """step 4: sizes by all_gather, payloads by one gather of padded buffers; rank 0 returns the whole stream"""
    import torch
    n = torch.tensor([len(out)], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    mine = torch.zeros(mx, dtype=torch.uint8, device=device)
    if len(out):
        mine[:len(out)] = torch.frombuffer(bytearray(out), dtype=torch.uint8).to(device)
    bufs = [torch.empty(mx, dtype=torch.uint8, device=device) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, bufs, dst=0)
    if rank != 0:
        return None
    return b"".join(bytes(b[:s].cpu().numpy()) for b, s in zip(bufs, sizes))


def encode_stream(enc, data, rank, world, dist, device="cpu"):
    """encode `data` (the WHOLE stream, identical on every rank; each rank only touches its range)"""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    lo, hi = block_ranges(data.size, world)[rank]
    out = encode_range(enc, data[lo:hi], rank, world, dist, device)
    return gather_framed(out, rank, world, dist, device)
