#!/usr/bin/env python
"""Analysis aid for DESIGN.md §10 item 0 (segment-parallel resolution): how fast does the speculative token chain
(frozen decisions, "no word hit") entered at an arbitrary position merge with the true chain, and how many segments
of a window could a speculative walker finish without deviating from the truth?

    ZL_V3_DUMP=/tmp/x.dump <parse_v3_sim> file level      # host replay writes frozen lengths + final marks
    python scripts/chain_sync_stats.py /tmp/x.dump
"""
import sys

import numpy as np

path = sys.argv[1]
raw = np.fromfile(path, dtype=np.uint8)
n = raw.size // 3
flen = raw[: 2 * n].view(np.uint16).astype(np.int64)
mark = raw[2 * n: 3 * n]
kind = mark & 7
explicit = (mark & 8) != 0
starts = np.flatnonzero(kind)
print("positions %d, tokens %d, explicit (non-frozen) %.2f%%, word hits %.2f%%" % (n, starts.size, 100 * explicit[starts].mean(), 100 * np.isin(kind[starts], (3, 4)).mean()))
is_start = kind != 0
nxt = np.arange(n) + np.where(flen > 0, flen, 1)        # speculative successor of every position

rng = np.random.default_rng(1)
for margin in (8, 16, 32, 64):
    merged_by, trials = [], 0
    for p in rng.integers(1000, n - 2000, 4000):
        x, steps = int(p), 0
        while x < n and not is_start[x] and x - p < 400:
            x = int(nxt[x]); steps += 1
        merged_by.append(x - p if x < n and is_start[x] else 10 ** 6)
        trials += 1
    merged_by = np.array(merged_by)
    print("entry at a random position: merged with the true chain within %3d positions in %.1f%% of trials (median %d, p90 %d)"
          % (margin, 100 * (merged_by <= margin).mean(), np.median(merged_by), np.percentile(merged_by, 90)))

# a segment is "clean" if every true token starting inside it took its frozen decision and is not a word hit
bad = np.zeros(n, dtype=bool)
bad[starts] = explicit[starts] | np.isin(kind[starts], (3, 4))
for seg in (32, 64, 128, 254):
    m = (n // seg) * seg
    per = bad[:m].reshape(-1, seg).any(axis=1)
    toks = is_start[:m].reshape(-1, seg).sum(axis=1)
    print("segments of %3d positions (%.1f tokens): %.1f%% contain no deviation from the speculation" % (seg, toks.mean(), 100 * (1 - per.mean())))
