"""Seeded synthetic inputs for BASELINE.json's configs.  numpy only, deterministic per seed.

  ascii_words(n)      config 1: 3000-word lowercase vocabulary, uniform pick, single spaces.
  enwik8_shaped(n)    configs 2/3: Zipf vocabulary (>=100k words), recurring phrases, sentence/paragraph
                      structure, wiki/XML markup, digits/dates, ~3 % multi-byte UTF-8.  Acceptance band
                      (checked in tests/test_corpus.py): order-0 entropy 4.9-5.3 bit/B, zlib-6 ratio 0.33-0.40,
                      zling e0 ratio 0.28-0.35.
  mixed(n)            config 4: seeded segments of 1-8 MB: ~60 % text, ~25 % structured binary, ~15 % random.

Real enwik8 is not available offline; these are stand-ins of the same shape, and bench.py says "synthetic".
"""
import numpy as np

_LETTERS = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
_LETTER_P = np.array([12.7, 9.1, 8.2, 7.5, 7.0, 6.7, 6.3, 6.1, 6.0, 4.3, 4.0, 2.8, 2.8, 2.4, 2.4, 2.2, 2.0, 2.0,
                      1.9, 1.5, 1.0, 0.8, 0.15, 0.15, 0.1, 0.07])
_LETTER_P = _LETTER_P / _LETTER_P.sum()


def _vocab(rng, nwords, minlen, maxlen, mean=None):
    """returns (blob u8, starts i64, lens i64) of `nwords` distinct-ish lowercase words"""
    if mean is None:
        lens = rng.integers(minlen, maxlen + 1, size=nwords)
    else:
        lens = np.clip(rng.poisson(mean - minlen, size=nwords) + minlen, minlen, maxlen)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    blob = _LETTERS[rng.choice(26, size=int(lens.sum()), p=_LETTER_P)]
    return blob, starts.astype(np.int64), lens.astype(np.int64)


def _gather(blob, starts, lens, ids):
    """concatenate blob[starts[i]:starts[i]+lens[i]] for i in ids"""
    l = lens[ids]
    total = int(l.sum())
    if total == 0:
        return np.zeros(0, dtype=np.uint8)
    out_start = np.concatenate([[0], np.cumsum(l)[:-1]])
    src = np.repeat(starts[ids] - out_start, l) + np.arange(total, dtype=np.int64)
    return blob[src]


def ascii_words(n, seed=1):
    """config 1: 3000 lowercase words of length 2-9, uniform, single-space separated"""
    rng = np.random.default_rng(seed)
    blob, starts, lens = _vocab(rng, 3000, 2, 9)
    # append the separator to every word once, then one gather
    wl = lens + 1
    ws = np.concatenate([[0], np.cumsum(wl)[:-1]])
    wblob = np.full(int(wl.sum()), 0x20, dtype=np.uint8)
    dst = np.repeat(ws - starts, lens) + np.arange(blob.size)
    wblob[dst] = blob
    out = np.zeros(0, dtype=np.uint8)
    parts = []
    have = 0
    while have < n:
        ids = rng.integers(0, 3000, size=200000)
        p = _gather(wblob, ws, wl, ids)
        parts.append(p)
        have += p.size
    out = np.concatenate(parts)[:n]
    return np.ascontiguousarray(out)


class _Enwik:
    """Vectorised article-text generator.  A table of byte-string UNITS (plain / Capitalised words, recurring
    phrases, numbers+dates, separators with wiki markup) is built once; a chunk is a sequence of unit ids
    (word, separator, word, separator, ...) expanded with one gather.  Page headers are rendered per chunk."""

    _SEPS = [b" ", b", ", b". ", b".\n\n", b" [[", b"]] ", b"]], ", b"]]. ", b" ''", b"'' ", b" &quot;", b"&quot; ",
             b".\n\n* ", b" (", b") ", b"; ", b": ", b"|", b" &amp; ", b" - ", b".\n*", b" '''", b"''' ", b"]]\n[[",
             b"&lt;br&gt; ", b" = ", b" {{", b"}} ", b".&lt;ref&gt;", b"&lt;/ref&gt; "]
    (S_SP, S_COMMA, S_DOT, S_PARA, S_LO, S_LC, S_LCC, S_LCD, S_IO, S_IC, S_QO, S_QC, S_BUL, S_PO, S_PC, S_SEMI, S_COL,
     S_BAR, S_AMP, S_DASH, S_BUL2, S_BO, S_BC, S_CAT, S_BR, S_EQ, S_TO, S_TC, S_RO, S_RC) = range(30)

    def __init__(self, seed):
        rng = np.random.default_rng(seed)
        self.rng = rng
        nv = self.nv = 131072
        # syllable-built vocabulary so that words share sub-word structure like natural language
        nsyl = 900
        cons = np.frombuffer(b"tnshrdlcmwfgypbvkjxqz", dtype=np.uint8)
        cons_p = np.array([9.1, 6.7, 6.3, 6.1, 6.0, 4.3, 4.0, 2.8, 2.4, 2.4, 2.2, 2.0, 2.0, 1.9, 1.5, 1.0, 0.8, 0.15,
                           0.15, 0.1, 0.07])
        cons_p /= cons_p.sum()
        vow = np.frombuffer(b"eaoiuy", dtype=np.uint8)
        vow_p = np.array([12.7, 8.2, 7.5, 7.0, 2.8, 0.6])
        vow_p /= vow_p.sum()
        syl = []
        for _ in range(nsyl):
            t = rng.random()
            on = bytes(cons[rng.choice(cons.size, size=int(rng.integers(0, 3)), p=cons_p)])
            vv = bytes(vow[rng.choice(vow.size, size=1 if t < 0.8 else 2, p=vow_p)])
            cd = bytes(cons[rng.choice(cons.size, size=int(rng.integers(0, 2)), p=cons_p)])
            syl.append(on + vv + cd)
        sz = 1.0 / np.arange(1, nsyl + 1) ** 0.8
        sz /= sz.sum()
        nsy = np.clip(rng.poisson(1.45, size=nv) + 1, 1, 5)
        sid = rng.choice(nsyl, size=int(nsy.sum()), p=sz)
        words, at = [], 0
        for k in nsy:
            words.append(b"".join(syl[j] for j in sid[at:at + k]))
            at += k
        head = [b"the", b"of", b"and", b"in", b"a", b"to", b"is", b"was", b"for", b"as", b"by", b"with", b"that",
                b"on", b"from", b"his", b"it", b"an", b"are", b"at", b"or", b"he", b"which", b"be", b"this", b"also",
                b"were", b"has", b"had", b"not", b"its", b"but", b"one", b"their", b"have", b"first", b"other", b"new"]
        words[:len(head)] = head
        for i in rng.choice(np.arange(1500, nv), size=34000, replace=False):   # ~3 % multi-byte UTF-8 overall
            w = bytearray(words[i])
            k = int(rng.integers(0, len(w) + 1))
            if rng.random() < 0.6:
                w[k:k] = bytes([0xC3, int(rng.integers(0xA0, 0xBF))])
            else:
                w[k:k] = bytes([int(rng.integers(0xE3, 0xE9)), int(rng.integers(0x80, 0xBF)),
                                int(rng.integers(0x80, 0xBF))]) * int(rng.integers(1, 4))
            words[i] = bytes(w)
        for i in rng.choice(np.arange(200, nv), size=42000, replace=False):    # proper nouns
            words[i] = words[i][:1].upper() + words[i][1:]
        for i in rng.choice(np.arange(3000, nv), size=6000, replace=False):    # acronyms
            words[i] = words[i][:4].upper()
        self.words = words
        zipf = 1.0 / np.arange(1, nv + 1) ** 1.05
        self.word_cdf = np.cumsum(zipf / zipf.sum())
        nph = self.nphrase = 30000
        wid = np.searchsorted(self.word_cdf, rng.random(nph * 5))
        klen = rng.integers(2, 6, size=nph)
        phrases, at = [], 0
        for k in klen:
            phrases.append(b" ".join(words[j] for j in wid[at:at + k]))
            at += 5
        pz = 1.0 / np.arange(1, nph + 1) ** 0.85
        self.phrase_cdf = np.cumsum(pz / pz.sum())
        self.phrases = phrases
        months = [b"January", b"February", b"March", b"April", b"May", b"June", b"July", b"August", b"September",
                  b"October", b"November", b"December"]
        nums = []
        for i in range(20000):
            r = i % 8
            v = int(rng.integers(0, 100000))
            if r < 3:
                nums.append(b"%d" % (1000 + v % 1025))
            elif r == 3:
                nums.append(b"%d %s %d" % (1 + v % 28, months[v % 12], 1500 + v % 510))
            elif r == 4:
                nums.append(b"%d" % v)
            elif r == 5:
                nums.append(b"%d.%d" % (v % 100, v % 10))
            elif r == 6:
                nums.append(b"%d,%03d" % (1 + v % 900, v % 1000))
            else:
                nums.append(b"%d%%" % (v % 100))
        self.nnum = len(nums)
        # unit table: [plain words | Capitalised words | phrases | Capitalised phrases | numbers | separators]
        units = (words + [w[:1].upper() + w[1:] for w in words] + phrases + [p[:1].upper() + p[1:] for p in phrases]
                 + nums + self._SEPS)
        self.off_cap, self.off_ph, self.off_phcap = nv, 2 * nv, 2 * nv + nph
        self.off_num = 2 * nv + 2 * nph
        self.off_sep = self.off_num + self.nnum
        lens = np.array([len(u) for u in units], dtype=np.int64)
        self.u_len = lens
        self.u_start = np.concatenate([[0], np.cumsum(lens)[:-1]])
        self.u_blob = np.frombuffer(b"".join(units), dtype=np.uint8)
        self.pageno = 0

    def chunk(self, nbytes):
        rng = self.rng
        n = max(1000, int(nbytes / 6.9))
        kind = rng.random(n)
        tok = np.searchsorted(self.word_cdf, rng.random(n)).astype(np.int64)
        isph = kind < 0.16
        isnum = (kind >= 0.16) & (kind < 0.20)
        tok[isph] = self.off_ph + np.searchsorted(self.phrase_cdf, rng.random(int(isph.sum())))
        tok[isnum] = self.off_num + rng.integers(0, self.nnum, size=int(isnum.sum()))
        # topical locality: tokens (words, phrases, numbers) recur within an article: ~28 % of the tokens repeat a
        # token seen up to ~1200 tokens earlier, a third of those together with their successor
        back = rng.integers(1, 1200, size=n)
        rep = np.nonzero(rng.random(n) < 0.20)[0]
        rep = rep[(rep >= 1200) & (rep < n - 1)]
        tok[rep] = tok[rep - back[rep]]
        rep2 = rep[::3]
        tok[rep2 + 1] = tok[rep2 + 1 - back[rep2]]
        isph = (tok >= self.off_ph) & (tok < self.off_phcap)
        isnum = tok >= self.off_num
        sep = np.full(n, self.S_SP, dtype=np.int64)
        r = rng.random(n)
        sep[r < 0.065] = self.S_COMMA
        sep[(r >= 0.065) & (r < 0.072)] = self.S_SEMI
        sep[(r >= 0.072) & (r < 0.078)] = self.S_COL
        sep[(r >= 0.078) & (r < 0.083)] = self.S_DASH
        sep[(r >= 0.083) & (r < 0.086)] = self.S_AMP
        # markup wrappers around phrase tokens: links, piped links, italics, quotes, parentheses, bold, templates
        wrap = rng.integers(0, 20, size=n)
        idx = np.nonzero(isph[1:-1])[0] + 1
        w = wrap[idx]
        for lo, hi, so, sc in ((0, 8, self.S_LO, self.S_LC), (8, 9, self.S_LO, self.S_LCC), (9, 11, self.S_IO, self.S_IC),
                               (11, 12, self.S_QO, self.S_QC), (12, 13, self.S_PO, self.S_PC),
                               (13, 14, self.S_BO, self.S_BC), (14, 15, self.S_TO, self.S_TC)):
            j = idx[(w >= lo) & (w < hi)]
            sep[j - 1] = so
            sep[j] = sc
        j = idx[w == 15]                       # piped link
        sep[j - 1] = self.S_LO
        sep[j] = self.S_BAR
        sep[np.minimum(j + 1, n - 1)] = self.S_LC
        j = np.nonzero(isnum[:-1] & (wrap[:-1] < 3))[0]
        sep[j] = self.S_RO
        sep[np.minimum(j + 2, n - 1)] = self.S_RC
        # sentences and paragraphs
        ends = np.cumsum(rng.geometric(1.0 / 17.0, size=n // 8 + 8) + 3)
        ends = ends[ends < n - 1]
        pk = rng.random(ends.size)
        plain = (sep[ends] == self.S_SP) | (sep[ends] == self.S_COMMA)
        e = ends[plain]
        pk = pk[plain]
        sep[e] = self.S_DOT
        sep[e[pk < 0.16]] = self.S_PARA
        sep[e[(pk >= 0.16) & (pk < 0.19)]] = self.S_BUL
        sep[e[(pk >= 0.19) & (pk < 0.20)]] = self.S_BR
        nxt = e + 1
        capw = nxt[tok[nxt] < self.nv]
        tok[capw] += self.off_cap
        capp = nxt[(tok[nxt] >= self.off_ph) & (tok[nxt] < self.off_phcap)]
        tok[capp] += self.nphrase
        ids = np.empty(2 * n, dtype=np.int64)
        ids[0::2] = tok
        ids[1::2] = self.off_sep + sep
        body = _gather(self.u_blob, self.u_start, self.u_len, ids)
        # cut into pages (~1.5-9 KB each) with rendered XML heads
        cuts = np.cumsum(rng.integers(1500, 9000, size=body.size // 3000 + 2))
        cuts = cuts[cuts < body.size]
        parts, prev = [], 0
        for c in cuts:
            parts.append(body[prev:c])
            prev = c
            parts.append(np.frombuffer(self._page_head(), dtype=np.uint8))
        parts.append(body[prev:])
        return np.concatenate(parts)

    def _page_head(self):
        rng = self.rng
        self.pageno += 1
        n = self.pageno * 7 + int(rng.integers(0, 7))
        title = self.phrases[int(np.searchsorted(self.phrase_cdf, rng.random()))].title()
        user = self.words[300 + int(rng.integers(0, 3000))].capitalize()
        v = int(rng.integers(0, 1 << 30))
        if v % 5 == 0:
            who = b"        <ip>%d.%d.%d.%d</ip>\n" % (v & 255, (v >> 8) & 255, (v >> 16) & 255, (v >> 24) & 63)
        else:
            who = b"        <username>%s</username>\n        <id>%d</id>\n" % (user, v % 99991)
        cat = b"".join(b"[[Category:%s]]\n" % self.phrases[int(np.searchsorted(self.phrase_cdf, rng.random()))].title()
                       for _ in range(v % 4))
        return (b"\n\n" + cat + b"</text>\n    </revision>\n  </page>\n  <page>\n    <title>" + title
                + b"</title>\n    <id>%d</id>\n    <revision>\n      <id>%d</id>\n"
                  b"      <timestamp>200%d-%02d-%02dT%02d:%02d:%02dZ</timestamp>\n      <contributor>\n%s"
                  b"      </contributor>\n      <text xml:space=\"preserve\">"
                % (n, n * 31 + v % 31, 2 + v % 5, 1 + v % 12, 1 + v % 28, (v >> 5) % 24, (v >> 10) % 60, (v >> 16) % 60, who)
                + (b"'''" + title + b"''' " if v % 3 else b""))


def enwik8_shaped(n, seed=8):
    g = _Enwik(seed)
    parts = []
    have = 0
    while have < n:
        c = g.chunk(2 << 20)             # fixed request size: output is prefix-stable in n
        parts.append(c)
        have += c.size
    return np.ascontiguousarray(np.concatenate(parts)[:n])


def _binary_segment(rng, n):
    kind = int(rng.integers(0, 3))
    if kind == 0:      # little-endian int32 counters with small strides
        k = n // 4 + 1
        base = int(rng.integers(0, 1 << 20))
        v = (base + np.cumsum(rng.integers(0, 4, size=k))).astype("<i4")
        return v.view(np.uint8)[:n]
    if kind == 1:      # float32 ramp with noise
        k = n // 4 + 1
        v = (np.linspace(0, 1000, k) + rng.normal(0, 0.01, size=k)).astype("<f4")
        return v.view(np.uint8)[:n]
    rec = int(rng.integers(32, 65))   # repeated fixed-size records with a few mutating fields
    k = n // rec + 1
    proto = rng.integers(0, 256, size=rec, dtype=np.uint8)
    a = np.tile(proto, (k, 1))
    a[:, 0:4] = np.arange(k, dtype="<u4").view(np.uint8).reshape(k, 4)
    a[:, rec // 2] = rng.integers(0, 8, size=k, dtype=np.uint8)
    return a.reshape(-1)[:n]


def mixed(n, seed=4):
    """config 4: text / structured binary / random segments of 1-8 MB, not block aligned"""
    rng = np.random.default_rng(seed)
    g = _Enwik(seed + 100)
    parts = []
    have = 0
    while have < n:
        seg = int(rng.integers(1 << 20, 8 << 20))
        seg = min(seg, n - have)
        r = rng.random()
        if r < 0.60:
            sub, got = [], 0
            while got < seg:
                c = g.chunk(2 << 20)
                sub.append(c)
                got += c.size
            p = np.concatenate(sub)[:seg]
        elif r < 0.85:
            p = _binary_segment(rng, seg)
        else:
            p = rng.integers(0, 256, size=seg, dtype=np.uint8)
        parts.append(np.ascontiguousarray(p, dtype=np.uint8))
        have += seg
    return np.ascontiguousarray(np.concatenate(parts)[:n])


def order0_entropy(a):
    c = np.bincount(np.asarray(a, dtype=np.uint8), minlength=256).astype(np.float64)
    p = c[c > 0] / c.sum()
    return float(-(p * np.log2(p)).sum())
