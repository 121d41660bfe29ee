"""Seeded synthetic FASTQ generators for the parity tests (numpy; small sizes).

Shapes follow SURVEY.md 8(d): Illumina-like fixed-length reads with (i) a 4-level binned quality
regime and (ii) a 41-level regime, and 454/Ion-like variable-length reads with restricted IUPAC codes
(only N,R,W,S may survive into the DNA stream -- SURVEY 8-Q9).

The large-scale generator used by bench.py lives in the product (dsrc_b200/csrc/synth.cuh); this
one exists so tests can build awkward inputs (ragged lengths, mixed titles, CRLF, ...).
"""
import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def _markov_quals(rng, n, length, levels, stay=0.85):
    """sticky Markov chain over `levels` (array of phred values) -> (n, length) uint8 phred."""
    k = len(levels)
    state = rng.integers(0, k, size=n)
    out = np.empty((n, length), dtype=np.uint8)
    for j in range(length):
        move = rng.random(n) > stay
        step = rng.integers(-2, 3, size=n)
        state = np.where(move, np.clip(state + step, 0, k - 1), state)
        out[:, j] = levels[state]
    return out


def illumina(n_reads, length=150, seed=1234, regime="binned", n_rate=2e-3, tail_frac=0.10,
             start_index=1, crlf=False, plus_rep=False, barcode_var=False, small_field=False):
    rng = np.random.default_rng(seed)
    if regime == "binned":
        levels = np.array([2, 12, 23, 37], dtype=np.uint8)
    elif regime == "mid12":          # 12 sticky levels (+ the transferred N): 9..16 symbols, still the 16-symbol models
        levels = np.array([2, 6, 9, 12, 15, 18, 21, 24, 27, 30, 34, 38], dtype=np.uint8)
    elif regime == "two":            # <= 4 symbols
        levels = np.array([11, 37], dtype=np.uint8)
    else:
        levels = np.arange(0, 41, dtype=np.uint8)
    q = _markov_quals(rng, n_reads, length, levels)
    seq = BASES[rng.integers(0, 4, size=(n_reads, length))]
    # '#' tails
    tails = rng.random(n_reads) < tail_frac
    tl = rng.integers(1, 41, size=n_reads)
    for i in np.nonzero(tails)[0]:
        q[i, length - tl[i]:] = 2
    # N with quality '#'
    nmask = rng.random((n_reads, length)) < n_rate
    seq = np.where(nmask, ord("N"), seq).astype(np.uint8)
    q = np.where(nmask, 2, q).astype(np.uint8)
    qual = (q + 33).astype(np.uint8)
    lines = []
    eol = b"\r\n" if crlf else b"\n"
    x = rng.integers(1000, 32000, size=n_reads)
    yinc = rng.integers(0, 7, size=n_reads)
    y = 1000 + np.cumsum(yinc)
    bcs = [b"ACGTACGT", b"ACGTACGA", b"TTGTACGT", b"ACGAACGT"]
    for i in range(n_reads):
        idx = start_index + i
        lane = 1 + (idx >> 12) % 4
        tile = 1101 + (idx >> 8) % 96
        bc = bcs[int(x[i]) % 4] if barcode_var else bcs[0]
        title = b"@SIM.%d A00123:45:HXXXXXXX:%d:%d:%d:%d 1:N:0:%s" % (idx, lane, tile, x[i], y[i], bc)
        if small_field:  # few-valued, non-run-length field early in the title -> ValueVar + Huffman, hit by SURVEY 8-Q1
            title = b"@S.%d." % (int(x[i]) % 37) + title[1:] + b" %d" % (int(x[i]) % 11)
        lines.append(title + eol + seq[i].tobytes() + eol + (b"+" + title[1:] if plus_rep else b"+") + eol
                     + qual[i].tobytes() + eol)
    return b"".join(lines)


def ion454(n_reads, seed=5, iupac=True):
    """variable-length reads, ~45 distinct qualities decreasing along the read, titles with
    key=value fields; ambiguity codes N (mostly q<7), R, W, S (q>=7)."""
    rng = np.random.default_rng(seed)
    lines = []
    for i in range(n_reads):
        L = int(np.clip(rng.normal(350, 90), 40, 600))
        s = BASES[rng.integers(0, 4, size=L)].copy()
        # homopolymer bias
        rep = rng.random(L) < 0.3
        for j in range(1, L):
            if rep[j]:
                s[j] = s[j - 1]
        base = np.linspace(40, 8, L) + rng.normal(0, 4, L)
        q = np.clip(base, 0, 44).astype(np.uint8)
        if iupac:
            amb = rng.random(L) < 0.004
            for j in np.nonzero(amb)[0]:
                c = rng.choice(list(b"NNNRWS"))
                s[j] = c
                if c == ord("N") and rng.random() < 0.8:
                    q[j] = rng.integers(0, 7)
                else:
                    q[j] = max(int(q[j]), 7)
        name = "".join(rng.choice(list("ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789"), size=14))
        title = ("@%s rank=%07d x=%d y=%d length=%d" % (name, i + 1, rng.integers(1, 4096), rng.integers(1, 4096), L)).encode()
        lines.append(title + b"\n" + s.tobytes() + b"\n+\n" + (q + 33).astype(np.uint8).tobytes() + b"\n")
    return b"".join(lines)


def mixed_titles(n_reads, length=50, seed=3):
    """titles whose field structure changes mid-block -> TagRawEncoder path (FLAG_MIXED_FIELD_FORMATTING)."""
    rng = np.random.default_rng(seed)
    lines = []
    for i in range(n_reads):
        s = BASES[rng.integers(0, 4, size=length)]
        q = rng.integers(2, 41, size=length).astype(np.uint8) + 33
        if i % 7 == 3:
            title = b"@lonely%d" % i
        else:
            title = b"@read_%d/1" % i
        lines.append(title + b"\n" + s.tobytes() + b"\n+\n" + q.tobytes() + b"\n")
    return b"".join(lines)


def skewed(n_reads, length=150, seed=8, p_major=0.97, levels=(2, 12, 23, 37)):
    """almost-constant bases and qualities: a handful of contexts hold tens of thousands of symbols each, so the
    adaptive rows hit TSymbolCoderRC::Rescale (src/SymbolCoderRC.h:69-73) several times inside one block."""
    rng = np.random.default_rng(seed)
    seq = np.where(rng.random((n_reads, length)) < p_major, ord("A"), BASES[rng.integers(0, 4, size=(n_reads, length))]).astype(np.uint8)
    lv = np.array(levels, dtype=np.uint8) + 33
    qual = np.where(rng.random((n_reads, length)) < p_major, ord("I"), lv[rng.integers(0, len(lv), size=(n_reads, length))]).astype(np.uint8)
    return b"".join(b"@SK.%d 1:N:0\n" % (i + 1) + seq[i].tobytes() + b"\n+\n" + qual[i].tobytes() + b"\n" for i in range(n_reads))


def varlen_binned(n_reads, seed=21, levels=(2, 12, 23, 37), lo=30, hi=170):
    """variable-length reads with few sticky quality levels: the 16-symbol quality models with per-record position buckets
    (TTranslationalQualityEncoder::Encode, j * rescale / len with a different len per read) over many tiles."""
    rng = np.random.default_rng(seed)
    lv = np.array(levels, dtype=np.uint8)
    lines = []
    for i in range(n_reads):
        L = int(rng.integers(lo, hi + 1))
        s = BASES[rng.integers(0, 4, size=L)]
        q = _markov_quals(rng, 1, L, lv)[0]
        lines.append(b"@VL.%d len=%d\n" % (i + 1, L) + s.tobytes() + b"\n+\n" + (q + 33).astype(np.uint8).tobytes() + b"\n")
    return b"".join(lines)


def random_quals(n_reads, length=150, seed=9, n_levels=60):
    """uniformly random qualities over many levels: tens of thousands of distinct order-2 contexts in one block
    (exercises the 64-symbol models and, on the GPU decode path, the full-table retry of a filled context hash)."""
    rng = np.random.default_rng(seed)
    seq = BASES[rng.integers(0, 4, size=(n_reads, length))]
    qual = (rng.integers(0, n_levels, size=(n_reads, length)) + 33).astype(np.uint8)
    return b"".join(b"@RQ.%d\n" % (i + 1) + seq[i].tobytes() + b"\n+\n" + qual[i].tobytes() + b"\n" for i in range(n_reads))


def changing_titles(seed=12):
    """a file whose title structure changes between blocks: 4-field titles (the reference's field vector ends at capacity 4), then
    13-field Illumina titles (the vector grows 4 -> 8 -> 16 in the middle of a block's first title), then 454-style titles --
    SURVEY 8-Q1 in its general form (capacity carried across blocks in file order)."""
    rng = np.random.default_rng(seed)
    a = []
    for i in range(4200):
        s = BASES[rng.integers(0, 4, size=100)]
        q = (rng.integers(20, 41, size=100) + 33).astype(np.uint8)
        a.append(b"@r%d/%d x=%d\n" % (i + 1, 1 + i % 2, int(rng.integers(0, 19))) + s.tobytes() + b"\n+\n" + q.tobytes() + b"\n")
    return b"".join(a) + illumina(3200, seed=seed + 1, small_field=True) + ion454(1500, seed=seed + 2)


def exact_size(size, seed=3, crlf=False, base_reads=10000):
    """a FASTQ file of exactly `size` bytes (the last title is padded): hits the reader's EOF-on-refill branch when `size` makes a
    full chunk window end exactly at the end of the file (src/FastqStream.cpp:41-69: Read() returns 0, the chunk is the bare carry-over)"""
    big = illumina(base_reads, seed=seed, crlf=crlf)
    eol = b"\r\n" if crlf else b"\n"
    recs = big.split(eol)
    out, tot, i = [], 0, 0
    while True:
        rec = eol.join(recs[i:i + 4]) + eol
        assert i + 4 < len(recs), "base too small"
        if tot + len(rec) + 400 > size:
            break
        out.append(rec)
        tot += len(rec)
        i += 4
    ln = 50
    tl = size - tot - (len(eol) * 4 + ln * 2 + 1)
    out.append(b"@" + b"x" * (tl - 1) + eol + b"A" * ln + eol + b"+" + eol + b"I" * ln + eol)
    out = b"".join(out)
    assert len(out) == size
    return out


IUPAC_ALL = b"NRWSKMDVHBYXU.-"      # indices 4..18 of the reference's symbol table (src/RecordsProcessor.cpp:180-207)


def block_scale(n_reads=1000, length=150, n_levels=41, sticky=False, seed=1, low_amb=0.0, low_codes=b"N", high_amb=0.0,
                high_codes=b"N", tail_frac=0.0, tail_max=60, varlen=False, q_lo=0):
    """Vectorised block-scale generator (>= 256 KiB for 1000 x 150 bp) for the model shapes the small catalogue does not reach:
    n_levels quality values (iid, or a sticky chain), ambiguity codes with q < 7 (`low_*`: transferred into the quality byte, any
    IUPAC code -- SURVEY a4) and with q >= 7 (`high_*`: they stay in the DNA stream, so only N/R/W/S -- SURVEY 8-Q9), '#' tails,
    optional per-read lengths. Quality values start at q_lo (>= 7 keeps `low` codes the only symbols below 7)."""
    rng = np.random.default_rng(seed)
    lv = (q_lo + np.arange(n_levels)).astype(np.uint8)
    if sticky:
        q = _markov_quals(rng, n_reads, length, lv, stay=0.9)
    else:
        q = lv[rng.integers(0, n_levels, size=(n_reads, length))]
    seq = BASES[rng.integers(0, 4, size=(n_reads, length))].copy()
    lens = rng.integers(max(20, length // 3), length + 1, size=n_reads) if varlen else np.full(n_reads, length)
    if tail_frac:
        t = np.where(rng.random(n_reads) < tail_frac, rng.integers(1, tail_max + 1, size=n_reads), 0)
        q = np.where(np.arange(length)[None, :] >= (lens - np.minimum(t, lens - 1))[:, None], 2, q).astype(np.uint8)
    if high_amb:
        m = rng.random((n_reads, length)) < high_amb
        codes = np.frombuffer(high_codes, dtype=np.uint8)[rng.integers(0, len(high_codes), size=(n_reads, length))]
        seq = np.where(m, codes, seq)
        q = np.where(m, np.maximum(q, 7), q).astype(np.uint8)
    if low_amb:
        m = rng.random((n_reads, length)) < low_amb
        codes = np.frombuffer(low_codes, dtype=np.uint8)[rng.integers(0, len(low_codes), size=(n_reads, length))]
        seq = np.where(m, codes, seq)
        q = np.where(m, rng.integers(0, 7, size=(n_reads, length)), q).astype(np.uint8)
    qual = (q + 33).astype(np.uint8)
    x = rng.integers(1000, 30000, size=n_reads)
    out = []
    for i in range(n_reads):
        ln = int(lens[i])
        out.append(b"@B.%d %d:%d\n" % (i + 1, x[i], 7 * i) + seq[i, :ln].tobytes() + b"\n+\n" + qual[i, :ln].tobytes() + b"\n")
    return b"".join(out)


def read_id_styles(style, n_reads=600, seed=51):
    """read IDs as other platforms / archives write them (the tag tokenizer's envelope): `sra` = SRA dumps
    ("@SRR1770413.17 HWI-ST1106:...:1101:1249:2062 length=100"), `ont` = Nanopore (UUID, runid hash, key=value pairs, ISO time stamp:
    ~40 separator-delimited fields, 240-byte titles), `pacbio` = "@m64011_190830_220126/4194376/ccs", `bgi` = "@V350012345L1C001R0010000001/1"."""
    rng = np.random.default_rng(seed)
    hexd = np.frombuffer(b"0123456789abcdef", dtype=np.uint8)
    lines = []
    for i in range(n_reads):
        if style == "sra":
            L = 100
            title = b"@SRR1770413.%d HWI-ST1106:418:D1WJWACXX:3:1101:%d:%d length=%d" % (i + 1, 1200 + int(rng.integers(0, 20000)), 2000 + 5 * i, L)
        elif style == "ont":
            L = int(rng.integers(200, 900))
            u = hexd[rng.integers(0, 16, size=32)].tobytes()
            title = (b"@%s-%s-%s-%s-%s runid=5d1f6bd2c0b1ee3ad2b5fa2ac6bdd7a2a1a0f0c3 sampleid=S1 read=%d ch=%d start_time=2021-03-%02dT%02d:%02d:%02dZ "
                     b"flow_cell_id=FAO12345 protocol_group_id=run_7 barcode=barcode%02d"
                     % (u[:8], u[8:12], u[12:16], u[16:20], u[20:32], 100 + 7 * i, int(rng.integers(1, 513)), 1 + i % 28, int(rng.integers(0, 24)),
                        int(rng.integers(0, 60)), int(rng.integers(0, 60)), int(rng.integers(1, 13))))
        elif style == "pacbio":
            L = int(rng.integers(300, 1200))
            title = b"@m64011_190830_220126/%d/ccs" % (4194376 + 131 * i + int(rng.integers(0, 100)))
        else:
            L = 100
            title = b"@V350012345L%dC%03dR%03d%07d/1" % (1 + i % 4, 1 + (i // 40) % 8, 1 + (i // 7) % 60, i * 3 + int(rng.integers(0, 3)))
        sq = BASES[rng.integers(0, 4, size=L)]
        q = _markov_quals(rng, 1, L, np.array([2, 12, 23, 37], dtype=np.uint8))[0]
        lines.append(title + b"\n" + sq.tobytes() + b"\n+\n" + (q + 33).astype(np.uint8).tobytes() + b"\n")
    return b"".join(lines)
