"""Shared helpers for the tests: FASTQ text, seeded random sequences, adversarial read sets.

Mirrors the helpers of the reference's tests/testthat/setup.R (GENERATE_RANDOM_SEQ,
ADD_FLANKS, CHOOSE_STRAND_FUN) with a numpy generator, since R's RNG stream is not
reproducible here.
"""
import numpy as np

BASES = "ACGT"
_COMP = str.maketrans("ACGTNacgtnRYSWKMBDHVryswkmbdhv", "TGCANTGCANYRSWMKVHDBYRSWMKVHDB")


def revcomp(s):
    return s.translate(_COMP)[::-1]


def random_seq(rng, n):
    return "".join(BASES[i] for i in rng.integers(0, 4, size=n))


def fastq(seqs, names=None, crlf=False):
    nl = "\r\n" if crlf else "\n"
    out = []
    for i, s in enumerate(seqs):
        name = names[i] if names is not None else "r%d" % (i + 1)
        out.append("@%s%s%s%s+%s%s%s" % (name, nl, s, nl, nl, "I" * len(s), nl))
    return "".join(out).encode("latin-1")


def add_flanks(rng, seqs, nleft=50, nright=50):
    """setup.R:34-44 ADD_FLANKS: 1..nleft / 1..nright random bases either side (0 = none)."""
    out = []
    for s in seqs:
        left = random_seq(rng, int(rng.integers(1, nleft + 1))) if nleft else ""
        right = random_seq(rng, int(rng.integers(1, nright + 1))) if nright else ""
        out.append(left + s + right)
    return out


def choose_strand(rng, seqs, strand):
    """setup.R:46-59 CHOOSE_STRAND_FUN."""
    if strand == "original":
        return list(seqs)
    if strand == "reverse":
        return [revcomp(s) for s in seqs]
    n = len(seqs)
    sel = set(rng.choice(n, size=n // 2, replace=False).tolist())
    return [revcomp(s) if i in sel else s for i, s in enumerate(seqs)]


def distinct_pool(rng, n, length):
    seen = set()
    out = []
    while len(out) < n:
        s = random_seq(rng, length)
        if s not in seen:
            seen.add(s)
            out.append(s)
    return out


def dense_pool(rng, n, length, frac_neighbours=0.3):
    """Distinct barcodes where a fraction are 1-substitution variants of earlier entries
    (so ties / ambiguity / best-vs-first differences actually occur)."""
    seen = set()
    out = []
    while len(out) < n:
        if out and rng.random() < frac_neighbours:
            base = out[int(rng.integers(0, len(out)))]
            pos = int(rng.integers(0, length))
            alt = BASES[int(rng.integers(0, 4))]
            s = base[:pos] + alt + base[pos + 1:]
        else:
            s = random_seq(rng, length)
        if s not in seen:
            seen.add(s)
            out.append(s)
    return out


def mutate(rng, s, sub_rate=0.0, n_rate=0.0, lower_rate=0.0):
    if not (sub_rate or n_rate or lower_rate):
        return s
    out = list(s)
    r = rng.random(size=(3, len(out)))
    for i, c in enumerate(out):
        if r[0, i] < sub_rate:
            out[i] = BASES[int(rng.integers(0, 4))]
        if r[1, i] < n_rate:
            out[i] = "N"
        elif r[2, i] < lower_rate:
            out[i] = out[i].lower()
    return "".join(out)


def fill_template(template, *variables):
    """Replace each maximal run of '-' in `template` by the next variable sequence."""
    out = []
    i = 0
    k = 0
    while i < len(template):
        if template[i] == "-":
            j = i
            while j < len(template) and template[j] == "-":
                j += 1
            assert len(variables[k]) == j - i
            out.append(variables[k])
            k += 1
            i = j
        else:
            out.append(template[i])
            i += 1
    return "".join(out)


def adversarial_reads(rng, n, template, pools, strand="both", read_len=None,
                      sub_rate=0.03, n_rate=0.01, lower_rate=0.02, junk_frac=0.1,
                      double_frac=0.1, short_frac=0.02, edge_frac=0.1):
    """Reads around constructs built from `template` (with '-' runs) and per-region `pools`.

    Includes: substitutions / N / lower-case anywhere, constructs at offset 0 and at the
    read end, reads shorter than the template, junk reads, and reads with two constructs
    (first-vs-best differences).
    """
    T = len(template)
    reads = []
    for _ in range(n):
        u = rng.random()
        if u < short_frac:
            reads.append(random_seq(rng, int(rng.integers(0, T))))
            continue
        if u < short_frac + junk_frac:
            reads.append(mutate(rng, random_seq(rng, int(rng.integers(T, T + 40))), 0, n_rate, lower_rate))
            continue

        def construct():
            vs = [p[int(rng.integers(0, len(p)))] for p in pools]
            c = fill_template(template, *vs)
            c = mutate(rng, c, sub_rate, n_rate, lower_rate)
            if strand == "reverse" or (strand == "both" and rng.random() < 0.5):
                c = revcomp(c)
            return c

        c = construct()
        if rng.random() < double_frac:
            c = c + random_seq(rng, int(rng.integers(0, 6))) + construct()
        e = rng.random()
        if e < edge_frac / 2:
            left, right = "", random_seq(rng, int(rng.integers(0, 20)))
        elif e < edge_frac:
            left, right = random_seq(rng, int(rng.integers(0, 20))), ""
        else:
            left, right = random_seq(rng, int(rng.integers(0, 25))), random_seq(rng, int(rng.integers(0, 25)))
        r = left + c + right
        if read_len is not None:
            r = (r + random_seq(rng, read_len))[:read_len]
        reads.append(r)
    return reads


def bgzf(data, block=65280, level=6):
    """Block gzip (BGZF, as written by bgzip / bcl-convert): gzip members of at most 64 KiB with a 'BC' extra field
    giving the member's size, closed by the standard empty member."""
    import struct
    import zlib
    out = []
    for at in list(range(0, len(data), block)) + [None]:
        chunk = b"" if at is None else data[at:at + block]
        comp = zlib.compressobj(level, zlib.DEFLATED, -15)
        raw = comp.compress(chunk) + comp.flush()
        bsize = 12 + 6 + len(raw) + 8
        out.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1) + raw +
                   struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    return b"".join(out)
