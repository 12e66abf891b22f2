"""Host-side logic of the product that needs no GPU: the C ABI loads and exports what include/scg.h
declares, the FASTQ reader/packer agrees with the reference's parser, validation errors carry the
reference's texts, the synthetic generator is reproducible, and -- without a device -- counting
fails loudly instead of falling back to a CPU path."""
import gzip
import os
import re

import numpy as np
import pytest

from fastq_cases import GOOD, BAD
from util import fastq, random_seq, mutate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_abi_exports_every_declared_symbol():
    from screencounter_b200._lib import lib, EXPORTS
    header = open(os.path.join(ROOT, "include", "scg.h")).read()
    declared = set(re.findall(r"\b(scg_[a-z0-9_]+)\s*\(", header))
    assert declared == set(EXPORTS), declared ^ set(EXPORTS)
    L = lib()
    for name in EXPORTS:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.scg_version()


def test_product_does_not_use_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "screencounter_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "kaori_port" not in text and "libkaori_ref" not in text and "liboracle" not in text, f


def _norm(seq):
    return "".join(c.upper() if c in "ACGTacgt" else "N" for c in seq)


@pytest.mark.parametrize("name", sorted(GOOD))
@pytest.mark.parametrize("nthreads", [1, 3])
def test_reader_matches_reference_parser(kref, name, nthreads):
    from screencounter_b200 import rcpp
    data = GOOD[name]
    assert rcpp.host_pack_roundtrip(data, nthreads) == [_norm(s) for s in kref.parse(data)]


@pytest.mark.parametrize("name", sorted(BAD))
def test_reader_errors(name):
    from screencounter_b200 import rcpp
    data, msg = BAD[name]
    with pytest.raises(Exception) as err:
        rcpp.host_pack_roundtrip(data)
    assert str(err.value) == msg


def test_reader_large_ragged_files_raw_and_gz(kref, tmp_path):
    from screencounter_b200 import rcpp
    rng = np.random.default_rng(0)
    reads = [mutate(rng, random_seq(rng, int(rng.integers(0, 300))), 0, 0.02, 0.05) for _ in range(30000)]
    reads[5] = random_seq(rng, 5000)
    data = fastq(reads)
    want = [_norm(s) for s in reads]
    assert [_norm(s) for s in kref.parse(data)] == want
    raw = tmp_path / "x.fastq"
    raw.write_bytes(data)
    gz = tmp_path / "x.fastq.gz"
    with gzip.open(gz, "wb") as f:
        f.write(data)
    for src in (data, str(raw), str(gz)):
        assert rcpp.host_pack_roundtrip(src, 4) == want
    with pytest.raises(Exception, match="failed to open file at"):
        rcpp.host_pack_roundtrip(str(tmp_path / "missing.fastq"))


VALIDATION = [
    # (call, args, message) -- texts of the reference
    ("single", dict(constant="ACGN--", pool=["AA"]), "unknown base 'N'"),
    ("single", dict(constant="ACGN--", pool=["AA"], strand=1), "cannot complement unknown base 'N'"),
    ("single", dict(constant="ACGT", pool=["AA"]), "expected one variable region in the constant template"),
    ("single", dict(constant="AC--GT--", pool=["AA"]), "expected one variable region in the constant template"),
    ("single", dict(constant="AC---", pool=["AA"]), "length of barcode_pool sequences (2) should be the same as the barcode_pool region (3)"),
    ("single", dict(constant="AC--", pool=["AA", "AAA"]), "variable regions should all have the same length (2)"),
    ("single", dict(constant="AC--", pool=["AA", "CC", "AA"]), "duplicate sequences detected (1, 3) when constructing the trie"),
    ("single", dict(constant="AC--", pool=["AR", "AG"]), "duplicate sequences detected (1, 2) when constructing the trie"),
    ("single", dict(constant="AC--", pool=["A!"]), "unknown base '!' detected when constructing the trie"),
    ("single", dict(constant="A" * 250 + "-" * 7, pool=["ACGTACG"]), "lacking compile-time support for constant regions longer than 256 bp"),
]


@pytest.mark.parametrize("kind,kw,msg", VALIDATION)
def test_validation_errors_need_no_device(kref, kind, kw, msg):
    from screencounter_b200 import rcpp
    args = (fastq(["ACGT"]), kw["constant"], kw.get("strand", 0), kw["pool"], 0, True)
    with pytest.raises(Exception) as err:
        rcpp.count_single_barcodes(*args, 1)
    assert str(err.value) == msg
    with pytest.raises(Exception) as ref_err:   # and the reference says the same
        kref.count_single(*args)
    assert str(ref_err.value) == msg


def test_other_entry_points_validate_like_the_reference():
    from screencounter_b200 import rcpp
    f = fastq(["ACGT"])
    with pytest.raises(Exception, match="expected 2 variable regions in the constant template"):
        rcpp.count_combo_barcodes_single(f, "AC--", 0, [["AA"], ["CC"]], 0, True, 1)
    with pytest.raises(Exception, match="currently expecting only 2 variable regions"):
        rcpp.count_combo_barcodes_single(f, "AC--", 0, [["AA"]], 0, True, 1)
    with pytest.raises(Exception, match=re.escape("length of variable region 2 (3) should be the same as its sequences (2)")):
        rcpp.count_combo_barcodes_single(f, "AC--G---", 0, [["AA"], ["CC"]], 0, True, 1)
    with pytest.raises(Exception, match="length of 'barcode_pools' should equal the number of variable regions"):
        rcpp.count_dual_barcodes_single_end(f, "AC--", [["AA"], ["CC"]], 0, 0, True, False, 1)
    with pytest.raises(Exception, match="both barcode pools should be of the same length"):
        rcpp.count_dual_barcodes(f, "AC--", False, 0, ["AA", "CC"], f, "AC--", False, 0, ["AA"], False, True, False, 1)
    with pytest.raises(Exception, match="expected one variable region in the second constant template"):
        rcpp.count_dual_barcodes(f, "AC--", False, 0, ["AA"], f, "ACGT", False, 0, ["AA"], False, True, False, 1)
    with pytest.raises(Exception, match=re.escape("duplicate sequences detected (1, 2) when constructing the trie")):
        rcpp.count_dual_barcodes(f, "AC--", False, 0, ["AA", "AA"], f, "AC--", False, 0, ["CC", "CC"], False, True, False, 1)
    with pytest.raises(Exception, match="expected at least one variable region"):
        rcpp.count_random_barcodes(f, "ACGT", 0, 0, True, 1)
    with pytest.raises(Exception, match="variable regions should all have the same length"):
        rcpp.match_barcodes(["AAA"], ["AAA", "AA"], 0, False)


@pytest.mark.skipif(_has_gpu(), reason="only meaningful where no CUDA device exists")
def test_no_device_is_a_loud_error_not_a_cpu_fallback():
    from screencounter_b200 import rcpp
    with pytest.raises(Exception, match="no usable CUDA device: this engine has no CPU fallback"):
        rcpp.count_single_barcodes(fastq(["ACGTAA"]), "AC--", 0, ["GT"], 0, True, 1)


def test_synthetic_generator_is_shard_reproducible(kref):
    from screencounter_b200.device import SynthSpec
    rng = np.random.default_rng(0)
    pool = ["".join("ACGT"[i] for i in rng.integers(0, 4, 20)) for _ in range(64)]
    template = "CAGCTACGTACG" + "-" * 20 + "CCAGCTCGATCG"
    spec = SynthSpec(template, [pool], seed=42, read_len=75, strand=2)
    whole = spec.fastq(0, 3000)
    per = 2 * 75 + 7
    assert len(whole) == 3000 * per
    assert spec.fastq(1000, 500) == whole[1000 * per:1500 * per]
    reads = kref.parse(whole)
    assert all(len(r) == 75 for r in reads)
    index, info = kref.trace_single(whole, template, 2, pool, 1, True)
    assert 0.75 < (index >= 0).mean() < 0.95           # ~90 % carry a construct, 1 % substitutions
    assert 0.3 < info[index >= 0, 1].mean() < 0.7      # about half on the reverse strand
    assert sum("N" in r for r in reads) > 0


@pytest.mark.parametrize("template,strand,mm,words", [
    ("CAGCTACGTACG" + "-" * 20 + "CCAGCTCGATCG", 2, 1, 3),
    ("ACGT----------TGCA", 0, 0, 2),
    ("ACGT----------TGCA", 1, 3, 5),
    ("A" * 60 + "-" * 8 + "C" * 60, 2, 2, 6),
])
def test_runtime_specialised_kernel_compiles(template, strand, mm, words):
    """The NVRTC step of the specialised scan kernel needs no device: 0 = compiled and loaded,
    2 = compiled, nothing to load it on."""
    import ctypes as C
    from screencounter_b200._lib import lib
    buf = C.create_string_buffer(16384)
    rc = lib().scg_jit_selftest(template.encode(), strand, mm, words, buf, C.c_size_t(16384))
    assert rc in (0, 2), buf.value.decode()


@pytest.mark.parametrize("template,strand,mm,read_len", [
    ("CAGCTACGTACG" + "-" * 20 + "CCAGCTCGATCG", 2, 1, 75),
    ("CAGCTACGTACG" + "-" * 20 + "CCAGCTCGATCG", 2, 0, 60),
    ("CAGCTACGTACGAA" + "-" * 20 + "CCAGCTCGAT", 2, 2, 50),      # asymmetric flanks: the strands' regions start apart
    ("ACGTACGT" + "-" * 10 + "TGCATGCA", 0, 1, 40),
    ("ACGTACGT" + "-" * 10 + "TGCATGCA", 1, 3, 26),
    ("A" * 40 + "-" * 30 + "C" * 40, 2, 1, 120),
    ("CAGCTACGTACG" + "-" * 20 + "CCAGCTCGATCG", 2, 1, 150),     # four window blocks
    ("CAGCTACGTACG" + "-" * 20 + "CCAGCTCGATCG", 2, 2, 192),     # five
    ("CAGCTACGTACG" + "-" * 20 + "CCAGCTCGATCG", 2, 1, 320),     # nine, the longest reads the kernel takes
    ("ACGTACGT" + "-" * 10 + "TGCATGCA", 0, 0, 101),
])
def test_uniform_length_kernel_compiles(template, strand, mm, read_len):
    """The uniform-length (filter + verify) variant through NVRTC, without a device."""
    import ctypes as C
    from screencounter_b200._lib import lib
    buf = C.create_string_buffer(16384)
    rc = lib().scg_jit_selftest_uniform(template.encode(), strand, mm, read_len, buf, C.c_size_t(16384))
    assert rc in (0, 2), buf.value.decode()


F12, R12 = "CAGCTACGTACG", "CCAGCTCGATCG"


@pytest.mark.parametrize("kind,a,strand_a,mm_a,b,strand_b,mm_b,read_len,use_first", [
    (1, F12 + "-" * 20 + R12, 0, 1, "GATTACAGGCTA" + "-" * 20 + "TTGACCGTAGCA", 0, 1, 75, 1),   # BASELINE configs[2]
    (1, F12 + "-" * 20 + R12, 1, 1, "GATTACAGGCTA" + "-" * 20 + "TTGACCGTAGCA", 1, 0, 75, 0),   # reverse strands, best mode
    (1, F12 + "-" * 8 + R12, 0, 2, "GATTACAGGCTA" + "-" * 30 + "TTGACCGTAGCA", 0, 0, 100, 1),   # uneven regions, key of 38 bases
    (1, "ACGTACGT" + "-" * 32 + "TGCATGCA", 0, 1, "GATTACAG" + "-" * 16 + "TTGACCGT", 1, 1, 60, 1),   # 32 + 16 = 48 bases
    (2, "CAGCTACG" + "-" * 20 + "GGTACCTT" + "-" * 20 + "CGATCGAG", 2, 1, "", 0, 0, 75, 1),     # BASELINE configs[3]
    (2, "CAGCTACG" + "-" * 20 + "GGTACCTT" + "-" * 10 + "CGATCGAG", 0, 0, "", 0, 0, 75, 0),
    (2, "CAGCTACG" + "-" * 6 + "GGTACCTT" + "-" * 9 + "CGATCGAGAA", 1, 2, "", 0, 0, 150, 1),
    (3, F12 + "-" * 16 + R12, 2, 1, "", 0, 0, 75, 1),                                           # BASELINE configs[4]
    (3, F12 + "-" * 16 + "CCAGCTCG", 2, 1, "", 0, 0, 75, 0),                                    # asymmetric flanks (Quirk B), best mode
    (3, F12 + "-" * 21 + R12, 0, 0, "", 0, 0, 101, 1),
    (4, F12 + "-" * 16 + R12, 2, 1, "", 0, 0, 75, 1),                                           # the same, barcodes listed by table part
    (4, F12 + "-" * 9 + R12, 1, 0, "", 0, 0, 150, 0),
])
def test_handler_kernels_compile(kind, a, strand_a, mm_a, b, strand_b, mm_b, read_len, use_first):
    """The specialised dual / combinatorial / random-barcode kernels (spec_handlers.cuh) through NVRTC, without a device."""
    import ctypes as C
    from screencounter_b200._lib import lib
    buf = C.create_string_buffer(16384)
    f = lib().scg_jit_selftest_handler
    f.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
    rc = f(kind, a.encode(), strand_a, mm_a, b.encode(), strand_b, mm_b, read_len, use_first, buf, 16384)
    assert rc in (0, 2), buf.value.decode()


def _tricky_fastq(rng, n, wrap_every=0, bad_at=None):
    """Records that defeat naive boundary guessing: qualities starting with '@' or '+', wrapped
    sequences and qualities, names holding '@' and '+', an occasional empty read."""
    out = []
    for i in range(n):
        L = int(rng.integers(0, 120)) if rng.random() < 0.9 else 0
        s = random_seq(rng, L)
        q = "".join("@+I#"[int(x)] for x in rng.integers(0, 4, size=L))
        if L and rng.random() < 0.3:
            q = "@" + q[1:]
        if wrap_every and L > wrap_every and rng.random() < 0.3:
            s = "\n".join(s[k:k + wrap_every] for k in range(0, L, wrap_every))
            q = "\n".join(q[k:k + wrap_every] for k in range(0, L, wrap_every))
        if bad_at is not None and i == bad_at:
            q = q + "I"   # one quality too many
        out.append("@r%d @x+y\n%s\n+anything@here\n%s\n" % (i, s, q))
    return "".join(out).encode("latin-1")


@pytest.mark.parametrize("wrap", [0, 40])
def test_parallel_splitter_equals_serial_and_reference(kref, wrap):
    """Inputs above 4 MB are split and parsed on several threads; the result must be the serial parse."""
    from screencounter_b200 import rcpp
    rng = np.random.default_rng(11 + wrap)
    data = _tricky_fastq(rng, 60000, wrap_every=wrap)
    assert len(data) > (4 << 20)
    serial = rcpp.host_pack_roundtrip(data, 1)
    assert serial == [_norm(s) for s in kref.parse(data)]
    for threads in (2, 5, 16):
        assert rcpp.host_pack_roundtrip(data, threads) == serial


def test_parallel_splitter_reports_the_serial_error(kref):
    from screencounter_b200 import rcpp
    rng = np.random.default_rng(3)
    data = _tricky_fastq(rng, 60000, bad_at=41234)
    msgs = []
    for threads in (1, 8):
        with pytest.raises(Exception) as err:
            rcpp.host_pack_roundtrip(data, threads)
        msgs.append(str(err.value))
    assert msgs[0] == msgs[1] == "non-equal lengths for quality and sequence strings (starting line %d)" % (4 * 41234 + 1)
    with pytest.raises(Exception) as err:
        kref.parse(data)
    assert msgs[0] in str(err.value)


@pytest.mark.parametrize("block", [200, 4096, 65280])
@pytest.mark.parametrize("threads", [1, 8])
def test_block_gzip_is_inflated_in_parallel_and_reads_the_same(kref, tmp_path, block, threads):
    """BGZF members are found from their headers and inflated on the thread pool; the text, and so every record, is
    what zlib's member-by-member gzread gives the reference.  Records straddle members at every block size."""
    import gzip
    from util import bgzf
    from screencounter_b200 import rcpp
    rng = np.random.default_rng(block + threads)
    data = _tricky_fastq(rng, 20000, wrap_every=40)
    path = tmp_path / "reads.fastq.gz"
    path.write_bytes(bgzf(data, block))
    assert gzip.decompress(path.read_bytes()) == data          # a valid multi-member gzip file
    want = kref.parse(str(path))
    got = rcpp.host_pack_roundtrip(str(path), threads)
    plain = rcpp.host_pack_roundtrip(data, threads)
    assert got == plain
    assert len(got) == len(want) == 20000
    norm = lambda r: "".join(c if c in "ACGT" else "N" for c in r.upper())
    assert got[:50] == [norm(r) for r in want[:50]]


def test_block_gzip_edge_cases(tmp_path, monkeypatch):
    from util import bgzf
    from screencounter_b200 import rcpp
    empty = tmp_path / "empty.fastq.gz"
    empty.write_bytes(bgzf(b""))
    assert rcpp.host_pack_roundtrip(str(empty), 4) == []
    good = fastq(["ACGT", "GGTTAA"])
    rng = np.random.default_rng(77)
    noisy = fastq([random_seq(rng, 60) for _ in range(3000)])   # text that does not compress to nothing
    broken = bytearray(bgzf(noisy, 3000))
    broken[18 + 40] ^= 0x5A                                      # a flipped byte inside the first member's deflate stream
    bad = tmp_path / "bad.fastq.gz"
    bad.write_bytes(bytes(broken))
    with pytest.raises(Exception):
        rcpp.host_pack_roundtrip(str(bad), 4)
    # a plain gzip member after BGZF members: not a BGZF chain, the serial gzip reader takes the file
    import gzip
    mixed = tmp_path / "mixed.fastq.gz"
    mixed.write_bytes(bgzf(good)[:-28] + gzip.compress(good))
    assert len(rcpp.host_pack_roundtrip(str(mixed), 4)) == 4


def test_block_gzip_on_the_host(tmp_path):
    """Host side of block gzip (csrc/bgzf.cpp): the library's compressor writes what gzip and the BGZF readers accept (members of at
    most 64 KiB with their 'BC' size field, closed by the empty member), and the host reader takes such an image from a file or
    from memory, inflating member-parallel, with the text's records."""
    import gzip
    import struct
    from screencounter_b200 import rcpp
    from util import bgzf
    rng = np.random.default_rng(4)
    reads = [mutate(rng, random_seq(rng, int(rng.integers(0, 200))), 0, 0.02, 0.05) for _ in range(20000)]
    text = fastq(reads)
    want = [_norm(s) for s in reads]
    for level, block in ((6, 0), (1, 5000), (0, 65280)):
        image = rcpp.bgzf_compress(text, level=level, block_text=block, nthreads=3).tobytes()
        assert gzip.decompress(image) == text
        at, members = 0, 0
        while at < len(image):
            assert image[at:at + 4] == b"\x1f\x8b\x08\x04" and image[at + 12:at + 16] == b"BC\x02\x00"
            at += struct.unpack("<H", image[at + 16:at + 18])[0] + 1
            members += 1
        assert at == len(image) and members == -(-len(text) // (block or 65280)) + 1
        assert image[-28:] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")   # bgzip's end-of-file marker
        assert rcpp.host_pack_roundtrip(image, 3) == want               # an image in memory
    path = tmp_path / "reads.fastq.gz"
    path.write_bytes(bgzf(text, 30000))
    assert rcpp.host_pack_roundtrip(str(path), 4) == want               # a file
    assert gzip.decompress(rcpp.bgzf_compress(b"").tobytes()) == b""
    damaged = bytearray(bgzf(text, 30000))
    damaged[5000] ^= 0xFF
    with pytest.raises(Exception, match="corrupt member"):
        rcpp.host_pack_roundtrip(bytes(damaged), 2)


def test_block_gzip_index_of_a_large_image_is_walked_in_pieces(monkeypatch):
    """bgzf_index (csrc/bgzf.cpp) walks an image of more than 8 MiB in pieces on the host pool, each piece starting where three
    well-formed members follow each other, and accepts the pieces only if their chains link up; the records are those of the
    text either way, and an image whose chain is broken in the middle is still refused."""
    from screencounter_b200 import rcpp
    rng = np.random.default_rng(21)
    reads = [random_seq(rng, 100) for _ in range(2000)]
    text = fastq(reads) * 60                                               # 120 000 records, 25 MB of text
    want = rcpp.host_pack_roundtrip(text, 4)
    assert len(want) == 120000
    for level, block in ((0, 65280), (1, 9000)):                           # stored members (image = 25 MB) and small deflated ones
        image = rcpp.bgzf_compress(text, level=level, block_text=block, nthreads=4).tobytes()
        if level == 0:
            assert len(image) > (16 << 20)                                 # at least four pieces
        assert rcpp.host_pack_roundtrip(image, 4) == want
        if level == 0:
            # the magic of a member in the middle is damaged: no chain from the start reaches the end
            at, k = 0, 0
            while k < 200:
                at += int.from_bytes(image[at + 16:at + 18], "little") + 1
                k += 1
            damaged = bytearray(image)
            damaged[at + 1] ^= 0x40
            with pytest.raises(Exception):
                rcpp.host_pack_roundtrip(bytes(damaged), 4)
