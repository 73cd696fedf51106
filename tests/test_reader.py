"""CPU: the native FASTA/FASTQ(+gzip) reader (ntl_reader_*) against the plain Python reader on awkward inputs:
multi-line FASTA, CRLF line ends, blank lines, no trailing newline, several files, gzip and plain, nrec chunking."""
import gzip

import numpy as np
import pytest


def _native(paths, fmt, nrec):
    from nanotel_b200.nanotel import NativeReader
    out = []
    for names, buf, off in NativeReader(paths, fmt, nrec):
        out.append([(names[i], buf[int(off[i]):int(off[i + 1])].tobytes()) for i in range(len(names))])
    return out


def test_fasta_variants(tmp_path):
    from nanotel_b200.nanotel import iter_chunks
    a = tmp_path / "a.fasta"
    a.write_bytes(b">r1 first read\nACGT\nACG\n\n>r2\r\nTTAGGG\r\nTT\r\n>r3 empty\n>r4\nGG")
    b = tmp_path / "b.fa.gz"
    with gzip.open(b, "wb") as f:
        f.write(b"; comment line before the first header\n>z1\n" + b"ACGT" * 5000 + b"\n>z2\nA\n")
    paths = [str(a), str(b)]
    for nrec in (1, 2, 3, 100, 0):
        got = _native(paths, "fasta", nrec)
        exp = list(iter_chunks(paths, "fasta", nrec))
        assert got == exp, nrec
    flat = [r for c in _native(paths, "fasta", 0) for r in c]
    assert [n for n, _ in flat] == ["r1 first read", "r2", "r3 empty", "r4", "z1", "z2"]
    assert flat[0][1] == b"ACGTACG" and flat[1][1] == b"TTAGGGTT" and flat[2][1] == b"" and flat[3][1] == b"GG"


def test_fastq_variants(tmp_path):
    from nanotel_b200.nanotel import iter_chunks
    rng = np.random.default_rng(3)
    recs = [("read%05d len=%d" % (i, L), bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), L)))
            for i, L in enumerate(rng.integers(1, 5000, 57))]
    gz = tmp_path / "x.fastq.gz"
    with gzip.open(gz, "wb") as f:
        for n, s in recs[:30]:
            f.write(b"@" + n.encode() + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")
    plain = tmp_path / "y.fastq"
    with open(plain, "wb") as f:
        for k, (n, s) in enumerate(recs[30:]):
            f.write(b"@" + n.encode() + b"\r\n" + s + b"\r\n+" + n.encode() + b"\r\n" + b"@" * len(s) + (b"" if k == 26 else b"\r\n"))
    paths = [str(gz), str(plain)]
    for nrec in (1, 7, 30, 31, 1000):
        got = _native(paths, "fastq", nrec)
        assert got == list(iter_chunks(paths, "fastq", nrec))
        assert [r for c in got for r in c] == recs


def test_reader_errors(tmp_path):
    from nanotel_b200.nanotel import NativeReader
    bad = tmp_path / "bad.fastq"
    bad.write_bytes(b"@r1\nACGT\n+\nIIII\nACGT\n")
    with pytest.raises(ValueError):
        list(NativeReader([str(bad)], "fastq", 10))
    with pytest.raises(ValueError):
        NativeReader([str(bad)], "bam", 10)
    with pytest.raises(ValueError):
        list(NativeReader([str(tmp_path / "missing.fastq")], "fastq", 10))


@pytest.mark.parametrize("ahead", ["1", "3", "8"])
def test_many_files_inflated_side_by_side(tmp_path, monkeypatch, ahead):
    """A list of many small files (what a nanopore run produces): the files are inflated by their own threads, up to
    NTL_READER_FILES ahead of the parser; records, their order and the nrec chunking must not depend on that, also when
    the reader is closed early or a file in the middle is missing."""
    from nanotel_b200.nanotel import NativeReader, iter_chunks
    monkeypatch.setenv("NTL_READER_FILES", ahead)
    monkeypatch.setenv("NTL_READER_MB", "16" if ahead == "3" else "2048")   # tiny budget: the file threads must block
    rng = np.random.default_rng(11)
    paths, recs = [], []
    for j in range(23):
        p = tmp_path / ("p%02d.fastq%s" % (j, ".gz" if j % 3 else ""))
        with (gzip.open(p, "wb") if j % 3 else open(p, "wb")) as f:
            for i in range(int(rng.integers(0, 9))):                     # some files are empty
                L = int(rng.integers(1, 20000 if j == 5 else 300))      # one file is much larger than a queue block
                n, s = "f%02d_r%d" % (j, i), bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), L))
                recs.append((n, s))
                f.write(b"@" + n.encode() + b"\n" + s + b"\n+\n" + b"#" * L + b"\n")
        paths.append(str(p))
    for nrec in (1, 5, 64, 0):
        got = _native(paths, "fastq", nrec)
        assert got == list(iter_chunks(paths, "fastq", nrec))
        assert [r for c in got for r in c] == recs
    it = iter(NativeReader(paths, "fastq", 3))                           # abandon the reader after the first chunk
    assert len(next(it)[0]) == 3
    it.close()
    missing = paths[:7] + [str(tmp_path / "nope.fastq.gz")] + paths[7:]
    with pytest.raises(ValueError, match="cannot open"):
        list(NativeReader(missing, "fastq", 1000))


def test_file_thread_blocks_on_a_full_queue(tmp_path, monkeypatch):
    """A file much larger than its share of the memory budget: its thread has to wait for the consumer again and
    again; nothing may be lost, reordered or deadlock, also when the reader is dropped half way."""
    from nanotel_b200.nanotel import NativeReader
    monkeypatch.setenv("NTL_READER_FILES", "8")
    monkeypatch.setenv("NTL_READER_MB", "16")                           # 4 batches of 2 MiB per file
    rng = np.random.default_rng(5)
    big = tmp_path / "big.fastq.gz"
    base = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 3000))
    n = 8000                                                            # 24 MB of sequence
    with gzip.open(big, "wb", compresslevel=1) as f:
        for i in range(n):
            s = base[i % 100:] + base[:i % 100]
            f.write(b"@r%d\n" % i + s + b"\n+\n" + b"I" * len(s) + b"\n")
    small = tmp_path / "small.fastq"
    small.write_bytes(b"@last\nACGT\n+\nIIII\n")
    got = 0
    for names, buf, off in NativeReader([str(big), str(small)], "fastq", 1500):
        for j, nm in enumerate(names):
            i = got + j
            if i < n:
                assert nm == "r%d" % i and buf[int(off[j]):int(off[j + 1])].tobytes() == base[i % 100:] + base[:i % 100]
            else:
                assert nm == "last" and buf[int(off[j]):int(off[j + 1])].tobytes() == b"ACGT"
        got += len(names)
    assert got == n + 1
    it = iter(NativeReader([str(big), str(small)], "fastq", 10))
    next(it)
    it.close()                                                          # the blocked file thread must let go


def test_varying_nrec_never_drops_records(tmp_path):
    """The ABI allows a different nrec on every ntl_reader_next call; the chunk read ahead with the previous size
    must be handed out (split or topped up), never dropped."""
    import ctypes as C
    import gzip
    from nanotel_b200 import _lib
    rng = np.random.default_rng(5)
    names, seqs = [], []
    paths = []
    for f in range(3):
        p = tmp_path / ("v%d.fastq.gz" % f)
        with gzip.open(p, "wb") as g:
            for i in range(41):
                s = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), int(rng.integers(1, 300))))
                nm = "r%d_%d" % (f, i)
                g.write(b"@" + nm.encode() + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")
                names.append(nm); seqs.append(s)
        paths.append(str(p))
    L = _lib.load()
    arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
    h = C.c_void_p()
    assert L.ntl_reader_open(C.byref(h), arr, len(paths), b"fastq") == 0
    sizes = [5, 1, 17, 3, 3, 40, 2, 9, 1000]
    got_names, got_seqs, k = [], [], 0
    sb, so, nb, no = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    while True:
        nrec = sizes[k % len(sizes)]; k += 1
        n = L.ntl_reader_next(h, nrec, C.byref(sb), C.byref(so), C.byref(nb), C.byref(no))
        assert n >= 0
        if n == 0:
            break
        assert n <= nrec
        soff = np.ctypeslib.as_array(C.cast(so, C.POINTER(C.c_int64)), (n + 1,))
        noff = np.ctypeslib.as_array(C.cast(no, C.POINTER(C.c_int64)), (n + 1,))
        sraw = C.string_at(sb, int(soff[-1]))
        nraw = C.string_at(nb, int(noff[-1]))
        for i in range(n):
            got_seqs.append(sraw[int(soff[i]):int(soff[i + 1])])
            got_names.append(nraw[int(noff[i]):int(noff[i + 1])].decode())
    L.ntl_reader_close(h)
    assert got_names == names and got_seqs == seqs


def _bgzf(data: bytes, block: int = 60000, level: int = 6) -> bytes:
    """bgzip's container: gzip members of at most 64 KiB, each with a "BC" extra field holding its size - 1, then the
    empty end-of-file block."""
    import struct
    import zlib
    out = []
    for k in list(range(0, len(data), block)) + [None]:
        chunk = b"" if k is None else data[k:k + block]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        cdata = c.compress(chunk) + c.flush()
        bsize = len(cdata) + 25                                        # header 12 + extra 6 + cdata + crc 4 + isize 4 - 1
        out.append(b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize) +
                   cdata + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
    return b"".join(out)


@pytest.mark.parametrize("threads", ["1", "4"])
def test_bgzf_files_are_inflated_by_a_pool(tmp_path, monkeypatch, threads):
    """A BGZF (bgzip) file is a chain of small gzip members that announce their size: with NTL_READER_BGZF_THREADS > 1
    its blocks are inflated by a pool of threads (1: zlib's gzread walks the members).  Same records either way, for
    FASTQ and multi-line FASTA, blocks cutting lines anywhere; gzip reads it too (it IS gzip); a corrupt block or a
    truncated file is an error, not silence."""
    from nanotel_b200.nanotel import NativeReader, iter_chunks
    monkeypatch.setenv("NTL_READER_BGZF_THREADS", threads)
    rng = np.random.default_rng(21)
    recs, fq, fa = [], [], []
    for i in range(700):
        L = int(rng.integers(1, 30000 if i % 50 == 0 else 3000))
        n, s = "read%04d some description" % i, bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), L))
        recs.append((n, s))
        fq.append(b"@" + n.encode() + b"\n" + s + b"\n+\n" + b"I" * L + b"\n")
        fa.append(b">" + n.encode() + b"\n" + b"\n".join(s[k:k + 70] for k in range(0, L, 70)) + b"\n")
    fq, fa = b"".join(fq), b"".join(fa)
    assert gzip.decompress(_bgzf(fq)) == fq                              # the container is what it claims to be
    for fmt, text in (("fastq", fq), ("fasta", fa)):
        for block in (60000, 4097):
            p = tmp_path / ("b%d.%s.gz" % (block, fmt))
            p.write_bytes(_bgzf(text, block))
            for nrec in (100, 0):
                got = _native([str(p)], fmt, nrec)
                assert [r for c in got for r in c] == recs
                assert got == list(iter_chunks([str(p)], fmt, nrec))
    # a list that mixes BGZF, plain gzip and uncompressed files
    p1, p2, p3 = tmp_path / "m1.fastq.gz", tmp_path / "m2.fastq.gz", tmp_path / "m3.fastq"
    p1.write_bytes(_bgzf(fq)); p2.write_bytes(gzip.compress(fq)); p3.write_bytes(fq)
    got = _native([str(p1), str(p2), str(p3)], "fastq", 512)
    assert [r for c in got for r in c] == recs * 3
    # corruption inside a block's deflate data, and a file cut in the middle of a block
    good = _bgzf(fq)
    bad = bytearray(good); bad[len(bad) // 2] ^= 0x55
    (tmp_path / "bad.fastq.gz").write_bytes(bytes(bad))
    with pytest.raises(ValueError):
        list(NativeReader([str(tmp_path / "bad.fastq.gz")], "fastq", 0))
    (tmp_path / "cut.fastq.gz").write_bytes(good[:len(good) // 2])
    with pytest.raises(ValueError):
        list(NativeReader([str(tmp_path / "cut.fastq.gz")], "fastq", 0))


def test_views_instead_of_copies(tmp_path):
    """copy=False (the chunk loop's mode): the sequence buffer of a chunk is a view of the reader's own buffer, valid
    until the next chunk is requested; consumed chunk by chunk it yields the same records as the copying mode."""
    from nanotel_b200.nanotel import NativeReader
    rng = np.random.default_rng(8)
    p = tmp_path / "v.fastq.gz"
    recs = []
    with gzip.open(p, "wb") as f:
        for i in range(900):
            L = int(rng.integers(1, 4000))
            n, s = "v%03d" % i, bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), L))
            recs.append((n, s))
            f.write(b"@" + n.encode() + b"\n" + s + b"\n+\n" + b"I" * L + b"\n")
    got = []
    for names, buf, off in NativeReader([str(p)], "fastq", 64, copy=False):
        assert not buf.flags["OWNDATA"]
        got += [(names[i], buf[int(off[i]):int(off[i + 1])].tobytes()) for i in range(len(names))]
    assert got == recs
    assert [r for c in _native([str(p)], "fastq", 64) for r in c] == recs
