"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol include/nanotel_b200.h declares,
struct layouts match, the NVRTC specialisation compiles for sm_100a without a device, compute entry points fail loudly
without a GPU (no CPU fallback), and the host-side Serial / window logic equals the oracle's restatement."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from nanotel_b200 import _lib
    return _lib, _lib.load()


def test_every_declared_symbol_is_exported():
    _l, L = _lib()
    hdr = open(os.path.join(ROOT, "include", "nanotel_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ntl_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 19
    for sym in sorted(declared):
        assert hasattr(L, sym), "libnanotel_b200.so does not export %s" % sym
    assert declared == set(_l.EXPORTS), declared ^ set(_l.EXPORTS)
    assert L.ntl_version() == 200


def test_struct_layouts():
    _l, L = _lib()
    assert _l.RESULT_DTYPE.itemsize == 64
    assert C.sizeof(_l.Stage) == 32
    assert C.sizeof(_l.Timings) == 8 * 8 + 5 * 8 + 4 * 4


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from nanotel_b200 import NanoTelError, Scanner
    with pytest.raises(NanoTelError) as e:
        Scanner("TTAGGG")
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)


@pytest.mark.parametrize("pats,tvr,S", [("TTAGGG", None, 100), ("YYAGGG", "TTGGG CCAGGG TCAGGG", 100),
                                         ("CCCTAA CCCTGA", "CCCAA", 500), ("N", None, 16), ("ACGTACGTACGTACGTAC", None, 200),
                                         ("TTAGGG", None, 20), ("TTAGGG", None, 150), ("TTAGGG", None, 64)])
def test_nvrtc_specialisation_compiles_for_sm_100a(pats, tvr, S, tmp_path):
    _l, L = _lib()
    P = _l.make_params(pats, tvr, subseq_length=S, rc=True)
    log = C.create_string_buffer(1 << 16)
    cubin = str(tmp_path / "k.cubin").encode()
    n = L.ntl_jit_compile_check(C.byref(P), b"sm_100a", log, 1 << 16, cubin)
    assert n > 1000, log.value.decode(errors="replace")
    assert os.path.getsize(cubin) == n


@pytest.mark.parametrize("S", [7, 37, 333, 1, 43690])
def test_subseq_lengths_without_a_span_geometry_are_left_to_the_generic_kernel(S):
    """No divisor >= 16 of S has an odd part <= 32: there is no specialised build (ntl_create then runs the generic
    scan kernel and says so through ntl_scan_path / ntl_scan_path_note)."""
    _l, L = _lib()
    P = _l.make_params("TTAGGG", None, subseq_length=S)
    log = C.create_string_buffer(4096)
    assert L.ntl_jit_compile_check(C.byref(P), b"sm_100a", log, 4096, None) == _l.NTL_ERR_JIT
    assert b"generic" in log.value
    P = _l.make_params("TTAGGG", None, subseq_length=43691)
    assert L.ntl_jit_compile_check(C.byref(P), b"sm_100a", log, 4096, None) == _l.NTL_ERR_ARG


def test_precompiled_cubins_ship_next_to_the_library():
    """build.py specialises the span scan ahead of time for the default pattern sets: files named by their key, with
    a verified header, in nanotel_b200/precompiled/."""
    from nanotel_b200 import _lib as _l
    d = os.path.join(os.path.dirname(_l.LIB_PATH), "precompiled")
    files = [f for f in os.listdir(d) if f.endswith(".cubin")]
    assert len(files) >= 8
    for f in files:
        raw = open(os.path.join(d, f), "rb").read()
        assert raw[:8] == b"NTLCUBN2" and len(f) == 32 + 6
        ka, kb, size = np.frombuffer(raw[8:32], np.uint64)
        assert f == "%016x%016x.cubin" % (int(ka), int(kb)) and len(raw) == 40 + int(size)
        assert raw[40:44] == b"\x7fELF"


@pytest.mark.parametrize("bad,code", [((["TTAGGG", ""], None), -2), (("TTAGGX", None), -2), (("A" * 19, None), -2),
                                      (("TTAGGG", "TT-GG"), -2)])
def test_pattern_validation(bad, code):
    _l, L = _lib()
    P = _l.make_params(bad[0], bad[1])
    log = C.create_string_buffer(4096)
    assert L.ntl_jit_compile_check(C.byref(P), b"sm_100a", log, 4096, None) == code


def test_count_windows_matches_split_telo():
    from nanotel_b200 import count_windows
    from oracle import oracle as O
    rng = np.random.default_rng(1)
    for S in (1, 7, 20, 64, 100, 128, 200, 500, 4096):
        for L in list(range(1, 40)) + rng.integers(1, 300000, 200).tolist():
            assert count_windows(L, S) == O.count_windows(L, S), (L, S)


def test_assign_serials_follows_the_8_way_split():
    """NanoTel.R:2234-2258 against the oracle's restatement, incl. gaps, the < 8 sequential branch and filtering."""
    from nanotel_b200 import READ_FILTERED, READ_KEEP, RESULT_DTYPE, assign_serials
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    for n in [0, 1, 5, 7, 8, 9, 16, 17, 100, 1001]:
        for trial in range(4):
            res = np.zeros(n, RESULT_DTYPE)
            keep = rng.random(n) < 0.4
            filt = rng.random(n) < (0.3 if trial % 2 else 0.0)
            res["status"] = np.where(filt, READ_FILTERED, np.where(keep, READ_KEEP, 0))
            start = int(rng.integers(1, 50))
            serial, order, nxt = assign_serials(res, start)
            idx = np.nonzero(~filt)[0]
            o_serial, o_order, o_next, _ = O.assign_serials((keep & ~filt)[idx].astype(np.int32), start, start - 1)
            exp = np.zeros(n, np.int32)
            exp[idx] = o_serial
            assert np.array_equal(serial, exp)
            assert np.array_equal(order, idx[o_order])
            assert nxt == o_next
    # worked example: 10 reads, all kept -> groups {0,8},{1,9},{2},...; group g starts at serial_start + sum of sizes
    res = np.zeros(10, RESULT_DTYPE)
    res["status"] = READ_KEEP
    serial, order, nxt = assign_serials(res, 1)
    assert serial.tolist() == [1, 3, 5, 6, 7, 8, 9, 10, 2, 4]
    assert order.tolist() == [0, 8, 1, 9, 2, 3, 4, 5, 6, 7] and nxt == 11


def test_fasta_fastq_reader_and_chunking(tmp_path):
    import gzip
    from nanotel_b200.nanotel import iter_chunks, list_input_files, revcomp
    fa = tmp_path / "a.fasta"
    fa.write_text(">r1 desc\nACGT\nAC\n>r2\nTTAGGG\n")
    fq = tmp_path / "b.fastq.gz"
    with gzip.open(fq, "wt") as f:
        f.write("@q1 x\nACGTN\n+\nIIIII\n@q2\nGG\n+\nII\n")
    assert [c for c in iter_chunks([str(fa)], "fasta", 1)] == [[("r1 desc", b"ACGTAC")], [("r2", b"TTAGGG")]]
    assert [c for c in iter_chunks([str(fq)], "fastq", 10)] == [[("q1 x", b"ACGTN"), ("q2", b"GG")]]
    assert list_input_files(str(tmp_path)) == sorted([str(fa), str(fq)])
    from oracle import oracle as O
    for s in (b"ACGTNRYKMacgtn", b"TTAGGG"):
        assert revcomp(s) == O.revcomp(s)


def test_shard_bounds_cover_and_balance():
    from nanotel_b200.shard import shard_bounds
    rng = np.random.default_rng(0)
    lens = rng.integers(1000, 250000, 5000)
    for w in (1, 2, 4, 8):
        b = shard_bounds(lens, w)
        assert b[0][0] == 0 and b[-1][1] == len(lens)
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        per = [int(lens[s:e].sum()) for s, e in b]
        assert max(per) - min(per) <= 2 * 250000


def test_native_fasta_gz_writer_reproduces_the_reference_files(tmp_path):
    """ntl_write_fasta_gz (the writer behind ntl_write_read_outputs): the reference's reads/<Serial>.fasta byte for byte
    after gunzip; --rc frame = Biostrings::reverseComplement of the input (IUPAC letters, lower case folded)."""
    import gzip
    import hashlib
    import json
    from nanotel_b200 import _lib
    from nanotel_b200.nanotel import fasta_record, iter_chunks, revcomp
    L = _lib.load()
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sha = json.load(open(os.path.join(gold, "example_reads_fasta_sha256.json")))
    (chunk,) = list(iter_chunks([os.path.join(gold, "sample.fasta")], "fasta", 10000))
    for serial, (name, seq) in enumerate(chunk, 1):
        path = str(tmp_path / ("%d.fasta.gz" % serial))
        assert L.ntl_write_fasta_gz(path.encode(), name.encode(), seq, len(seq), 0) == 0
        raw = gzip.open(path).read()
        assert len(raw) == sha["bytes"][str(serial)] and hashlib.sha256(raw).hexdigest() == sha["sha256"][str(serial)]
    for seq in (b"ACGTNRYKMSWBDHVacgtnrykm" * 7, b"A", b"", b"ACGT" * 40, b"ACGT" * 40 + b"C"):
        path = str(tmp_path / "x.fasta.gz")
        assert L.ntl_write_fasta_gz(path.encode(), b"some read  with spaces", seq, len(seq), 1) == 0
        assert gzip.open(path).read() == fasta_record("some read  with spaces", revcomp(seq))
    assert L.ntl_write_fasta_gz(str(tmp_path / "no_such_dir" / "x.gz").encode(), b"n", b"A", 1, 0) == -8


def test_fasta_writer_reproduces_the_reference_files():
    """reads/<Serial>.fasta of the reference's own example run (writeXStringSet): same bytes from fasta_record()."""
    import hashlib
    import json
    from nanotel_b200.nanotel import fasta_record, iter_chunks
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sha = json.load(open(os.path.join(gold, "example_reads_fasta_sha256.json")))
    (chunk,) = list(iter_chunks([os.path.join(gold, "sample.fasta")], "fasta", 10000))
    assert len(chunk) == 4
    for serial, (name, seq) in enumerate(chunk, 1):
        raw = fasta_record(name, seq)
        assert len(raw) == sha["bytes"][str(serial)] and hashlib.sha256(raw).hexdigest() == sha["sha256"][str(serial)]


def test_fmt_double_is_readr_style():
    from nanotel_b200.nanotel import fmt_double
    assert fmt_double(1.0) == "1" and fmt_double(0.0) == "0" and fmt_double(0.5) == "0.5"
    assert fmt_double(0.9919354838709677) == "0.9919354838709677" and fmt_double(8 / 21) == repr(8 / 21)
    assert fmt_double(1234.5) == "1234.5" and fmt_double(2981.0) == "2981"


def test_analysis_post_processing(tmp_path):
    """--analysis (NanoTel.R:2438-2508): filter, sort, running median, second filter, the two output files --
    against a literal restatement of the R pipeline on a table with NA rows, ties and both filters biting."""
    import csv
    import statistics
    from nanotel_b200.nanotel import _frame, run_analysis
    rng = np.random.default_rng(4)
    rows = []
    for i in range(240):
        L = int(rng.choice([1200, 2000, 3000, 5000, 5000, 8000, 12000, 20000]))   # ties in sequence_length
        tl = int(rng.integers(200, min(L, 9000)))
        dens_mm = float(rng.choice([0.5, 0.74, 0.75, 0.8, 0.97, 1.0]))
        st_mm = int(rng.choice([1, 20, 134, 135, 400]))
        row = [i + 1, "read%02d extra" % i, L, 0.9, 1, tl, tl, dens_mm, st_mm, st_mm + tl - 1, tl]
        if i % 11 == 0:
            row[7:11] = [None, None, None, None]                               # NA on the mismatch track
        rows.append(row)
    df = _frame(rows, 2)
    got, for_plot = run_analysis(df, str(tmp_path), "bc01")
    # R: filter -> arrange(desc(sequence_length)) (stable) -> running median -> filter
    kept = [r for r in rows if r[7] is not None and r[7] >= 0.75 and r[8] <= 134]
    kept.sort(key=lambda r: -r[2])
    exp = []
    for i, r in enumerate(kept):
        med = statistics.median([x[10] for x in kept[:i + 1]])
        exp.append((r[0], r[2], r[10], float(med), r[2] - float(med)))
    assert [tuple(x) for x in for_plot[["Serial", "sequence_length", "Telomere_length_mismatch", "TelLenMM_RunningMed",
                                        "SeqLen_minus_RunMed"]].itertuples(index=False)] == exp
    assert for_plot["read_index"].tolist() == list(range(1, len(exp) + 1))
    final = [e for e in exp if e[4] >= 134]
    assert 0 < len(final) < len(exp) < len(rows)
    assert got["Serial"].tolist() == [e[0] for e in final]
    out = list(csv.reader(open(tmp_path / "bc01_filtered_sorted_summary.csv")))
    assert out[0][-2:] == ["TelLenMM_RunningMed", "SeqLen_minus_RunMed"] and len(out) == len(final) + 1
    assert [r[0] for r in out[1:]] == [str(e[0]) for e in final]
    txt = open(tmp_path / "bc01_results.txt").read().split("\n")
    assert txt[0] == "Results for bc01" and txt[2].endswith(": %d" % len(final))
    med = statistics.median([e[2] for e in final])
    assert txt[3].endswith(": " + (str(int(med)) if med == int(med) else repr(float(med))))
    pct = round(100.0 * sum(e[2] < 2000 for e in final) / len(final), 1)
    assert txt[4].endswith(": " + (str(int(pct)) if pct == int(pct) else repr(pct)) + "%")
    assert len(list(csv.reader(open(tmp_path / "bc01_telomere_plot_data.csv")))) == len(exp) + 1
    # nothing survives: R prints NA and NaN
    got0, _ = run_analysis(df.iloc[:0], str(tmp_path), "empty")
    assert len(got0) == 0 and "NA" in open(tmp_path / "empty_results.txt").read()


def test_host_memory_probe_runs_without_a_device():
    """ntl_host_read_gbs (the packer's ceiling that bench.py reports beside the packer's rate): plain reads and the
    read + quarter-size write pattern both return a positive rate; bad arguments return 0."""
    from nanotel_b200 import _lib
    L = _lib.load()
    a = np.random.default_rng(3).integers(65, 85, 32 << 20, dtype=np.uint8)
    assert L.ntl_host_read_gbs(a.ctypes.data, a.nbytes, 2, 2) > 0.05
    assert L.ntl_host_read_gbs(a.ctypes.data, a.nbytes, 2, -2) > 0.05
    assert L.ntl_host_read_gbs(None, a.nbytes, 2, 2) == 0.0 and L.ntl_host_read_gbs(a.ctypes.data, 0, 2, 2) == 0.0


def test_summary_rows_from_a_reader_chunk_equal_rows_from_a_list():
    """_rows_from_results on a reader chunk (lengths from the offsets, no read is copied) = on a list of (name, seq)."""
    from nanotel_b200 import _lib
    from nanotel_b200 import nanotel as N
    res = np.zeros(3, _lib.RESULT_DTYPE)
    res["track"]["start"][:] = [[5, -1, 0], [1, 2, 0], [7, 8, 0]]
    res["track"]["end"][:] = [[50, 0, 0], [10, 20, 0], [70, 80, 0]]
    res["track"]["density"][:] = 0.5
    buf, off = np.frombuffer(b"ACGTACGTAAACG", np.uint8), np.array([0, 4, 8, 13])
    chunk = N._LazyChunk(["a", "b x", "c"], buf, off)
    lst = [("a", b"ACGT"), ("b x", b"ACGT"), ("c", b"AAACG")]
    serial = np.array([1, 2, 3])
    assert N._rows_from_results(chunk, res, serial, [2, 0], 2) == N._rows_from_results(lst, res, serial, [2, 0], 2)
    assert N._rows_from_results(chunk, res, serial, [2, 0], 2)[0][:3] == [3, "c", 5]
