"""GPU: the NanoTel command line / host mirror produces the files NanoTel.R produces, with every value equal to the
oracle's restatement: <basename>_summary.csv, reads_ids.txt, reads/<Serial>.fasta.gz, density vectors."""
import csv
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

from nanotel_b200.nanotel import fmt_double
from oracle import oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _expected_rows(chunks, patterns, tvr, min_density, S, right_edge, rc, use_filter):
    """Summary rows per the oracle, chunk by chunk, with the reference's Serial logic."""
    P = O.make_params(patterns, tvr, min_density, S, right_edge)
    T = 3 if tvr else 2
    rows, serial_start, mx = [], 1, 0
    for chunk in chunks:
        seqs = [s for _, s in chunk]
        recs, passed, _, _ = O.scan_batch(P, seqs, do_rc=rc, use_filter=use_filter, n_threads=4, want_windows=False)
        idx = np.nonzero(passed)[0]
        serial, order, serial_start, mx = O.assign_serials(recs["keep"][idx].astype(np.int32), serial_start, mx)
        for j in order:
            i = idx[j]
            row = [str(serial[j]), chunk[i][0], str(len(seqs[i]))]
            for t in range(T):
                tr = recs[i]["t"][t]
                if int(tr["start"]) == -1:
                    row += ["NA"] * 4
                else:
                    row += [fmt_double(float(tr["density"])), str(int(tr["start"])), str(int(tr["end"])),
                            str(int(tr["end"]) - int(tr["start"]) + 1)]
            rows.append(row)
    return rows


def test_cli_on_the_reference_example(tmp_path):
    from nanotel_b200.nanotel import iter_chunks, main
    inp = os.path.join(GOLD, "sample.fasta")
    out = str(tmp_path / "out")
    fasta_sha = json.load(open(os.path.join(GOLD, "example_reads_fasta_sha256.json")))
    assert main(["-i", inp, "--save_path", out, "--format", "fasta", "--patterns", "TTAGGG", "--min_density", "0.6"]) == 0
    got = list(csv.reader(open(os.path.join(out, "sample.fasta_summary.csv"))))
    chunks = list(iter_chunks([inp], "fasta", 10000))
    exp = _expected_rows(chunks, "TTAGGG", None, 0.6, 100, False, False, False)
    assert got[0][:4] == ["Serial", "sequence_ID", "sequence_length", "telo_density"] and len(got[0]) == 11
    assert got[1:] == exp and len(exp) == 4
    ids = open(os.path.join(out, "reads_ids.txt")).read().split("\n")[:-1]
    assert ids == [r[1] for r in exp]
    for serial, (name, seq) in enumerate(chunks[0], 1):                     # all 4 reads are telomeric, Serial = order
        raw = gzip.open(os.path.join(out, "reads", "%d.fasta.gz" % serial)).read()
        txt = raw.decode().split("\n")
        assert txt[0] == ">" + name and "".join(txt[1:]) == seq.decode() and max(map(len, txt[1:])) == 80
        assert hashlib.sha256(raw).hexdigest() == fasta_sha["sha256"][str(serial)]      # the reference's own file
        dv = list(csv.DictReader(open(os.path.join(out, "density_vectors", "read%d.csv" % serial))))
        res = O.analyze_read(O.make_params("TTAGGG"), seq)
        st, en = O.split_telo(len(seq), 100)
        assert [int(r["start_index"]) for r in dv] == st.tolist()
        assert [r["density"] for r in dv] == [fmt_double(float(c) / float(w)) for c, w in zip(res.win_counts[0], en - st + 1)]
        assert [r["density_mismatch"] for r in dv] == [fmt_double(float(c) / float(w)) for c, w in zip(res.win_counts[1], en - st + 1)]


def test_cli_chunked_rc_filter_tvr_serials(tmp_path):
    """--nrec chunking with >= 8 reads per chunk (8-way Serial split), --rc, --use_filter, --tvr_patterns, gzip FASTQ."""
    from nanotel_b200.nanotel import iter_chunks, main
    from nanotel_b200.synth import as_list, synth_reads
    buf, off, meta = synth_reads(150, seed=5, median_len=3000, min_len=600, max_len=20000, telomeric_frac=0.5)
    seqs = as_list(buf, off)
    fq = str(tmp_path / "reads.fastq.gz")
    with gzip.open(fq, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b"@read%08d len=%d\n" % (i, len(s)) + s + b"\n+\n" + b"I" * len(s) + b"\n")
    out = str(tmp_path / "o")
    args = ["-i", fq, "--save_path", out, "-n", "37", "--rc", "--patterns", "YYAGGG", "--tvr_patterns",
            "TTGGG CCAGGG TCAGGG", "--use_filter", "--check_right_edge", "--min_density", "0.5", "--subseq_length", "200"]
    assert main(args) == 0
    got = list(csv.reader(open(os.path.join(out, "reads.fastq.gz_summary.csv"))))
    chunks = list(iter_chunks([fq], "fastq", 37))
    assert [len(c) for c in chunks] == [37, 37, 37, 37, 2]
    exp = _expected_rows(chunks, "YYAGGG", "TTGGG CCAGGG TCAGGG", 0.5, 200, True, True, True)
    assert len(got[0]) == 15 and got[1:] == exp and len(exp) > 5
    # the saved reads are in the reverse-complemented frame (NanoTel.R:2219-2221 before :1873)
    first = exp[0]
    i = int(first[1].split()[0][4:])
    txt = gzip.open(os.path.join(out, "reads", first[0] + ".fasta.gz")).read().decode().split("\n")
    assert "".join(txt[1:]) == O.revcomp(seqs[i]).decode()


def test_search_patterns_and_filter_reads_functions(example_reads):
    from nanotel_b200.nanotel import filter_reads, search_patterns
    df = search_patterns(example_reads, "TTAGGG", serial_start=7, min_density=0.6)
    assert df["Serial"].tolist() == [7, 8, 9, 10]
    P = O.make_params("TTAGGG")
    for row, (name, seq) in zip(df.itertuples(index=False), example_reads):
        r = O.analyze_read(P, seq).rec
        assert row.sequence_ID == name and row.sequence_length == len(seq)
        assert (row.Telomere_start, row.Telomere_end) == (int(r["t"][0]["start"]), int(r["t"][0]["end"]))
        assert row.telo_density_mismatch == float(r["t"][1]["density"])
    kept = filter_reads(example_reads, "TTAGGG", do_rc=False, right_edge=True)
    Pf = O.make_params("TTAGGG", right_edge=True)
    assert [n for n, _ in kept] == [n for n, s in example_reads if O.filter_read(Pf, s)]
    kept_left = filter_reads(example_reads, "TTAGGG", do_rc=False, right_edge=False)       # only the all-telomere read
    Pl = O.make_params("TTAGGG", right_edge=False)
    assert [n for n, _ in kept_left] == [n for n, s in example_reads if O.filter_read(Pl, s)] == [example_reads[0][0]]
    assert filter_reads(example_reads[1:], "TTAGGG", do_rc=False, right_edge=False) is None  # reference returns NA
