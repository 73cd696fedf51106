"""Python host mirror of NanoTel.R's interface for the per-read telomere-detection path.

R is not installable in this image, so the host side that `north_star` places in R is mirrored here in Python with
the reference's own function names, argument meaning and outputs; the R glue a maintainer would use instead is in
../R/ (INTEGRATION.md).  Every number below comes from libnanotel_b200.so (CUDA); this module only moves data,
assigns Serial numbers through the library's ntl_assign_serials, and writes the files NanoTel.R writes.

    search_patterns(...)             NanoTel.R:2001-2078   one call = one sequential pass over a read set
    filter_reads(...)                NanoTel.R:2123-2163
    run_future_worker_chuncks(...)   NanoTel.R:2171-2268   chunk loop: read --nrec records, rc, filter, 8-way Serial
    main(argv)                       NanoTel.R:19-99, 2304-2433   same command line flags, same output files
"""
from __future__ import annotations

import argparse
import gzip
import io
import os
import sys
import time
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .scanner import Scanner, assign_serials

VERSION = "Telomere Analyzer  version v1.1.9-beta 2026-02-19 (nanotel_b200 CUDA path)"

Read = Tuple[str, bytes]          # (header line without '>' / '@', sequence) -- one element of a DNAStringSet

SUMMARY_COLUMNS = ["Serial", "sequence_ID", "sequence_length",
                   "telo_density", "Telomere_start", "Telomere_end", "Telomere_length",
                   "telo_density_mismatch", "Telomere_start_mismatch", "Telomere_end_mismatch",
                   "Telomere_length_mismatch"]
TVR_COLUMNS = ["telo_density_mismatch_tvr", "Telomere_start_mismatch_tvr", "Telomere_end_mismatch_tvr",
               "Telomere_length_mismatch_tvr"]


# ------------------------------------------------------------------------------------------------ input
def _open_maybe_gzip(path: str):
    f = open(path, "rb")
    magic = f.read(2)
    f.seek(0)
    if magic == b"\x1f\x8b":
        return io.BufferedReader(gzip.GzipFile(fileobj=f), 1 << 20)
    return io.BufferedReader(f, 1 << 20)


def iter_records(paths: Sequence[str], fmt: str) -> Iterator[Read]:
    """readDNAStringSet semantics: FASTA may be multi-line, FASTQ is 4-line records, names are
    the full header line, qualities are ignored, gzip is transparent."""
    for p in paths:
        with _open_maybe_gzip(p) as f:
            if fmt == "fasta":
                name, parts = None, []
                for line in f:
                    line = line.rstrip(b"\r\n")
                    if line.startswith(b">"):
                        if name is not None:
                            yield name, b"".join(parts)
                        name, parts = line[1:].decode("utf-8", "replace"), []
                    elif line and name is not None:
                        parts.append(line)
                if name is not None:
                    yield name, b"".join(parts)
            elif fmt == "fastq":
                while True:
                    h = f.readline()
                    if not h:
                        break
                    s = f.readline().rstrip(b"\r\n")
                    f.readline()
                    f.readline()
                    if not h.startswith(b"@"):
                        raise ValueError("%s: malformed FASTQ record header %r" % (p, h[:40]))
                    yield h[1:].rstrip(b"\r\n").decode("utf-8", "replace"), s
            else:
                raise ValueError('format should be "fastq" or "fasta"')


def list_input_files(input_path: str) -> List[str]:
    """dir(full.names = TRUE, recursive = TRUE) for a directory, else the file itself (NanoTel.R:2174-2178)."""
    if os.path.isdir(input_path):
        out = []
        for root, _, files in os.walk(input_path):
            out.extend(os.path.join(root, f) for f in files)
        return sorted(out)
    return [input_path]


def iter_chunks(paths: Sequence[str], fmt: str, nrec: int) -> Iterator[List[Read]]:
    """readDNAStringSet(files, nrec = nrec) in a loop (NanoTel.R:2209-2217)."""
    chunk: List[Read] = []
    for rec in iter_records(paths, fmt):
        chunk.append(rec)
        if nrec > 0 and len(chunk) >= nrec:
            yield chunk
            chunk = []
    if chunk:
        yield chunk


class NativeReader:
    """ntl_reader_*: the library's FASTA/FASTQ(+gzip) reader.  Iterating yields (names, buf, offsets) per chunk of
    `nrec` records: one contiguous uint8 sequence buffer + n+1 offsets, copied out of the reader's buffers -- or, with
    copy=False, views of those buffers that are valid until the next chunk is requested (what the chunk loop uses: it
    is done with a chunk before it asks for the next one, and a copy of 230 MB per 10 000 reads is not free)."""

    def __init__(self, paths: Sequence[str], fmt: str, nrec: int, copy: bool = True):
        import ctypes as C
        self._C = C
        self._L = _lib.load()
        arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
        self._h = C.c_void_p()
        rc = self._L.ntl_reader_open(C.byref(self._h), arr, len(paths), fmt.encode())
        if rc != 0:
            raise ValueError('format should be "fastq" or "fasta"' if fmt not in ("fastq", "fasta") else "ntl_reader_open failed")
        self.nrec = int(nrec)
        self.copy = bool(copy)

    def __iter__(self):
        C = self._C
        sb, so, nb, no = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        try:
            while True:
                n = self._L.ntl_reader_next(self._h, self.nrec, C.byref(sb), C.byref(so), C.byref(nb), C.byref(no))
                if n < 0:
                    raise ValueError(self._L.ntl_reader_error(self._h).decode(errors="replace"))
                if n == 0:
                    return
                soff = np.ctypeslib.as_array(C.cast(so, C.POINTER(C.c_int64)), (n + 1,)).copy()
                noff = np.ctypeslib.as_array(C.cast(no, C.POINTER(C.c_int64)), (n + 1,)).copy()
                buf = np.ctypeslib.as_array(C.cast(sb, C.POINTER(C.c_uint8)), (max(int(soff[-1]), 1),))
                if self.copy:
                    buf = buf.copy()
                nraw = C.string_at(nb, int(noff[-1])) if noff[-1] > 0 else b""
                names = [nraw[int(noff[i]):int(noff[i + 1])].decode("utf-8", "replace") for i in range(n)]
                yield names, buf, soff
        finally:
            self.close()

    def close(self):
        if self._h and self._h.value:
            self._L.ntl_reader_close(self._h)
            self._h = self._C.c_void_p()


# ------------------------------------------------------------------------------------------------ tables
class _LazyChunk:
    """Sequence-of-(name, sequence) view over a reader chunk; sequences are materialised only for the kept reads."""

    def __init__(self, names, buf, off):
        self.names, self.buf, self.off = names, buf, off

    def __len__(self):
        return len(self.names)

    def __getitem__(self, i):
        return self.names[i], self.buf[int(self.off[i]):int(self.off[i + 1])].tobytes()

    def length(self, i) -> int:
        return int(self.off[i + 1] - self.off[i])


def _tokens(p) -> List[str]:
    if p is None:
        return []
    if isinstance(p, str):
        return p.split()
    return [str(x) for x in p]


def _rows_from_results(reads: Sequence[Read], res: np.ndarray, serial: np.ndarray, order: np.ndarray,
                       n_tracks: int):
    """One summary row per kept read (analyze_read's add_row, NanoTel.R:1923-1974), in `order`."""
    rows = []
    lazy = isinstance(reads, _LazyChunk)                     # a reader chunk knows the lengths without copying a read
    for i in order:
        r = res[i]
        row = [int(serial[i]), reads.names[i], reads.length(i)] if lazy else [int(serial[i]), reads[i][0], len(reads[i][1])]
        for t in range(n_tracks):
            tr = r["track"][t]
            if int(tr["start"]) == -1:                       # NanoTel.R:1926-1931: NA for this track
                row += [None, None, None, None]
            else:
                row += [float(tr["density"]), int(tr["start"]), int(tr["end"]), int(tr["end"]) - int(tr["start"]) + 1]
        rows.append(row)
    return rows


def _frame(rows, n_tracks: int):
    import pandas as pd
    cols = SUMMARY_COLUMNS + (TVR_COLUMNS if n_tracks == 3 else [])
    df = pd.DataFrame(rows, columns=cols)
    for c in cols:
        if c in ("Serial", "sequence_length") or c.startswith("Telomere_"):
            df[c] = df[c].astype("Int64")
    return df


def search_patterns(sample_telomeres: Sequence[Read], pattern_list, max_length=1e5, output_dir: Optional[str] = None,
                    serial_start: int = 1, min_density: float = 0.6, title: str = "Telomeric repeat density",
                    tvr_patterns=None, right_edge: bool = False, *, subseq_length: int = 100,
                    scanner: Optional[Scanner] = None, device: int = 0):
    """NanoTel.R:2001-2078.  `sample_telomeres` are the reads exactly as analyze_read would see them (already
    reverse-complemented / filtered by the caller, as in the reference).  Returns the summary data frame; when
    output_dir is given, reads/<Serial>.fasta.gz and density_vectors/read<Serial>.csv are written for every row
    (the reference writes the FASTA and three plots there, NanoTel.R:1870-1918).  `max_length` and `title` only
    affect the reference's plots and are accepted for signature parity."""
    own = scanner is None
    sc = scanner or Scanner(_tokens(pattern_list), _tokens(tvr_patterns) or None, min_density, subseq_length,
                            rc=False, use_filter=False, right_edge=right_edge, device=device)
    try:
        res = sc.scan([s for _, s in sample_telomeres])
        keep = (res["status"] & _lib.READ_KEEP) != 0
        serial = np.zeros(len(res), np.int32)
        serial[keep] = serial_start + np.arange(int(keep.sum()), dtype=np.int32)   # :2050-2069, one sequential pass
        order = np.nonzero(keep)[0]
        rows = _rows_from_results(sample_telomeres, res, serial, order, sc.n_tracks)
        if output_dir:
            write_read_outputs(output_dir, sample_telomeres, sc, res, serial, order)
        return _frame(rows, sc.n_tracks)
    finally:
        if own:
            sc.close()


def filter_reads(samples: Sequence[Read], patterns, do_rc: bool = True, num_of_cores: int = 10,
                 subread_width: int = 200, right_edge: bool = True, trimm_length: int = 70, *,
                 min_density: float = 0.6, device: int = 0) -> Optional[List[Read]]:
    """NanoTel.R:2123-2163.  Returns the reads that pass the edge filter (reverse-complemented first if do_rc, as the
    reference does), or None when none passes (the reference returns NA).  subread_width / trimm_length are fixed at
    200 / 70 in the kernel, as every call site of the reference uses them (NanoTel.R:2229)."""
    if subread_width != 200 or trimm_length != 70:
        raise ValueError("the CUDA edge filter implements the reference's only call: subread_width=200, trimm_length=70")
    with Scanner(_tokens(patterns), None, min_density, 100, rc=do_rc, use_filter=True, right_edge=right_edge,
                 device=device) as sc:
        res = sc.scan([s for _, s in samples])
    passed = (res["status"] & _lib.READ_FILTERED) == 0
    out = []
    for (name, seq), ok in zip(samples, passed):
        if ok:
            out.append((name, revcomp(seq) if do_rc else seq))
    return out or None


_COMP = bytes.maketrans(b"ACGTMRWSYKVHDBNacgtmrwsykvhdbn", b"TGCAKYWSRMBDHVNTGCAKYWSRMBDHVN")


def revcomp(seq: bytes) -> bytes:
    """Biostrings::reverseComplement (IUPAC aware, upper-case result) -- used only to WRITE reads/<Serial>.fasta.gz in
    the frame NanoTel.R writes them (NanoTel.R:2219-2221 then :1873); the scan itself complements on the fly."""
    return seq.translate(_COMP)[::-1]


# ------------------------------------------------------------------------------------------------ outputs
def fmt_double(x: float) -> str:
    """A double as readr::write_csv prints it: the shortest decimal string that reads back to the same value, and no
    decimal point for integral values (1 -> "1", 0 -> "0")."""
    x = float(x)
    if x == int(x) and abs(x) < 1e15:
        return str(int(x))
    return repr(x)


_fmt_double = fmt_double


def write_summary_csv(df, path: str) -> None:
    """readr::write_csv(df_summary) (NanoTel.R:2431-2432): NA for missing, integers plain, doubles shortest."""
    import csv
    with open(path, "w", newline="") as f:
        w = csv.writer(f, quoting=csv.QUOTE_MINIMAL, lineterminator="\n")
        w.writerow(list(df.columns))
        for row in df.itertuples(index=False):
            out = []
            for v in row:
                if v is None or v is np.nan or (hasattr(v, "__class__") and v.__class__.__name__ == "NAType"):
                    out.append("NA")
                elif isinstance(v, float):
                    out.append("NA" if v != v else _fmt_double(v))
                else:
                    out.append(str(v))
            w.writerow(out)


def fasta_record(name: str, seq: bytes) -> bytes:
    """One record as Biostrings::writeXStringSet prints it (NanoTel.R:1870-1873): '>' + the full header line, then the
    upper-case sequence 80 letters per line."""
    up = seq.upper()
    return b">" + name.encode() + b"\n" + b"\n".join(up[k:k + 80] for k in range(0, len(up), 80)) + b"\n"


def write_read_outputs(output_dir: str, reads: Sequence[Read], sc: Scanner, res: np.ndarray, serial: np.ndarray,
                       order: Iterable[int], rc_applied: bool = False) -> None:
    """Per telomeric read: reads/<Serial>.fasta.gz (writeXStringSet(compress = TRUE), NanoTel.R:1870-1873) and the
    per-window tables of every track (the `subs` data frames the plot functions receive, NanoTel.R:1876-1918) as
    density_vectors/read<Serial>.csv.  Plot rendering itself stays with the reference's R functions.
    The files are written by the library (ntl_write_read_outputs: all host threads, zlib level 6 = R's gzfile()
    default, doubles as shortest round-trip strings); `reads` is a reader chunk or any sequence of (name, sequence)."""
    import ctypes as C
    order = np.ascontiguousarray(np.asarray(list(order), dtype=np.int32))
    if len(order) == 0:
        os.makedirs(os.path.join(output_dir, "reads"), exist_ok=True)
        os.makedirs(os.path.join(output_dir, "density_vectors"), exist_ok=True)
        return
    if isinstance(reads, _LazyChunk):
        names, buf, off = reads.names, reads.buf, np.ascontiguousarray(reads.off, dtype=np.int64)
    else:
        names = [r[0] for r in reads]
        seqs = [bytes(r[1]) for r in reads]
        off = np.zeros(len(seqs) + 1, np.int64)
        np.cumsum([len(x) for x in seqs], out=off[1:])
        buf = np.frombuffer(b"".join(seqs) or b"\0", np.uint8)
    nb = [n.encode() for n in names]
    noff = np.zeros(len(nb) + 1, np.int64)
    np.cumsum([len(x) for x in nb], out=noff[1:])
    nblob = b"".join(nb) + b"\0"
    serial = np.ascontiguousarray(serial, dtype=np.int32)
    buf = np.ascontiguousarray(buf)
    os.makedirs(output_dir, exist_ok=True)
    rc_ = sc._L.ntl_write_read_outputs(sc._h, output_dir.encode(), buf.ctypes.data, off.ctypes.data, nblob,
                                       noff.ctypes.data, serial.ctypes.data, order.ctypes.data, len(order), sc.n_tracks,
                                       float(sc.params.min_density), int(bool(rc_applied)), os.cpu_count() or 1)
    if rc_ != 0:
        raise OSError("writing the per-read outputs under %s failed (ntl_write_read_outputs: %d)" % (output_dir, rc_))


# ------------------------------------------------------------------------------------------------ chunk loop
def run_future_worker_chuncks(input_path: str, output_path: Optional[str], format: str = "fastq", nrec: int = 10000,
                              patterns=None, do_rc: bool = False, use_filter: bool = False, right_edge: bool = True,
                              tvr_patterns=None, *, min_density: float = 0.6, subseq_length: int = 100,
                              device: int = 0, devices: Optional[Sequence[int]] = None, verbose: bool = True):
    """NanoTel.R:2171-2268.  The 8 forked search_patterns() futures of the reference are replaced by one
    ntl_scan_batch() per chunk -- sharded over `devices` when several GPUs are given -- ; Serial numbers and row order
    follow the reference's 8-way round-robin split (ntl_assign_serials).  A read on which NanoTel.R itself would
    stop() (NTL_READ_REF_ERROR, DESIGN.md "degenerate inputs") stops the run here too.
    Returns {"df_summary": DataFrame, "all_reads_length_vec": int array}."""
    files = list_input_files(input_path)
    all_len: List[np.ndarray] = []
    rows = []
    serial_start = 1
    with Scanner(_tokens(patterns), _tokens(tvr_patterns) or None, min_density, subseq_length, rc=do_rc,
                 use_filter=use_filter, right_edge=right_edge, device=device, devices=devices) as sc:
        if sc.note and verbose:
            print("note:", sc.note, file=sys.stderr)
        tm = {"reader_wait": 0.0, "scan": 0.0, "rows": 0.0, "outputs": 0.0}
        t_mark = time.perf_counter()
        for ci, (names, buf, soff) in enumerate(NativeReader(files, format, nrec, copy=False), 1):
            tm["reader_wait"] += time.perf_counter() - t_mark
            t_mark = time.perf_counter()
            if verbose:
                print(time.strftime("%Y-%m-%d %H:%M:%S"))
                print("processing chunk", ci, "...")
            all_len.append(np.diff(soff))                                               # :2225 (before the filter)
            res = sc.scan_concat(buf, soff)
            tm["scan"] += time.perf_counter() - t_mark
            t_mark = time.perf_counter()
            bad = np.flatnonzero(res["status"] & _lib.READ_REF_ERROR)
            if len(bad):
                raise RuntimeError("NanoTel.R would have stopped on read(s) %s of chunk %d (%s)" %
                                   (" ".join(str(int(i) + 1) for i in bad[:10]), ci, names[int(bad[0])]))
            serial, order, serial_start = assign_serials(res, serial_start)             # :2234-2258
            chunk = _LazyChunk(names, buf, soff)
            rows += _rows_from_results(chunk, res, serial, order, sc.n_tracks)
            tm["rows"] += time.perf_counter() - t_mark
            t_mark = time.perf_counter()
            if output_path:
                write_read_outputs(output_path, chunk, sc, res, serial, order, rc_applied=do_rc)
            tm["outputs"] += time.perf_counter() - t_mark
            t_mark = time.perf_counter()
        if verbose:
            print("timing (s): " + ", ".join("%s %.2f" % kv for kv in tm.items()), file=sys.stderr)
        n_tracks = sc.n_tracks
    return {"df_summary": _frame(rows, n_tracks),
            "all_reads_length_vec": np.concatenate(all_len) if all_len else np.zeros(0, np.int64)}


# ------------------------------------------------------------------------------------------------ command line
def build_parser() -> argparse.ArgumentParser:
    """The option list of NanoTel.R:30-92 (optparse), same names and defaults."""
    ap = argparse.ArgumentParser(prog="NanoTel", description="Telomere pattern finder (CUDA hot path)")
    ap.add_argument("-i", "--input_path", default=None, help="Path to input files.( dir or single file)")
    ap.add_argument("--save_path", default=None, help="A path to a directory for storing the output files.")
    ap.add_argument("--format", default="fastq", help='input files format ("fastq" (the default) or "fasta", gzip is supported)')
    ap.add_argument("-n", "--nrec", type=int, default=10000, help="maximum number of records to read per iteration")
    ap.add_argument("-r", "--rc", action="store_true", default=False, help="reverse complement the given reads")
    ap.add_argument("--patterns", default=None, help="Space separated list of pattern(s). Must be in double quotes.")
    ap.add_argument("--min_density", type=float, default=0.6)
    ap.add_argument("--subseq_length", type=int, default=100)
    ap.add_argument("--use_filter", action="store_true", default=False)
    ap.add_argument("--check_right_edge", action="store_true", default=False)
    ap.add_argument("--tvr_patterns", default=None)
    ap.add_argument("--version", action="store_true", default=False)
    ap.add_argument("--analysis", action="store_true", default=False,
                    help="post-processing of the reference (NanoTel.R:2438-2508): filtered, sorted summary + results.txt")
    ap.add_argument("--device", type=int, default=0, help="CUDA device ordinal (extension)")
    ap.add_argument("--devices", default=None,
                    help='space or comma separated CUDA ordinals, or "all": every --nrec chunk is sharded over them '
                         "(extension; replaces the 8 forked workers of NanoTel.R:2207)")
    return ap


def _r_number(x) -> str:
    """paste0(<numeric>) of R: up to 15 significant digits, no trailing zeros; NA / NaN as R prints them."""
    if x is None:
        return "NA"
    x = float(x)
    if x != x:
        return "NaN"
    return fmt_double(float("%.15g" % x))


def run_analysis(df, save_path: str, barcode_name: str):
    """The --analysis post-processing of NanoTel.R:2438-2508 (host side, no GPU): keep rows with telo_density_mismatch
    >= 0.75 and Telomere_start_mismatch <= 134, sort by sequence_length (longest first, ties in row order), add the
    running median of Telomere_length_mismatch and sequence_length minus it, drop rows where that difference is below
    134; writes <barcode>_filtered_sorted_summary.csv and <barcode>_results.txt.  The reference also draws
    <barcode>_telomere_plot.png with ggplot; here the three curves of that plot are written as
    <barcode>_telomere_plot_data.csv (read_index, sequence_length, Telomere_length_mismatch, TelLenMM_RunningMed).
    Returns (df_filtered, df_for_plot)."""
    import bisect
    import pandas as pd
    keep = (df["telo_density_mismatch"].astype("float64") >= 0.75) & (df["Telomere_start_mismatch"].astype("float64") <= 134)
    d = df[keep.fillna(False)].copy()                                                    # NA comparisons drop the row
    d = d.iloc[np.argsort(-d["sequence_length"].to_numpy(dtype=np.int64), kind="stable")].reset_index(drop=True)
    run_med, sorted_vals = [], []
    for v in d["Telomere_length_mismatch"].to_numpy(dtype=np.float64):
        bisect.insort(sorted_vals, float(v))
        m = len(sorted_vals)
        run_med.append(sorted_vals[m // 2] if m % 2 else 0.5 * (sorted_vals[m // 2 - 1] + sorted_vals[m // 2]))
    d["TelLenMM_RunningMed"] = np.array(run_med, dtype=np.float64)
    d["SeqLen_minus_RunMed"] = d["sequence_length"].to_numpy(dtype=np.float64) - d["TelLenMM_RunningMed"].to_numpy()
    df_for_plot = d.copy()
    df_for_plot["read_index"] = np.arange(1, len(d) + 1, dtype=np.int64)
    d = d[d["SeqLen_minus_RunMed"] >= 134].reset_index(drop=True)
    write_summary_csv(d, os.path.join(save_path, barcode_name + "_filtered_sorted_summary.csv"))
    n_reads = len(d)
    tl = d["Telomere_length_mismatch"].to_numpy(dtype=np.float64)
    med_telo = float(np.median(tl)) if n_reads else None
    pct_short = round(100.0 * float((tl < 2000).sum()) / n_reads, 1) if n_reads else float("nan")
    with open(os.path.join(save_path, barcode_name + "_results.txt"), "w") as f:
        f.write("Results for %s\n" % barcode_name)
        f.write("==========================================\n")
        f.write("Number of telomeric reads after filtration : %d\n" % n_reads)
        f.write("Median telomere length with mismatch (bp)  : %s\n" % _r_number(med_telo))
        f.write("%% of telomeres shorter than 2kb            : %s%%\n" % _r_number(pct_short))
    write_summary_csv(df_for_plot[["read_index", "sequence_length", "Telomere_length_mismatch", "TelLenMM_RunningMed"]],
                      os.path.join(save_path, barcode_name + "_telomere_plot_data.csv"))
    return d, df_for_plot


def main(argv: Optional[Sequence[str]] = None) -> int:
    opt = build_parser().parse_args(argv)
    if opt.version:
        print(VERSION)
        return 0
    if opt.patterns is None:
        sys.exit("Missing required parameter:  --patterns")
    if opt.save_path is None:
        sys.exit("Missing required parameter:  --save_path")
    if opt.input_path is None:
        sys.exit("Missing required parameter:  --input_path")
    os.makedirs(opt.save_path, exist_ok=True)
    t1 = time.time()
    devices = None
    if opt.devices:
        if opt.devices.strip() == "all":
            import torch
            devices = list(range(torch.cuda.device_count()))
        else:
            devices = [int(x) for x in opt.devices.replace(",", " ").split()]
    ans = run_future_worker_chuncks(opt.input_path, opt.save_path, opt.format, opt.nrec, opt.patterns, opt.rc,
                                    opt.use_filter, opt.check_right_edge, opt.tvr_patterns,
                                    min_density=opt.min_density, subseq_length=opt.subseq_length, device=opt.device,
                                    devices=devices)
    df = ans["df_summary"]
    barcode_name = os.path.basename(os.path.normpath(opt.input_path))
    write_summary_csv(df, os.path.join(opt.save_path, barcode_name + "_summary.csv"))        # :2430-2432
    with open(os.path.join(opt.save_path, "reads_ids.txt"), "w") as f:                        # :2433
        for sid in df["sequence_ID"]:
            f.write(str(sid) + "\n")
    lens = ans["all_reads_length_vec"]
    os.makedirs(os.path.join(opt.save_path, "log"), exist_ok=True)
    with open(os.path.join(opt.save_path, "log", "run.log"), "w") as f:
        f.write(VERSION + "\n")
        f.write("The patterns to search: %s\n" % opt.patterns)
        f.write("The sub-sequence length  is: %d\n" % opt.subseq_length)
        f.write("The minimal density for a telomeric subseq: %s\n" % opt.min_density)
        f.write("Total reads in sample: %d\n" % len(lens))
        f.write("Number of reads which identified as Telomeric: %d\n" % len(df))
        if len(lens):
            f.write("%% of total reads: %s%%\n" % round(100.0 * len(df) / len(lens), 2))
        f.write("Elapsed: %.3f s\n" % (time.time() - t1))
    if opt.analysis:
        run_analysis(df, opt.save_path, barcode_name)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
