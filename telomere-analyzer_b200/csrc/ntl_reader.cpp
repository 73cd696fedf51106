/*
 * ntl_reader.cpp -- native FASTA/FASTQ (+gzip) record reader for the host side of the path.
 *
 * Replaces, for callers that do not have Biostrings, `open_input_files()` + `readDNAStringSet(files, nrec =, format =)`
 * (NanoTel.R:2180, :2213; semantics SURVEY App. B.6): records are streamed `nrec` at a time across a list of files,
 * gzip is transparent (zlib's gzread also passes plain files through), FASTA may be multi-line, FASTQ is 4-line
 * records, names are the full header line without '>' / '@', qualities are skipped.  A chunk comes back as ONE
 * contiguous sequence buffer + offsets -- exactly what ntl_scan_batch_concat() takes -- so no per-read allocation
 * happens between the file and the packer.  The next chunk is inflated and parsed by a background thread while the
 * caller works on the current one.
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nanotel_b200.h"

namespace {

struct Chunk {
    std::vector<char> seq, names;
    std::vector<int64_t> seq_off, name_off;
    int32_t n = 0;
    int status = 0;                 /* 0 ok, < 0 error */
    void clear() { seq.clear(); names.clear(); seq_off.assign(1, 0); name_off.assign(1, 0); n = 0; status = 0; }
};

} // namespace

struct ntl_reader {
    std::vector<std::string> paths;
    bool fastq = true;
    size_t file_idx = 0;
    gzFile gz = nullptr;
    std::vector<char> buf;          /* inflate window */
    size_t pos = 0, end = 0;
    bool eof_file = true;
    std::string pending_header;     /* FASTA: header of the record that follows the one just finished */
    bool have_pending = false;
    char err[512] = "";

    Chunk cur, next;
    std::thread worker;
    bool prefetching = false;
    int32_t prefetch_nrec = 0;

    /* ---- buffered line reader over the current file */
    bool fill()
    {
        if (!gz) return false;
        if (pos > 0 && pos < end) memmove(buf.data(), buf.data() + pos, end - pos);
        end -= pos; pos = 0;
        if (end == buf.size()) buf.resize(buf.size() * 2);
        int got = gzread(gz, buf.data() + end, (unsigned)(buf.size() - end));
        if (got < 0) { int e; snprintf(err, sizeof err, "%s: %s", paths[file_idx].c_str(), gzerror(gz, &e)); return false; }
        if (got == 0) { eof_file = true; return false; }
        end += (size_t)got;
        return true;
    }
    /* next line [*b, *e) without the terminator; false at end of file */
    bool line(const char **b, const char **e)
    {
        for (;;) {
            const char *nl = (const char *)memchr(buf.data() + pos, '\n', end - pos);
            if (nl) {
                *b = buf.data() + pos; *e = nl;
                pos = (size_t)(nl - buf.data()) + 1;
                if (*e > *b && (*e)[-1] == '\r') (*e)--;
                return true;
            }
            if (!fill()) {
                if (err[0]) return false;
                if (pos < end) {                 /* last line without '\n' */
                    *b = buf.data() + pos; *e = buf.data() + end; pos = end;
                    if (*e > *b && (*e)[-1] == '\r') (*e)--;
                    return true;
                }
                return false;
            }
        }
    }
    bool open_next_file()
    {
        if (gz) { gzclose(gz); gz = nullptr; }
        have_pending = false;
        if (file_idx >= paths.size()) return false;
        gz = gzopen(paths[file_idx].c_str(), "rb");
        if (!gz) { snprintf(err, sizeof err, "cannot open %s", paths[file_idx].c_str()); return false; }
        gzbuffer(gz, 1u << 20);
        pos = end = 0; eof_file = false;
        return true;
    }

    /* one record into c; returns 1, 0 at end of all files, < 0 on error */
    int record(Chunk &c)
    {
        for (;;) {
            if (!gz) {
                if (file_idx >= paths.size()) return 0;
                if (!open_next_file()) return err[0] ? NTL_ERR_ARG : 0;
            }
            const char *b, *e;
            if (fastq) {
                if (!line(&b, &e)) { if (err[0]) return NTL_ERR_SEQUENCE; gzclose(gz); gz = nullptr; file_idx++; continue; }
                if (b == e) continue;                                   /* blank line between records */
                if (*b != '@') { snprintf(err, sizeof err, "%s: malformed FASTQ header", paths[file_idx].c_str()); return NTL_ERR_SEQUENCE; }
                c.names.insert(c.names.end(), b + 1, e);
                if (!line(&b, &e)) { snprintf(err, sizeof err, "%s: truncated FASTQ record", paths[file_idx].c_str()); return NTL_ERR_SEQUENCE; }
                c.seq.insert(c.seq.end(), b, e);
                const char *pb, *pe;
                if (!line(&pb, &pe) || pb == pe || *pb != '+' || !line(&pb, &pe)) {
                    snprintf(err, sizeof err, "%s: truncated FASTQ record", paths[file_idx].c_str());
                    return NTL_ERR_SEQUENCE;
                }
            } else {
                if (!have_pending) {
                    bool found = false;
                    while (line(&b, &e)) { if (b < e && *b == '>') { pending_header.assign(b + 1, e); found = true; break; } }
                    if (!found) { if (err[0]) return NTL_ERR_SEQUENCE; gzclose(gz); gz = nullptr; file_idx++; continue; }
                }
                c.names.insert(c.names.end(), pending_header.begin(), pending_header.end());
                have_pending = false;
                while (line(&b, &e)) {
                    if (b < e && *b == '>') { pending_header.assign(b + 1, e); have_pending = true; break; }
                    c.seq.insert(c.seq.end(), b, e);
                }
                if (err[0]) return NTL_ERR_SEQUENCE;
            }
            c.seq_off.push_back((int64_t)c.seq.size());
            c.name_off.push_back((int64_t)c.names.size());
            c.n++;
            return 1;
        }
    }
    void read_chunk(Chunk &c, int32_t nrec)
    {
        c.clear();
        while (nrec <= 0 || c.n < nrec) {
            int r = record(c);
            if (r < 0) { c.status = r; return; }
            if (r == 0) break;
        }
    }
};

extern "C" int ntl_reader_open(ntl_reader **out, const char *const *paths, int32_t n_paths, const char *format)
{
    if (!out || !paths || n_paths < 1 || !format) return NTL_ERR_ARG;
    *out = nullptr;
    const bool fq = strcmp(format, "fastq") == 0;
    if (!fq && strcmp(format, "fasta") != 0) return NTL_ERR_ARG;
    ntl_reader *r = new (std::nothrow) ntl_reader();
    if (!r) return NTL_ERR_NOMEM;
    r->fastq = fq;
    for (int32_t i = 0; i < n_paths; i++) {
        if (!paths[i]) { delete r; return NTL_ERR_ARG; }
        r->paths.push_back(paths[i]);
    }
    r->buf.resize(4u << 20);
    *out = r;
    return NTL_OK;
}

extern "C" int32_t ntl_reader_next(ntl_reader *r, int32_t nrec, const char **seq_buf, const int64_t **seq_off,
                                   const char **name_buf, const int64_t **name_off)
{
    if (!r) return NTL_ERR_ARG;
    if (r->prefetching && r->prefetch_nrec == nrec) {
        r->worker.join();
        r->prefetching = false;
        std::swap(r->cur, r->next);
    } else {
        if (r->prefetching) { r->worker.join(); r->prefetching = false; }   /* nrec changed: cannot happen in NanoTel */
        r->read_chunk(r->cur, nrec);
    }
    if (r->cur.status < 0) return r->cur.status;
    if (r->cur.n > 0) {                                  /* read ahead while the caller scans this chunk */
        r->prefetch_nrec = nrec;
        r->prefetching = true;
        r->worker = std::thread([r, nrec]() { r->read_chunk(r->next, nrec); });
    }
    if (seq_buf) *seq_buf = r->cur.seq.data();
    if (seq_off) *seq_off = r->cur.seq_off.data();
    if (name_buf) *name_buf = r->cur.names.data();
    if (name_off) *name_off = r->cur.name_off.data();
    return r->cur.n;
}

extern "C" const char *ntl_reader_error(const ntl_reader *r) { return r ? r->err : "NULL reader"; }

extern "C" void ntl_reader_close(ntl_reader *r)
{
    if (!r) return;
    if (r->prefetching) r->worker.join();
    if (r->gz) gzclose(r->gz);
    delete r;
}
