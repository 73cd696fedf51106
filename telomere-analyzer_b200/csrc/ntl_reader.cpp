/*
 * ntl_reader.cpp -- native FASTA/FASTQ (+gzip) record reader for the host side of the path.
 *
 * Replaces, for callers that do not have Biostrings, `open_input_files()` + `readDNAStringSet(files, nrec =, format =)`
 * (NanoTel.R:2180, :2213; semantics SURVEY App. B.6): records are streamed `nrec` at a time across a list of files,
 * gzip is transparent (zlib's gzread also passes plain files through), FASTA may be multi-line, FASTQ is 4-line
 * records, names are the full header line without '>' / '@', qualities are skipped.  A chunk comes back as ONE
 * contiguous sequence buffer + offsets -- exactly what ntl_scan_batch_concat() takes -- so no per-read allocation
 * happens between the file and the packer.  The next chunk is parsed by a background thread while the caller works on
 * the current one, and the files of the list are inflated side by side: up to NTL_READER_FILES (default 8) files ahead
 * of the parser each have their own zlib thread feeding a bounded queue of 4 MiB blocks (a nanopore run is thousands
 * of small fastq.gz files; one gzip stream cannot be inflated in parallel, a list of them can).
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <zlib.h>
#include <algorithm>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nanotel_b200.h"

namespace {

/* One file being inflated by its own thread into a bounded queue of blocks. */
struct FileStream {
    static constexpr size_t BLOCK = 4u << 20;
    static constexpr size_t DEPTH = 8;
    struct Block { std::vector<char> data; size_t n = 0; };
    std::string path;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Block> q;
    bool finished = false;          /* producer pushed everything it will ever push */
    bool cancel = false;
    std::string error;              /* set before finished */

    explicit FileStream(const std::string &p) : path(p) { th = std::thread([this]() { produce(); }); }
    ~FileStream()
    {
        { std::lock_guard<std::mutex> g(mu); cancel = true; }
        cv.notify_all();
        if (th.joinable()) th.join();
    }
    void finish(const std::string &err)
    {
        { std::lock_guard<std::mutex> g(mu); error = err; finished = true; }
        cv.notify_all();
    }
    void produce()
    {
        gzFile gz = gzopen(path.c_str(), "rb");
        if (!gz) { finish("cannot open " + path); return; }
        gzbuffer(gz, 1u << 20);
        for (;;) {
            Block b;
            b.data.resize(BLOCK);
            const int got = gzread(gz, b.data.data(), (unsigned)BLOCK);
            if (got < 0) { int e; std::string m = path + ": " + gzerror(gz, &e); gzclose(gz); finish(m); return; }
            if (got == 0) break;
            b.n = (size_t)got;
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [this]() { return cancel || q.size() < DEPTH; });
            if (cancel) { lk.unlock(); gzclose(gz); return; }
            q.push_back(std::move(b));
            lk.unlock();
            cv.notify_all();
        }
        gzclose(gz);
        finish("");
    }
    /* next block, or false at the end of the file / on error (then *err is non-empty) */
    bool pop(Block *out, std::string *err)
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [this]() { return !q.empty() || finished; });
        if (!q.empty()) {
            *out = std::move(q.front());
            q.pop_front();
            lk.unlock();
            cv.notify_all();
            return true;
        }
        *err = error;
        return false;
    }
};

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
    bool gz = false;                /* a file is open (its stream is streams[file_idx]) */
    std::vector<std::unique_ptr<FileStream>> streams;   /* one slot per path; created up to `ahead` files early */
    size_t started = 0, ahead = 8;
    FileStream::Block blk;          /* block being copied into buf */
    size_t blk_pos = 0;
    std::vector<char> buf;          /* parse window */
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
        if (blk_pos == blk.n) {                       /* next block of the current file */
            std::string e;
            blk.n = 0; blk_pos = 0;
            if (!streams[file_idx]->pop(&blk, &e)) {
                if (!e.empty()) { snprintf(err, sizeof err, "%s", e.c_str()); return false; }
                eof_file = true;
                return false;
            }
        }
        const size_t take = std::min(buf.size() - end, blk.n - blk_pos);
        memcpy(buf.data() + end, blk.data.data() + blk_pos, take);
        blk_pos += take;
        end += take;
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
    void close_file()
    {
        if (gz && file_idx < streams.size()) streams[file_idx].reset();     /* joins the file's thread */
        gz = false;
    }
    bool open_next_file()
    {
        gz = false;
        have_pending = false;
        if (file_idx >= paths.size()) return false;
        if (streams.size() < paths.size()) streams.resize(paths.size());
        for (; started < paths.size() && started < file_idx + ahead; started++)
            streams[started].reset(new FileStream(paths[started]));
        gz = true;
        blk.n = 0; blk_pos = 0;
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
                if (!line(&b, &e)) { if (err[0]) return NTL_ERR_SEQUENCE; close_file(); file_idx++; continue; }
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
                    if (!found) { if (err[0]) return NTL_ERR_SEQUENCE; close_file(); file_idx++; continue; }
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
    if (const char *a = getenv("NTL_READER_FILES")) { const int v = atoi(a); if (v >= 1 && v <= 64) r->ahead = (size_t)v; }
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
    r->streams.clear();                                   /* cancels and joins the inflate threads */
    delete r;
}
