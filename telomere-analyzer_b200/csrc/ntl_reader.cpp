/*
 * ntl_reader.cpp -- native FASTA/FASTQ (+gzip) record reader for the host side of the path.
 *
 * Replaces, for callers that do not have Biostrings, `open_input_files()` + `readDNAStringSet(files, nrec =, format =)`
 * (NanoTel.R:2180, :2213; semantics SURVEY App. B.6): records are streamed `nrec` at a time across a list of files,
 * gzip is transparent (zlib's gzread also passes plain files through), FASTA may be multi-line, FASTQ is 4-line
 * records, names are the full header line without '>' / '@', qualities are skipped.  A chunk comes back as ONE
 * contiguous sequence buffer + offsets -- exactly what ntl_scan_batch_concat() takes -- so no per-read allocation
 * happens between the file and the packer.  The next chunk is parsed by a background thread while the caller works on
 * the current one, and the files of the list are inflated side by side: up to NTL_READER_FILES (default: the core count, at most 16) files ahead
 * of the parser each have their own thread that inflates AND parses the file into a bounded queue of record batches
 * (a nanopore run is thousands of small fastq.gz files; one gzip stream cannot be inflated in parallel, a list of
 * them can); the consumer only concatenates batches into --nrec chunks.
 * A file in BGZF form (bgzip / htslib: a chain of gzip members of <= 64 KiB, each announcing its compressed size in a
 * "BC" extra field) CAN be inflated in parallel: when fewer files than cores are open, such a file gets a pool of
 * inflate threads of its own (BgzfSource below: blocks are read sequentially, inflated in groups by the pool, handed
 * to the line parser in order).  Plain single-member gzip has no such seams and stays on one thread (zlib's gzread).
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

/* Records of one file, parsed: sequence letters and header lines back to back, with their cumulative ends. */
struct Batch {
    std::vector<char> seq, names;
    std::vector<int64_t> seq_end, name_end;
    int32_t n = 0;
    void add(const std::string &name) { names.insert(names.end(), name.begin(), name.end()); }
    void close_record() { seq_end.push_back((int64_t)seq.size()); name_end.push_back((int64_t)names.size()); n++; }
};

/* A BGZF file inflated by a pool of threads.  The owner (the file's producer thread) reads the compressed blocks in
 * order and groups ~1 MiB of them into a job; the pool inflates jobs (raw deflate per block, CRC32 checked); next()
 * hands out the decompressed bytes of the jobs in file order. */
struct BgzfSource {
    struct Job {
        std::vector<unsigned char> raw;                 /* the blocks' deflate data, back to back */
        std::vector<uint32_t> clen, isize, crc;         /* per block */
        std::vector<char> out;
        std::string err;
        bool done = false;
    };
    FILE *f = nullptr;
    bool eof = false;
    std::string read_err;
    std::deque<std::shared_ptr<Job>> order;             /* jobs in file order, oldest first */
    std::deque<std::shared_ptr<Job>> todo;              /* not yet taken by a pool thread */
    std::vector<std::thread> pool;
    std::mutex mu;
    std::condition_variable cv_todo, cv_done;
    bool stop = false;

    /* true if the file starts with a BGZF block header */
    static bool probe(const char *path)
    {
        FILE *g = fopen(path, "rb");
        if (!g) return false;
        unsigned char h[18];
        const bool ok = fread(h, 1, 18, g) == 18 && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 4) &&
                        h[10] == 6 && h[11] == 0 && h[12] == 'B' && h[13] == 'C' && h[14] == 2 && h[15] == 0;
        fclose(g);
        return ok;
    }
    BgzfSource(const char *path, int n_threads)
    {
        f = fopen(path, "rb");
        if (!f) { read_err = "cannot open file"; eof = true; return; }
        setvbuf(f, nullptr, _IOFBF, 4u << 20);
        for (int t = 0; t < n_threads; t++) pool.emplace_back([this]() { work(); });
    }
    ~BgzfSource()
    {
        { std::lock_guard<std::mutex> g(mu); stop = true; }
        cv_todo.notify_all();
        for (auto &t : pool) t.join();
        if (f) fclose(f);
    }
    void work()
    {
        for (;;) {
            std::shared_ptr<Job> j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_todo.wait(lk, [this]() { return stop || !todo.empty(); });
                if (stop) return;
                j = todo.front(); todo.pop_front();
            }
            size_t total = 0;
            for (uint32_t n : j->isize) total += n;
            j->out.resize(total);
            size_t in_pos = 0, out_pos = 0;
            for (size_t b = 0; b < j->clen.size() && j->err.empty(); b++) {
                z_stream zs;
                memset(&zs, 0, sizeof zs);
                if (inflateInit2(&zs, -15) != Z_OK) { j->err = "inflateInit2 failed"; break; }
                zs.next_in = j->raw.data() + in_pos; zs.avail_in = j->clen[b];
                zs.next_out = (Bytef *)j->out.data() + out_pos; zs.avail_out = j->isize[b];
                const int r = inflate(&zs, Z_FINISH);
                inflateEnd(&zs);
                if (r != Z_STREAM_END || zs.avail_out != 0) { j->err = "corrupt BGZF block"; break; }
                if (j->isize[b] && (uint32_t)crc32(0L, (const Bytef *)j->out.data() + out_pos, j->isize[b]) != j->crc[b]) {
                    j->err = "BGZF block fails its CRC"; break;
                }
                in_pos += j->clen[b]; out_pos += j->isize[b];
            }
            { std::lock_guard<std::mutex> g(mu); j->done = true; }
            cv_done.notify_all();
        }
    }
    /* the next ~1 MiB of blocks as a job; nullptr at the end of the file or on a malformed header (read_err set) */
    std::shared_ptr<Job> read_job()
    {
        if (eof) return nullptr;
        auto j = std::make_shared<Job>();
        while (j->raw.size() < (1u << 20)) {
            unsigned char h[12];
            const size_t got = fread(h, 1, 12, f);
            if (got == 0) { eof = true; break; }
            if (got != 12 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) { read_err = "not a BGZF block header"; eof = true; break; }
            const unsigned xlen = h[10] | (h[11] << 8);
            unsigned char extra[65536];
            if (fread(extra, 1, xlen, f) != xlen) { read_err = "truncated BGZF block"; eof = true; break; }
            int bsize = -1;
            for (unsigned o = 0; o + 4 <= xlen;) {
                const unsigned slen = extra[o + 2] | (extra[o + 3] << 8);
                if (extra[o] == 'B' && extra[o + 1] == 'C' && slen == 2 && o + 6 <= xlen) bsize = extra[o + 4] | (extra[o + 5] << 8);
                o += 4 + slen;
            }
            const long clen = (long)bsize - (long)xlen - 19;
            if (bsize < 0 || clen < 0) { read_err = "BGZF block without a size field"; eof = true; break; }
            const size_t at = j->raw.size();
            j->raw.resize(at + (size_t)clen);
            unsigned char tr[8];
            if (fread(j->raw.data() + at, 1, (size_t)clen, f) != (size_t)clen || fread(tr, 1, 8, f) != 8) {
                read_err = "truncated BGZF block"; eof = true; j->raw.resize(at); break;
            }
            j->clen.push_back((uint32_t)clen);
            j->crc.push_back(tr[0] | (tr[1] << 8) | (tr[2] << 16) | ((uint32_t)tr[3] << 24));
            j->isize.push_back(tr[4] | (tr[5] << 8) | (tr[6] << 16) | ((uint32_t)tr[7] << 24));
        }
        return j->clen.empty() ? nullptr : j;
    }
    /* appends the next job's bytes to buf at *end (growing buf); false at the end of the file or on error (*err set) */
    bool next(std::vector<char> &buf, size_t *end, std::string *err)
    {
        for (;;) {
            while (order.size() < 2 * pool.size() + 1) {         /* keep the pool busy */
                auto j = read_job();
                if (!j) break;
                { std::lock_guard<std::mutex> g(mu); todo.push_back(j); }
                order.push_back(j);
                cv_todo.notify_one();
            }
            if (order.empty()) { *err = read_err; return false; }
            auto j = order.front();
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_done.wait(lk, [&]() { return j->done; });
            }
            order.pop_front();
            if (!j->err.empty()) { *err = j->err; return false; }
            if (j->out.empty()) continue;                         /* only empty blocks (the end-of-file marker) */
            if (buf.size() < *end + j->out.size()) buf.resize(std::max(buf.size() * 2, *end + j->out.size()));
            memcpy(buf.data() + *end, j->out.data(), j->out.size());
            *end += j->out.size();
            return true;
        }
    }
};

/* One file of the list: its own thread inflates AND parses it into a bounded queue of record batches. */
struct FileStream {
    static constexpr size_t BATCH_BYTES = 2u << 20;
    static constexpr int32_t BATCH_RECORDS = 1024;
    size_t depth;                   /* batches a file may run ahead of the consumer (memory budget / files ahead) */
    std::string path;
    bool fastq;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Batch> q;
    bool finished = false;          /* the producer pushed everything it will ever push */
    bool cancel = false;
    std::string error;              /* set before finished */

    /* producer-private: buffered line reader over the gzip stream */
    int inflate_threads = 1;        /* > 1: a BGZF file is inflated by a pool of that many threads */
    std::unique_ptr<BgzfSource> bgzf;
    gzFile gz = nullptr;
    std::vector<char> buf;
    size_t pos = 0, end = 0;
    std::string zerr;

    FileStream(const std::string &p, bool fq, size_t d, int inflate_thr = 1) : depth(d), path(p), fastq(fq), inflate_threads(inflate_thr)
    {
        th = std::thread([this]() { produce(); });
    }
    ~FileStream()
    {
        { std::lock_guard<std::mutex> g(mu); cancel = true; }
        cv.notify_all();
        if (th.joinable()) th.join();
    }
    void finish(const std::string &err)
    {
        if (gz) { gzclose(gz); gz = nullptr; }
        bgzf.reset();
        { std::lock_guard<std::mutex> g(mu); error = err; finished = true; }
        cv.notify_all();
    }
    bool push(Batch &&b)            /* false: the reader was closed */
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [this]() { return cancel || q.size() < depth; });
        if (cancel) return false;
        q.push_back(std::move(b));
        lk.unlock();
        cv.notify_all();
        return true;
    }
    /* next batch, or false at the end of the file / on error (then *err is non-empty) */
    bool pop(Batch *out, std::string *err)
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

    bool fill()
    {
        if (pos > 0 && pos < end) memmove(buf.data(), buf.data() + pos, end - pos);
        end -= pos; pos = 0;
        if (bgzf) {
            std::string e;
            if (bgzf->next(buf, &end, &e)) return true;
            if (!e.empty()) zerr = path + ": " + e;
            return false;
        }
        if (end == buf.size()) buf.resize(buf.size() * 2);
        const int got = gzread(gz, buf.data() + end, (unsigned)std::min<size_t>(buf.size() - end, 1u << 30));
        if (got < 0) { int e; zerr = path + ": " + gzerror(gz, &e); return false; }
        if (got == 0) {                 /* end of file -- or a stream that stops in the middle of a member */
            int e = Z_OK;
            const char *msg = gzerror(gz, &e);
            if (e == Z_BUF_ERROR || e == Z_DATA_ERROR) zerr = path + ": " + (msg && *msg ? msg : "truncated gzip stream");
            return false;
        }
        end += (size_t)got;
        return true;
    }
    /* next line [*b, *e) without the terminator; false at end of file or on a zlib error (zerr set) */
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
                if (!zerr.empty()) return false;
                if (pos < end) {                 /* last line without '\n' */
                    *b = buf.data() + pos; *e = buf.data() + end; pos = end;
                    if (*e > *b && (*e)[-1] == '\r') (*e)--;
                    return true;
                }
                return false;
            }
        }
    }
    void produce()
    {
        if (inflate_threads > 1 && BgzfSource::probe(path.c_str())) bgzf.reset(new BgzfSource(path.c_str(), inflate_threads));
        else {
            gz = gzopen(path.c_str(), "rb");
            if (!gz) { finish("cannot open " + path); return; }
            gzbuffer(gz, 1u << 20);
        }
        buf.resize(4u << 20);
        Batch bt;
        std::string pending;            /* FASTA: header of the record that follows the one just finished */
        bool have_pending = false;
        const char *b, *e;
        for (;;) {
            if (fastq) {
                if (!line(&b, &e)) break;
                if (b == e) continue;                                   /* blank line between records */
                if (*b != '@') { if (bt.n && !push(std::move(bt))) return; finish(path + ": malformed FASTQ header"); return; }
                bt.names.insert(bt.names.end(), b + 1, e);
                bool ok = line(&b, &e);
                if (ok) {
                    bt.seq.insert(bt.seq.end(), b, e);
                    const char *pb, *pe;
                    ok = line(&pb, &pe) && pb != pe && *pb == '+' && line(&pb, &pe);
                }
                if (!ok) {
                    bt.seq.resize(bt.seq_end.empty() ? 0 : (size_t)bt.seq_end.back());          /* drop the partial record */
                    bt.names.resize(bt.name_end.empty() ? 0 : (size_t)bt.name_end.back());
                    if (bt.n && !push(std::move(bt))) return;
                    finish(!zerr.empty() ? zerr : path + ": truncated FASTQ record");
                    return;
                }
            } else {
                if (!have_pending) {
                    bool found = false;
                    while (line(&b, &e)) { if (b < e && *b == '>') { pending.assign(b + 1, e); found = true; break; } }
                    if (!found) break;
                }
                bt.add(pending);
                have_pending = false;
                while (line(&b, &e)) {
                    if (b < e && *b == '>') { pending.assign(b + 1, e); have_pending = true; break; }
                    bt.seq.insert(bt.seq.end(), b, e);
                }
                if (!zerr.empty()) break;
            }
            bt.close_record();
            if (bt.seq.size() >= BATCH_BYTES || bt.n >= BATCH_RECORDS) {
                if (!push(std::move(bt))) return;
                bt = Batch();
            }
        }
        if (!zerr.empty()) {
            bt.seq.resize(bt.seq_end.empty() ? 0 : (size_t)bt.seq_end.back());
            bt.names.resize(bt.name_end.empty() ? 0 : (size_t)bt.name_end.back());
        }
        if (bt.n && !push(std::move(bt))) return;
        finish(zerr);
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
    std::vector<std::unique_ptr<FileStream>> streams;   /* one slot per path; created up to `ahead` files early */
    size_t started = 0, ahead = 8;
    int inflate_threads = 1;        /* per file: cores / files that are open side by side (BGZF files only) */
    size_t budget_mb = 2048;        /* parsed records the file threads may hold in total: whole files of a typical run
                                       fit, so that the files really are inflated side by side */
    Batch bt;                       /* batch being handed out */
    int32_t bt_rec = 0;             /* its next record */
    char err[512] = "";

    Chunk cur, next;
    Chunk left;                     /* records read ahead with another nrec and not handed out yet */
    int32_t left_pos = 0;
    std::thread worker;
    bool prefetching = false;
    int32_t prefetch_nrec = 0;

    /* Up to `want` records (<= 0: no limit) of the current batch into c; refills the batch from the file list.
     * Returns the number of records copied, 0 at the end of all files, < 0 on error. */
    int take(Chunk &c, int32_t want)
    {
        while (bt_rec >= bt.n) {
            if (file_idx >= paths.size()) return 0;
            if (streams.size() < paths.size()) streams.resize(paths.size());
            for (; started < paths.size() && started < file_idx + ahead; started++)
                streams[started].reset(new FileStream(paths[started], fastq,
                                                      std::max<size_t>(4, (budget_mb << 20) / ahead / FileStream::BATCH_BYTES),
                                                      inflate_threads));
            std::string e;
            bt = Batch(); bt_rec = 0;
            if (!streams[file_idx]->pop(&bt, &e)) {
                if (!e.empty()) { snprintf(err, sizeof err, "%s", e.c_str()); return NTL_ERR_SEQUENCE; }
                streams[file_idx].reset();                                  /* end of this file: joins its thread */
                file_idx++;
            }
        }
        const int32_t k = want > 0 ? std::min(want, bt.n - bt_rec) : bt.n - bt_rec;
        const int64_t s0 = bt_rec ? bt.seq_end[bt_rec - 1] : 0, s1 = bt.seq_end[bt_rec + k - 1];
        const int64_t m0 = bt_rec ? bt.name_end[bt_rec - 1] : 0, m1 = bt.name_end[bt_rec + k - 1];
        const int64_t sb = (int64_t)c.seq.size() - s0, mb = (int64_t)c.names.size() - m0;
        c.seq.insert(c.seq.end(), bt.seq.begin() + s0, bt.seq.begin() + s1);
        c.names.insert(c.names.end(), bt.names.begin() + m0, bt.names.begin() + m1);
        for (int32_t i = 0; i < k; i++) {
            c.seq_off.push_back(bt.seq_end[bt_rec + i] + sb);
            c.name_off.push_back(bt.name_end[bt_rec + i] + mb);
        }
        c.n += k;
        bt_rec += k;
        return k;
    }
    /* records [from, from + k) of src appended to c */
    static void append(Chunk &c, const Chunk &src, int32_t from, int32_t k)
    {
        const int64_t s0 = src.seq_off[from], s1 = src.seq_off[from + k];
        const int64_t m0 = src.name_off[from], m1 = src.name_off[from + k];
        const int64_t sb = (int64_t)c.seq.size() - s0, mb = (int64_t)c.names.size() - m0;
        c.seq.insert(c.seq.end(), src.seq.begin() + s0, src.seq.begin() + s1);
        c.names.insert(c.names.end(), src.names.begin() + m0, src.names.begin() + m1);
        for (int32_t i = 1; i <= k; i++) {
            c.seq_off.push_back(src.seq_off[from + i] + sb);
            c.name_off.push_back(src.name_off[from + i] + mb);
        }
        c.n += k;
    }
    void take_left(Chunk &c, int32_t k) { append(c, left, left_pos, k); left_pos += k; }
    void read_chunk(Chunk &c, int32_t nrec)
    {
        c.clear();
        if (left_pos < left.n) {                         /* never drop parsed records: they come first */
            const int32_t avail = left.n - left_pos;
            take_left(c, nrec > 0 ? std::min(nrec, avail) : avail);
            if (left_pos >= left.n) {
                const int st = left.status;
                left.clear(); left_pos = 0;
                if (st < 0) { c.status = st; return; }   /* the read-ahead ended in an error: report it after its records */
            }
        }
        while (nrec <= 0 || c.n < nrec) {
            const int r = take(c, nrec > 0 ? nrec - c.n : 0);
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
    r->ahead = std::min<size_t>(16, std::max<size_t>(2, std::thread::hardware_concurrency()));
    if (const char *a = getenv("NTL_READER_FILES")) { const int v = atoi(a); if (v >= 1 && v <= 64) r->ahead = (size_t)v; }
    if (const char *a = getenv("NTL_READER_MB")) { const int v = atoi(a); if (v >= 16 && v <= (1 << 20)) r->budget_mb = (size_t)v; }
    {   /* cores left over when the list is shorter than the core count go to the inside of BGZF files */
        const size_t cores = std::max<size_t>(1, std::thread::hardware_concurrency());
        const size_t side_by_side = std::min<size_t>(r->ahead, (size_t)n_paths);
        /* two cores stay with the line parser and the consumer of the chunks */
        r->inflate_threads = (int)std::min<size_t>(32, std::max<size_t>(1, (cores > 3 ? cores - 2 : cores) / side_by_side));
        if (const char *a = getenv("NTL_READER_BGZF_THREADS")) { const int v = atoi(a); if (v >= 1 && v <= 64) r->inflate_threads = v; }
    }
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
        if (r->prefetching) {
            /* nrec changed (NanoTel.R never does that, the ABI allows it): nothing that was parsed is dropped.  The
             * chunk read ahead with the old size comes first (it was filled from `left`, then from the files), then
             * whatever `left` still holds; read_chunk splits or tops that up to the new size. */
            r->worker.join();
            r->prefetching = false;
            Chunk m; m.clear();
            if (r->next.n > 0) ntl_reader::append(m, r->next, 0, r->next.n);
            if (r->left_pos < r->left.n) ntl_reader::append(m, r->left, r->left_pos, r->left.n - r->left_pos);
            m.status = r->next.status < 0 ? r->next.status : r->left.status;
            std::swap(r->left, m);
            r->left_pos = 0;
            r->next.clear();
        }
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
