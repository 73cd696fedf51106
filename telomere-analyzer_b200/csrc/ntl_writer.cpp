/*
 * ntl_writer.cpp -- host side of the per-read outputs of analyze_read (NanoTel.R:1870-1918), native and threaded:
 *   reads/<Serial>.fasta.gz            writeXStringSet(..., compress = TRUE) of the read (NanoTel.R:1870-1873)
 *   density_vectors/read<Serial>.csv   the per-window tables (`subs` data frames, NanoTel.R:740-765) the plot functions
 *                                      receive (NanoTel.R:1876-1918): ID, start_index, end_index, then density and
 *                                      class per track
 * for every summary row of a chunk.  The command-line driver spent most of its wall clock formatting these files in
 * Python; here one call writes them with all host threads (zlib level 6 = R's gzfile() default; doubles as the
 * shortest decimal string that reads back to the same value, integral values without a decimal point: the form
 * readr::write_csv prints).
 * Plot rendering stays with the reference's own R functions (R/plot_density_vectors.R).
 */
#include "../../include/nanotel_b200.h"
#include "ntl_pack.h"

#include <sys/stat.h>
#include <zlib.h>

#include <atomic>
#include <charconv>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct CompTable {
    unsigned char comp[256], up[256];
    CompTable()
    {
        for (int i = 0; i < 256; i++) {
            up[i] = (unsigned char)((i >= 'a' && i <= 'z') ? i - 32 : i);
            comp[i] = up[i];
        }
        /* Biostrings::reverseComplement, IUPAC aware */
        const char *a = "ACGTMRWSYKVHDBN", *b = "TGCAKYWSRMBDHVN";
        for (int k = 0; a[k]; k++) {
            comp[(unsigned char)a[k]] = (unsigned char)b[k];
            comp[(unsigned char)(a[k] + 32)] = (unsigned char)b[k];
        }
    }
};
const CompTable g_tab;

/* '>' + header line, then the upper-case sequence 80 letters per line (writeXStringSet's layout) */
void fasta_text(const char *name, int64_t name_len, const char *seq, int64_t L, bool rc, std::string &out)
{
    out.clear();
    out.reserve((size_t)(L + L / 80 + name_len + 8));
    out.push_back('>');
    out.append(name, (size_t)name_len);
    out.push_back('\n');
    for (int64_t k = 0; k < L; k += 80) {
        const int64_t e = k + 80 < L ? k + 80 : L;
        if (rc) for (int64_t i = k; i < e; i++) out.push_back((char)g_tab.comp[(unsigned char)seq[L - 1 - i]]);
        else for (int64_t i = k; i < e; i++) out.push_back((char)g_tab.up[(unsigned char)seq[i]]);
        out.push_back('\n');
    }
    if (L == 0) out.push_back('\n');                    /* b"\n".join([]) + b"\n" */
}

bool gzip_to_file(const std::string &text, const char *path, int level, std::vector<unsigned char> &scratch)
{
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    scratch.resize(deflateBound(&zs, (uLong)text.size()) + 64);
    zs.next_in = (Bytef *)text.data(); zs.avail_in = (uInt)text.size();
    zs.next_out = scratch.data(); zs.avail_out = (uInt)scratch.size();
    const int r = deflate(&zs, Z_FINISH);
    const size_t n = scratch.size() - zs.avail_out;
    deflateEnd(&zs);
    if (r != Z_STREAM_END) return false;
    FILE *f = fopen(path, "wb");
    if (!f) return false;
    const bool ok = fwrite(scratch.data(), 1, n, f) == n;
    return fclose(f) == 0 && ok;
}

/* the shortest decimal string that round-trips, integral values without a decimal point: readr::write_csv's doubles */
void append_double(std::string &s, double x)
{
    char b[40];
    auto r = std::to_chars(b, b + sizeof b, x);
    s.append(b, (size_t)(r.ptr - b));
}
void append_int(std::string &s, long long v)
{
    char b[24];
    auto r = std::to_chars(b, b + sizeof b, v);
    s.append(b, (size_t)(r.ptr - b));
}

bool write_text(const std::string &s, const char *path)
{
    FILE *f = fopen(path, "wb");
    if (!f) return false;
    const bool ok = fwrite(s.data(), 1, s.size(), f) == s.size();
    return fclose(f) == 0 && ok;
}

} // namespace

extern "C" int ntl_write_fasta_gz(const char *path, const char *name, const char *seq, int64_t len, int32_t rc)
{
    if (!path || !name || (!seq && len > 0) || len < 0) return NTL_ERR_ARG;
    std::string text;
    std::vector<unsigned char> scratch;
    fasta_text(name, (int64_t)strlen(name), seq, len, rc != 0, text);
    return gzip_to_file(text, path, 6, scratch) ? NTL_OK : NTL_ERR_IO;
}

extern "C" int ntl_write_read_outputs(const ntl_ctx *ctx, const char *out_dir, const char *buf, const int64_t *offsets,
                                      const char *names, const int64_t *name_off, const int32_t *serial,
                                      const int32_t *order, int32_t n_rows, int32_t n_tracks, double min_density,
                                      int32_t rc_applied, int32_t threads)
{
    if (!ctx || !out_dir || !buf || !offsets || !names || !name_off || !serial || (!order && n_rows > 0) || n_rows < 0 ||
        n_tracks < 1 || n_tracks > 3)
        return NTL_ERR_ARG;
    const std::string rd = std::string(out_dir) + "/reads", dv = std::string(out_dir) + "/density_vectors";
    mkdir(out_dir, 0777); mkdir(rd.c_str(), 0777); mkdir(dv.c_str(), 0777);
    static const char *sfx[3] = {"", "_mismatch", "_mismatch_tvr"};
    std::string header = "ID,start_index,end_index";
    for (int t = 0; t < n_tracks; t++) { header += ",density"; header += sfx[t]; header += ",class"; header += sfx[t]; }
    header += "\n";
    std::atomic<int> failed(0);
    if (threads < 1) threads = 1;
    ntl_parallel_for(n_rows, threads, 1, [&](int64_t b, int64_t e) {
        std::string text, csv;
        std::vector<unsigned char> scratch;
        std::vector<int32_t> st, en;
        std::vector<double> den[3];
        char path[4096];
        for (int64_t j = b; j < e; j++) {
            const int32_t i = order[j], s = serial[i];
            const int64_t L = offsets[i + 1] - offsets[i];
            fasta_text(names + name_off[i], name_off[i + 1] - name_off[i], buf + offsets[i], L, rc_applied != 0, text);
            snprintf(path, sizeof path, "%s/%d.fasta.gz", rd.c_str(), s);
            if (!gzip_to_file(text, path, 6, scratch)) { failed.store(1); continue; }
            /* the window tables of a summary row belong to a kept read: served from the host copy, no device call */
            int n = ntl_get_windows(ctx, i, 0, 0, nullptr, nullptr, nullptr, nullptr);
            if (n < 0) { failed.store(1); continue; }
            st.resize((size_t)n + 1); en.resize((size_t)n + 1);
            bool ok = true;
            for (int t = 0; t < n_tracks; t++) {
                den[t].resize((size_t)n + 1);
                if (ntl_get_windows(ctx, i, t, n, t == 0 ? st.data() : nullptr, t == 0 ? en.data() : nullptr, nullptr,
                                    den[t].data()) != n) ok = false;
            }
            if (!ok) { failed.store(1); continue; }
            csv.assign(header);
            for (int k = 0; k < n; k++) {
                append_int(csv, k + 1); csv.push_back(',');
                append_int(csv, st[(size_t)k]); csv.push_back(',');
                append_int(csv, en[(size_t)k]);
                for (int t = 0; t < n_tracks; t++) {
                    const double d = den[t][(size_t)k];
                    csv.push_back(',');
                    append_double(csv, d);
                    csv.push_back(',');
                    append_int(csv, d < min_density ? (d < 0.1 ? 0 : 1) : -5);       /* NanoTel.R:749-758 */
                }
                csv.push_back('\n');
            }
            snprintf(path, sizeof path, "%s/read%d.csv", dv.c_str(), s);
            if (!write_text(csv, path)) failed.store(1);
        }
    });
    return failed.load() ? NTL_ERR_IO : NTL_OK;
}
