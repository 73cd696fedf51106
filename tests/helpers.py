"""Shared comparison code: CUDA path (through the C ABI) vs the CPU oracle (oracle/)."""
import numpy as np

from oracle import oracle as O


def oracle_batch(seqs, patterns, tvr=None, min_density=0.6, S=100, right_edge=False, rc=False, use_filter=False,
                 n_threads=4):
    P = O.make_params(patterns, tvr, min_density, S, right_edge)
    recs, passed, win_off, wc = O.scan_batch(P, seqs, do_rc=rc, use_filter=use_filter, n_threads=n_threads)
    return P, recs, passed, win_off, wc


def compare_batch(sc, res, seqs, recs, passed, win_off, wc, check_windows=True, check_stages=False, label=""):
    """Bit-exact comparison of every per-read record, window count and (optionally) intermediate stage."""
    from nanotel_b200 import READ_FILTERED, READ_KEEP, READ_REF_ERROR
    n = len(seqs)
    assert len(res) == n
    T = sc.n_tracks
    bad = []
    for i in range(n):
        r, o = res[i], recs[i]
        if not passed[i]:
            if not (r["status"] & READ_FILTERED):
                bad.append((i, "filter verdict", int(r["status"]), "oracle dropped it"))
            continue
        if r["status"] & READ_FILTERED:
            bad.append((i, "filter verdict", int(r["status"]), "oracle kept it"))
            continue
        o_err = bool(o["flags"] & O.FLAG_REF_ERROR)
        if bool(r["status"] & READ_REF_ERROR) != o_err:
            bad.append((i, "ref_error", int(r["status"]), int(o["flags"])))
            continue
        if o_err:
            continue
        if int(r["n_win"]) != int(o["n_win"]):
            bad.append((i, "n_win", int(r["n_win"]), int(o["n_win"])))
        if bool(r["status"] & READ_KEEP) != bool(o["keep"]):
            bad.append((i, "keep", int(r["status"]), int(o["keep"])))
        for t in range(T):
            g, e = r["track"][t], o["t"][t]
            if int(g["start"]) != int(e["start"]) or int(g["end"]) != int(e["end"]):
                bad.append((i, "track%d interval" % t, (int(g["start"]), int(g["end"])), (int(e["start"]), int(e["end"]))))
            elif float(g["density"]).hex() != float(e["density"]).hex():
                bad.append((i, "track%d density" % t, float(g["density"]), float(e["density"])))
            if check_stages:
                s = sc.stages(i, t)
                for k in ("coarse_start", "coarse_end", "acc_start", "acc_end", "edge_start", "edge_end"):
                    if int(s[k]) != int(e[k]):
                        bad.append((i, "track%d %s" % (t, k), int(s[k]), int(e[k])))
                if float(s["acc_density"]).hex() != float(e["acc_density"]).hex():
                    bad.append((i, "track%d acc_density" % t, s["acc_density"], float(e["acc_density"])))
        if check_windows and int(o["n_win"]) > 0:
            nw = int(o["n_win"])
            base = int(win_off[i])
            for t in range(T):
                st, en, cov, den = sc.windows(i, t, nw)
                exp = wc[base + t * nw: base + (t + 1) * nw]
                if not np.array_equal(cov, exp):
                    k = int(np.nonzero(cov != exp)[0][0])
                    bad.append((i, "track%d window %d count" % (t, k), int(cov[k]), int(exp[k])))
                est, een = O.split_telo(len(seqs[i]), sc.params.subseq_length)
                if not (np.array_equal(st, est) and np.array_equal(en, een)):
                    bad.append((i, "window geometry", None, None))
                width = (en - st + 1).astype(np.float64)
                if not np.array_equal(den, cov.astype(np.float64) / width):
                    bad.append((i, "track%d window densities" % t, None, None))
    assert not bad, "%s: %d mismatches, first: %s (read length %d)" % (label, len(bad), bad[:5], len(seqs[bad[0][0]]))
