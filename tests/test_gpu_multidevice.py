"""GPU: one context that shards every batch over several devices (ntl_params.device_ids) -- what replaces the 8 forked
workers of NanoTel.R:2207, :2234-2254.  The shards are contiguous and balanced by bases, every shard runs the whole
pipeline on its own device / stream / packer threads, records and window tables are gathered in input order: the
results must be byte-identical to a single-device context, whatever the number of shards.  On a box with one GPU the
shards all live on device 0 (a device may be listed more than once), on a box with several they use them all."""
import numpy as np
import pytest

from oracle.compare import full_parity

pytestmark = pytest.mark.gpu


def _devices(k):
    import torch
    n = torch.cuda.device_count()
    return [i % n for i in range(k)]


@pytest.fixture(scope="module")
def reads():
    from nanotel_b200.synth import synth_reads
    return synth_reads(6000, 20261018 + 9, telomeric_frac=0.3, median_len=7000.0, n_frac=0.02)


@pytest.mark.parametrize("k", [2, 3, 8])
def test_sharded_context_equals_single_device(reads, k):
    from nanotel_b200 import Scanner
    buf, off, meta = reads
    cfg = dict(patterns="YYAGGG", tvr_patterns="TTGGG CCAGGG TCAGGG", rc=True, debug_stages=True)
    with Scanner(**cfg) as one:
        ref = one.scan_concat(buf, off)
        ref_counts = [one.window_counts(t) for t in range(3)]
        ref_win = {i: one.windows(i, 1) for i in (0, 17, 2999, 3000, 5999)}
        ref_stage = one.stages(4242, 2)
    with Scanner(devices=_devices(k), **cfg) as sc:
        assert sc.n_devices == k
        res = sc.scan_concat(buf, off)
        bounds, devs = sc.shards()
        assert bounds[0] == 0 and bounds[-1] == len(res) and np.all(np.diff(bounds) > 0)
        # balanced by bases: every shard within one longest read of the mean
        per = np.add.reduceat(meta["lengths"], bounds[:-1])
        assert per.max() - per.min() <= 2 * meta["lengths"].max()
        a, b = ref.copy(), res.copy()
        a["win_offset"] = 0; b["win_offset"] = 0              # internal: position inside the owning device's planes
        assert a.tobytes() == b.tobytes()
        for t in range(3):
            assert np.array_equal(sc.window_counts(t), ref_counts[t])
        for i, tab in ref_win.items():
            got = sc.windows(i, 1)
            assert all(np.array_equal(x, y) for x, y in zip(got, tab))
        assert sc.stages(4242, 2) == ref_stage
        # the staged entry points run on every shard as well
        sc.pack_concat(buf, off); sc.upload(); sc.run()
        again = sc.download().copy()
        again["win_offset"] = 0
        assert again.tobytes() == b.tobytes()
        tm = sc.timings()
        assert tm["bases"] == meta["bases"] and tm["kernel_launches"] >= 3 * k


def test_sharded_filter_run_matches_the_oracle(reads):
    from nanotel_b200 import Scanner
    buf, off, meta = reads
    with Scanner("TTAGGG", None, 0.6, 200, use_filter=True, right_edge=True, devices=_devices(4)) as sc:
        res = sc.scan_concat(buf, off)
        v = full_parity(sc, res, (buf, off), "TTAGGG", None, 0.6, 200, True, False, True)
    assert v["mismatches"] == 0 and v["reads_kept"] > 100, v


def test_more_shards_than_reads():
    from nanotel_b200 import Scanner
    seqs = [b"TTAGGG" * 300, b"ACGT" * 500]
    with Scanner("TTAGGG", devices=_devices(5)) as sc:
        res = sc.scan(seqs)
        assert res[0]["status"] & 1 and not (res[1]["status"] & 1)
        assert sc.windows(0, 0)[2].sum() > 1700
        assert len(sc.scan([])) == 0
