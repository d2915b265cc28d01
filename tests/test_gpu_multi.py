"""GPU tests of the device-group layer (ldw_group_*): multi-GPU hdw and scan against the single-device path.

The one-member group runs on any GPU box; the others need >= 2 devices (`gpurun --gpus 2`)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_devices():
    import torch
    return torch.cuda.device_count()


def _case(nseq=300, nsnp=5200, seed=3):
    from ldweaver_b200 import synth
    sy = synth.generate(nseq=nseq, nsnp=nsnp, seed=seed)
    return sy


def _single(sy, blk, retain, flags=0):
    import ldweaver_b200 as ldw
    snp = ldw.snp_dat_from_codes(sy.codes, sy.POS, sy.g)
    hdw, cnt, _ = ldw.estimate_Hamming_distance_weights(snp, 0.1, return_parts=True)
    plan = ldw.MIPlan(snp, hdw, sy.paint, blk)
    from ldweaver_b200 import synth
    lra = synth.exact_lr_links_approx(sy.POS, sy.g, 20000.0)
    out = plan.scan(float(sy.g), 20000.0, retain, lra, flags)
    plan.close()
    return hdw, cnt, out, lra


def _same_tables(a, b):
    for k in ("pos1", "pos2", "clust1", "clust2", "len", "block", "MI"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


def test_one_member_group_equals_the_plain_entry_points():
    """world = 1: no NCCL call is made, but the whole group path (resident matrix, plan from device codes, shared
    short-range layout, merge) runs -- and must reproduce ldw_hdw / ldw_mi_scan bit for bit."""
    from ldweaver_b200 import api
    sy = _case()
    hdw, cnt, (sr, lr, bd, thr, prob, st), lra = _single(sy, 1000, 3000.0)
    grp = api.DeviceGroup([0])
    assert (grp.world, grp.n_local, grp.first_rank) == (1, 1, 0)
    grp.load_codes(sy.codes)
    w, c, sharded = grp.hdw(0.1, return_parts=True)
    np.testing.assert_array_equal(w, hdw)
    np.testing.assert_array_equal(c, cnt)
    assert not sharded
    gsr, glr, gbd, gthr, gprob, gst = grp.mi_scan(hdw, sy.POS, sy.paint, 1000, float(sy.g), 20000.0, 3000.0, lra)
    _same_tables(gsr, sr)
    _same_tables(glr, lr)
    _same_tables(gbd, bd)
    np.testing.assert_array_equal(gthr, thr)
    np.testing.assert_array_equal(gprob, prob)
    assert len(gst) == 1 and gst[0]["n_pairs"] == st["n_pairs"]
    # SR-only mode and the in-scan fp64 short-range MI through the group
    _, _, (sr2, lr2, _, _, _, _), _ = _single(sy, 1000, 3000.0, flags=api.SCAN_SR_ONLY | api.SCAN_SR_EXACT)
    gsr2, glr2, *_ = grp.mi_scan(hdw, sy.POS, sy.paint, 1000, float(sy.g), 20000.0, 3000.0, lra, flags=api.SCAN_SR_ONLY | api.SCAN_SR_EXACT)
    _same_tables(gsr2, sr2)
    assert len(glr2["MI"]) == 0 == len(lr2["MI"])
    grp.close()


@pytest.mark.skipif("n_devices() < 2")
def test_in_process_group_over_all_devices_is_bitwise_the_single_device_result():
    from ldweaver_b200 import api
    nd = min(n_devices(), 4)
    sy = _case(nseq=700, nsnp=9300, seed=5)
    hdw, cnt, (sr, lr, bd, thr, prob, st), lra = _single(sy, 1000, 5000.0)
    grp = api.DeviceGroup(list(range(nd)))
    assert (grp.world, grp.n_local) == (nd, nd)
    grp.load_codes(sy.codes)
    for force in (False, True):  # small problem: replicated unless forced; forced = tiles dealt + ncclAllReduce
        w, c, sharded = grp.hdw(0.1, force_shard=force, return_parts=True)
        assert sharded == force
        np.testing.assert_array_equal(c, cnt)
        np.testing.assert_array_equal(w, hdw)
    gsr, glr, gbd, gthr, gprob, gst = grp.mi_scan(hdw, sy.POS, sy.paint, 1000, float(sy.g), 20000.0, 5000.0, lra)
    _same_tables(gsr, sr)
    _same_tables(glr, lr)
    _same_tables(gbd, bd)
    np.testing.assert_array_equal(gthr, thr)
    np.testing.assert_array_equal(gprob, prob)
    assert len(gst) == nd and sum(d["n_pairs"] for d in gst) == st["n_pairs"]
    assert all(d["n_blocks"] > 0 for d in gst)
    grp.close()


@pytest.mark.skipif("n_devices() < 2")
def test_public_api_with_devices_matches_single_device(tmp_path):
    """perform_MI_computation(devices=[0, 1]) / estimate_Hamming_distance_weights(devices=...): same return value, same
    TSV bytes as the single-device call with the in-scan fp64 short-range MI."""
    import ldweaver_b200 as ldw
    from ldweaver_b200 import synth
    sy = _case(nseq=400, nsnp=6100, seed=8)
    snp = ldw.snp_dat_from_codes(sy.codes, sy.POS, sy.g)
    lra = synth.exact_lr_links_approx(sy.POS, sy.g, 20000.0)
    hdw1 = ldw.estimate_Hamming_distance_weights(snp, 0.1)
    hdw2 = ldw.estimate_Hamming_distance_weights(snp, 0.1, devices=[0, 1])
    np.testing.assert_array_equal(hdw1, hdw2)
    outs = []
    for tag, kw in (("one", dict(exact_sr="in_scan")), ("two", dict(devices=[0, 1]))):
        d = tmp_path / tag
        d.mkdir()
        res = ldw.perform_MI_computation(snp, hdw1, ldw.CdsVar(sy.paint, 3), lr_save_path=str(d / "lr.tsv"), sr_save_path=str(d / "sr.tsv"),
                                         plt_folder=str(d), lr_retain_links=4000, max_blk_sz=1000, lr_links_approx=lra, **kw)
        outs.append((res, (d / "lr.tsv").read_bytes(), (d / "sr.tsv").read_bytes()))
    (a, lra_bytes, sra_bytes), (b, lrb_bytes, srb_bytes) = outs
    _same_tables(a.sr, b.sr)
    _same_tables(a.lr, b.lr)
    assert lra_bytes == lrb_bytes and sra_bytes == srb_bytes
    for k in a.sr_links_red:
        np.testing.assert_array_equal(a.sr_links_red[k], b.sr_links_red[k])


@pytest.mark.skipif("n_devices() < 2")
def test_one_rank_per_process_groups(tmp_path):
    """The torchrun form: two processes, one rank each, NCCL id through a file.  Rank 0 alone holds the matrix; the union of
    the ranks' tables is the single-device table."""
    sy = _case(nseq=500, nsnp=7400, seed=11)
    hdw, cnt, (sr, lr, bd, thr, prob, st), lra = _single(sy, 1000, 4000.0)
    np.savez(tmp_path / "in.npz", codes=sy.codes, POS=sy.POS, paint=sy.paint, g=sy.g, hdw=hdw, lra=lra)
    worker = os.path.join(ROOT, "tests", "multi_rank_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(tmp_path), str(r), "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-3000:]
    parts = [dict(np.load(tmp_path / f"out{r}.npz")) for r in range(2)]
    for r in range(2):
        np.testing.assert_array_equal(parts[r]["hdw"], hdw)     # forced shard + all-reduce: every rank has all weights
        np.testing.assert_array_equal(parts[r]["cnt"], cnt)
    for name, ref in (("sr", sr), ("lr", lr)):
        blk = np.concatenate([parts[r][f"{name}_block"] for r in range(2)])
        order = np.argsort(blk, kind="stable")
        for k in ("pos1", "pos2", "clust1", "clust2", "len", "block", "MI"):
            got = np.concatenate([parts[r][f"{name}_{k}"] for r in range(2)])[order]
            np.testing.assert_array_equal(got, ref[k], err_msg=f"{name}.{k}")
    assert set(parts[0]["sr_block"]).isdisjoint(set(parts[1]["sr_block"]))
    merged_thr = np.where(np.isnan(parts[0]["thr"]), parts[1]["thr"], parts[0]["thr"])
    np.testing.assert_array_equal(merged_thr, thr)
    meta = [json.loads(o.strip().splitlines()[-1]) for o, _ in outs]
    assert all(m["world"] == 2 and m["n_local"] == 1 for m in meta) and {m["first_rank"] for m in meta} == {0, 1}
