"""CPU, world_size 2, gloo: the multi-rank host logic -- weights broadcast from rank 0, make_blocks rows dealt
round-robin, per-rank link tables gathered on the host -- reproduces the single-rank result exactly.
(The block computation itself is done by the oracle here; on the GPU box each rank calls ldw_mi_scan(n_parts, part).)"""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import c_oracle as CO
    import ldw_oracle as O
    from ldweaver_b200 import api
    e = dict(np.load(os.path.join(ROOT, "tests", "golden", "fixture_expected.npz")))
    snp = O.snp_dat_from_codes(e["codes"], e["relaxed_POS"], 50000)
    # weights: computed on rank 0 only, broadcast (NCCL over NVLink on the GPU box, gloo here)
    w = torch.zeros(snp.nseq, dtype=torch.float64)
    if rank == 0:
        w.copy_(torch.from_numpy(CO.hdw(snp.codes, 0.1)[0]))
    dist.broadcast(w, src=0)
    hdw = w.numpy()
    assert np.array_equal(hdw, e["hdw"])
    blocks = O.make_blocks(snp.nsnp, 1000)
    assert blocks == api.make_blocks(snp.nsnp, 1000)
    mine = api.partition_blocks(snp.nsnp, 1000, world, rank)
    POS = snp.POS.astype(np.float64)
    rows = []
    for b in mine:
        fs, fe, ts, te = blocks[b]
        f, t = np.arange(fs - 1, fe), np.arange(ts - 1, te)
        MI = CO.block_mi(snp.codes, hdw, snp.r, snp.uqe, f, t)
        L = CO.block_links(MI, POS, f, t, 50000.0, 20000.0, 1e4, 1e5)
        rows.append((b, POS[t[L["col"]]][L["is_sr"]], POS[f[L["row"]]][L["is_sr"]], L["MI"][L["is_sr"]], L["thr"]))
    gathered = [None] * world
    dist.all_gather_object(gathered, rows)
    if rank == 0:
        allrows = sorted([r for part in gathered for r in part], key=lambda r: r[0])
        assert [r[0] for r in allrows] == list(range(len(blocks)))  # every block exactly once
        p1 = np.concatenate([r[1] for r in allrows]).astype(np.int32)
        p2 = np.concatenate([r[2] for r in allrows]).astype(np.int32)
        mi = np.concatenate([r[3] for r in allrows])
        np.save(os.path.join(out_dir, "p1.npy"), p1)
        np.save(os.path.join(out_dir, "p2.npy"), p2)
        np.save(os.path.join(out_dir, "mi.npy"), mi)
        np.save(os.path.join(out_dir, "thr.npy"), np.array([r[4] for r in allrows]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_reproduces_single_rank(tmp_path, fixture_expected):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    e = fixture_expected
    np.testing.assert_array_equal(np.load(tmp_path / "p1.npy"), e["g50k_b1000_sr_pos1"])
    np.testing.assert_array_equal(np.load(tmp_path / "p2.npy"), e["g50k_b1000_sr_pos2"])
    assert np.abs(np.load(tmp_path / "mi.npy") - e["g50k_b1000_sr_MI"]).max() < 1e-12
    thr = np.load(tmp_path / "thr.npy")
    ok = ~np.isnan(e["g50k_b1000_thr"])
    assert np.abs(thr[ok] - e["g50k_b1000_thr"][ok]).max() < 1e-12


def test_partition_rule():
    from ldweaver_b200 import api
    for nsnp, blk in ((900, 1000), (2500, 1000), (100000, 10000), (300000, 10000), (31234, 10000)):
        blocks = api.make_blocks(nsnp, blk)
        cost = [(fe - fs + 1) * (fe - fs) // 2 if fs == ts else (fe - fs + 1) * (te - ts + 1) for fs, fe, ts, te in blocks]
        for w in (1, 2, 4, 8):
            parts = [api.partition_blocks(nsnp, blk, w, r) for r in range(w)]
            assert sorted(b for p in parts for b in p) == list(range(len(blocks)))     # every block exactly once
            assert all(p == sorted(p) for p in parts)                                   # make_blocks order within a part
            loads = [sum(cost[b] for b in p) for p in parts]
            if len(blocks) >= 4 * w:                                                    # enough blocks to balance
                assert max(loads) <= 1.08 * sum(loads) / w
    # C2 on 8 GPUs: 45 full + 10 half blocks -> at most 6.5 block units per rank (round-robin: 7)
    loads = [sum(1.0 if blocks_[0] != blocks_[2] else 0.5 for blocks_ in [api.make_blocks(100000, 10000)[b] for b in
             api.partition_blocks(100000, 10000, 8, r)]) for r in range(8)]
    assert max(loads) <= 6.5
