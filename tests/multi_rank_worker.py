"""Worker of tests/test_gpu_multi.py::test_one_rank_per_process_groups: one rank of a two-process device group.
usage: multi_rank_worker.py <dir> <rank> <world>"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    d, rank, world = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    from ldweaver_b200 import api
    idf = os.path.join(d, "nccl_id.bin")
    if rank == 0:
        uid = api.DeviceGroup.unique_id()
        with open(idf + ".tmp", "wb") as fh:
            fh.write(uid)
        os.replace(idf + ".tmp", idf)
    else:
        t0 = time.time()
        while not os.path.exists(idf):
            if time.time() - t0 > 120:
                raise TimeoutError("no NCCL id from rank 0")
            time.sleep(0.05)
        uid = open(idf, "rb").read()
    inp = np.load(os.path.join(d, "in.npz"))
    grp = api.DeviceGroup.from_rank(rank, rank, world, uid)
    n, S = inp["codes"].shape
    grp.load_codes(inp["codes"] if rank == 0 else None, n, S)   # only rank 0 holds the matrix on the host
    w, cnt, sharded = grp.hdw(0.1, force_shard=True, return_parts=True)
    assert sharded
    sr, lr, bd, thr, prob, st = grp.mi_scan(inp["hdw"], inp["POS"], inp["paint"], 1000, float(inp["g"]), 20000.0, 4000.0, float(inp["lra"]))
    out = {"hdw": w, "cnt": cnt, "thr": thr}
    for name, t in (("sr", sr), ("lr", lr)):
        for k, v in t.items():
            out[f"{name}_{k}"] = v
    np.savez(os.path.join(d, f"out{rank}.npz"), **out)
    print(json.dumps({"world": grp.world, "n_local": grp.n_local, "first_rank": grp.first_rank, "n_sr": int(len(sr["MI"]))}))
    grp.close()


if __name__ == "__main__":
    main()
