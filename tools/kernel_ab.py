#!/usr/bin/env python
"""Developer tool: same-box A/B timing of scan-kernel variants (tools/build_variant.sh).

    python tools/kernel_ab.py [--nsnp 30000] [--steps 5] name1 name2 ...     (names under ldweaver_b200/variants/, or 'base')

Every variant runs in its own process on the same synthetic C2-shaped data (616 sequences); prints per-variant kernel
time per scan, pairs/s and a digest of the results (link counts, MI sums, threshold) that must agree between variants."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]


def worker(args):
    import numpy as np
    import torch
    from ldweaver_b200 import api, synth
    d = np.load(args.data)
    codes, POS, paint, g = d["codes"], d["POS"], d["paint"], int(d["g"])
    n, S = codes.shape
    grp = api.DeviceGroup.from_rank(0, 0, 1, None)
    grp.load_codes(torch.from_numpy(codes).pin_memory().numpy(), n, S)
    hdw = grp.hdw(0.1)
    lra = synth.exact_lr_links_approx(POS, g, 20000.0)
    blk = api.round_half_even_thousands(args.blk)
    lr_retain = 1e6 * (n / 100000.0) ** 2

    def scan(flags):
        return grp.mi_scan(hdw, POS, paint, blk, float(g), 20000.0, lr_retain, lra, flags, copy=False)
    for _ in range(3):
        scan(api.SCAN_NO_D2H)
    ks, ds, pairs = [], [], 0
    for _ in range(args.steps):
        *_, st = scan(api.SCAN_NO_D2H)
        st = st[0]
        ks.append(st["t_kernel_ms"]); ds.append(st["t_scan_ms"] + st["t_select_ms"]); pairs = st["n_pairs"]
    sr, lr, bd, thr, prob, st = scan(0)
    srv, lrv = sr.views(), lr.views()
    out = {"variant": args.worker, "kernel_ms": min(ks), "kernel_ms_all": ks, "device_ms": min(ds), "Gpairs_s": pairs / min(ds) / 1e6,
           "n_blocks": st[0]["n_blocks"], "n_scan_launches": st[0]["n_scan_launches"], "n_tiles": st[0]["n_tiles"],
           "digest": {"n_sr": int(sr.n), "n_lr": int(lr.n), "sr_MI_sum": float(np.sum(srv["MI"], dtype=np.float64)) if sr.n else 0.0,
                      "lr_MI_sum": float(np.sum(lrv["MI"], dtype=np.float64)) if lr.n else 0.0,
                      "lr_pos_sum": int(np.sum(lrv["pos1"].astype(np.int64) * 3 + lrv["pos2"])) if lr.n else 0,
                      "thr_sum": float(np.nansum(thr)), "eps": st[0]["eps_obs_max"], "reruns": st[0]["n_reruns"]}}
    print("AB " + json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("names", nargs="*")
    ap.add_argument("--nsnp", type=int, default=30000)
    ap.add_argument("--blk", type=int, default=10000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--worker", default=None)
    ap.add_argument("--data", default="/tmp/kernel_ab_data.npz")
    args = ap.parse_args()
    if args.worker:
        return worker(args)
    import numpy as np
    from ldweaver_b200 import synth
    if not os.path.exists(args.data):
        sy = synth.generate(616, args.nsnp, 616100)
        np.savez(args.data, codes=sy.codes, POS=sy.POS, paint=sy.paint, g=sy.g)
    res = []
    for nm in args.names:
        env = dict(os.environ)
        if nm != "base":
            env["LDW_LIBRARY_PATH"] = os.path.join(ROOT, "ldweaver_b200", "variants", f"libldwgpu_{nm}.so")
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", nm, "--data", args.data, "--steps", str(args.steps),
                            "--blk", str(args.blk)], env=env, capture_output=True, text=True)
        lines = [l for l in p.stdout.splitlines() if l.startswith("AB ")]
        if p.returncode != 0 or not lines:
            print(f"variant {nm}: FAILED rc={p.returncode}\n{p.stdout[-1500:]}\n{p.stderr[-3000:]}")
            continue
        r = json.loads(lines[0][3:])
        res.append(r)
        print(f"{nm:12s} kernel {r['kernel_ms']:8.3f} ms  device {r['device_ms']:8.3f} ms  {r['Gpairs_s']:7.2f} G pairs/s  launches {r['n_scan_launches']}  {json.dumps(r['digest'])}")
        if p.stderr.strip():
            print(p.stderr[-2500:])
    if res:
        d0 = res[0]["digest"]
        for r in res[1:]:
            same = all(r["digest"][k] == d0[k] for k in ("n_sr", "n_lr", "lr_pos_sum", "thr_sum"))
            close = abs(r["digest"]["sr_MI_sum"] - d0["sr_MI_sum"]) <= 1e-6 * max(1.0, abs(d0["sr_MI_sum"]))
            print(f"{r['variant']:12s} vs {res[0]['variant']}: links/thresholds identical={same} sr_MI_sum close={close} speed x{res[0]['kernel_ms'] / r['kernel_ms']:.3f}")


if __name__ == "__main__":
    main()
