"""Deterministic synthetic alignments shaped like the BASELINE.json configurations.

Common generator: genome length g = 2 221 315; POS = sorted sample without replacement; F = max(8, S/16) founder
haplotypes; per site an allele set (bi/tri/tetra-allelic) and allele frequencies with MAF ~ U(0.02, 0.5) (minor
alleles split evenly); each sequence copies a founder (Zipf-ish cluster sizes) and re-draws each site with
probability 0.03; N/gap injected i.i.d. per cell.  Sites that fail the reference's default filter are re-drawn, so
every returned site would be retained by parse_fasta_alignment(method="default")."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

G_DEFAULT = 2221315

CONFIGS = {
    # name: (nseq, nsnp, seed, allele-set probabilities (2,3,4 alleles), N rate)
    "C2": (616, 100_000, 616100, (0.847, 0.147, 0.006), 0.01),
    "C3": (10_000, 50_000, 1000050, (0.847, 0.147, 0.006), 0.01),
    "C4": (5_000, 300_000, 5000300, (0.847, 0.147, 0.006), 0.01),
    "C5": (2_000, 500_000, 2000500, (0.60, 0.30, 0.10), 0.10),
}


@dataclass
class Synth:
    codes: np.ndarray   # [nsnp, nseq] uint8 0..4
    POS: np.ndarray     # int32, ascending
    g: int
    paint: np.ndarray   # int32 cluster label per SNP (three contiguous thirds)
    nclust: int


def _draw_sites(rng, m, S, founders_of_seq, F, allele_probs, n_rate):
    """m candidate sites -> (codes [m, S], pass mask under the default filter)."""
    k = rng.choice([2, 3, 4], size=m, p=allele_probs)
    maf = rng.uniform(0.02, 0.5, size=m)
    # allele frequency table [m, 4]: major = 1 - maf, minors share maf evenly; unused alleles 0
    freq = np.zeros((m, 4))
    freq[:, 0] = 1 - maf
    for kk in (2, 3, 4):
        sel = k == kk
        freq[sel, 1:kk] = (maf[sel] / (kk - 1))[:, None]
    cdf = np.cumsum(freq, axis=1)
    # random letter assignment per site: permutation of ACGT
    perm = np.argsort(rng.random((m, 4)), axis=1).astype(np.uint8)
    fa = (rng.random((m, F))[:, :, None] > cdf[:, None, :]).sum(axis=2).clip(0, 3)      # founder allele ranks
    seq_rank = fa[:, founders_of_seq]                                                   # [m, S]
    mut = rng.random((m, S)) < 0.03
    redraw = (rng.random((m, S))[:, :, None] > cdf[:, None, :]).sum(axis=2).clip(0, 3)
    seq_rank = np.where(mut, redraw, seq_rank)
    codes = np.take_along_axis(perm, seq_rank.astype(np.int64), axis=1).astype(np.uint8)
    codes[rng.random((m, S)) < n_rate] = 4
    # reference default filter (src/getACGTNsites.cpp:104-134): >= 2 nucleotides present, gap share < 0.15,
    # second-largest nucleotide count > int(S * 0.01)
    cnt = np.stack([(codes == a).sum(axis=1) for a in range(5)], axis=1)
    present = (cnt[:, :4] > 0).sum(axis=1)
    second = np.sort(cnt[:, :4], axis=1)[:, 2]
    ok = (present > 1) & (cnt[:, 4] / S < 0.15) & (second > int(S * 0.01))
    return codes, ok


def generate(nseq: int, nsnp: int, seed: int, allele_probs=(0.847, 0.147, 0.006), n_rate: float = 0.01,
             g: int = G_DEFAULT, chunk: int = 20000) -> Synth:
    rng = np.random.default_rng(seed)
    POS = np.sort(rng.choice(np.arange(1, g + 1), size=nsnp, replace=False)).astype(np.int32)
    F = max(8, nseq // 16)
    wts = 1.0 / np.arange(1, F + 1)  # Zipf-ish founder popularity, exponent 1
    founders_of_seq = rng.choice(F, size=nseq, p=wts / wts.sum())
    out = np.empty((nsnp, nseq), dtype=np.uint8)
    have = 0
    while have < nsnp:
        m = min(chunk, int((nsnp - have) * 1.3) + 64)
        codes, ok = _draw_sites(rng, m, nseq, founders_of_seq, F, allele_probs, n_rate)
        good = codes[ok]
        take = min(len(good), nsnp - have)
        out[have:have + take] = good[:take]
        have += take
    paint = np.ones(nsnp, dtype=np.int32)
    paint[nsnp // 3:] = 2
    paint[2 * nsnp // 3:] = 3
    return Synth(codes=out, POS=POS, g=g, paint=paint, nclust=3)


def generate_config(name: str, nsnp_override: int = 0) -> Synth:
    S, n, seed, probs, nr = CONFIGS[name]
    if nsnp_override:
        n = nsnp_override
    return generate(S, n, seed, probs, nr)


def codes_to_alignment(codes: np.ndarray, rng_seed: int = 0, lowercase_frac: float = 0.0) -> np.ndarray:
    """ASCII alignment [nseq, nsnp] for the encoding path (classes 0..3 -> ACGT, 4 -> N or '-')."""
    rng = np.random.default_rng(rng_seed)
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    aln = lut[codes.T].copy()
    gaps = (codes.T == 4) & (rng.random(aln.shape) < 0.5)
    aln[gaps] = ord("-")
    if lowercase_frac > 0:
        lower = (rng.random(aln.shape) < lowercase_frac) & (aln != ord("-"))
        aln[lower] |= 0x20
    return aln


def exact_lr_links_approx(POS: np.ndarray, g: float, sr_dist: float) -> float:
    """Exact number of SNP pairs farther apart than sr_dist (circular), i.e. the quantity
    R/computePairwiseMI.R:94-97 estimates from a 10 % sample; used where R's RNG stream is not wanted."""
    P = np.sort(np.asarray(POS, dtype=np.int64))
    n = len(P)
    # pairs with linear distance d <= sr or d >= g - sr are short range
    hi = np.searchsorted(P, P + int(np.floor(sr_dist)), side="right")
    sr_lin = int((hi - np.arange(n) - 1).sum())
    wrap = np.searchsorted(P, P + int(np.ceil(g - sr_dist)), side="left")
    sr_wrap = int((n - wrap).sum())
    total = n * (n - 1) // 2
    return float(total - sr_lin - sr_wrap)


def cheap_codes(S, n, seed, n_rate, multi):
    """uint8-only generator for the big shapes: founders + 3 % re-draws + N; every site has >= 2 alleles."""
    rng = np.random.default_rng(seed)
    F = max(8, S // 16)
    wts = 1.0 / np.arange(1, F + 1)
    fos = rng.choice(F, size=S, p=wts / wts.sum())
    out = np.empty((n, S), dtype=np.uint8)
    for lo in range(0, n, 4096):
        m = min(4096, n - lo)
        k = rng.choice([2, 3, 4], size=m, p=multi)
        fa = (rng.integers(0, 256, (m, F), dtype=np.uint8) % 8)
        fa = np.where(fa < 5, 0, np.minimum(fa - 4, (k - 1)[:, None])).astype(np.uint8)   # major allele ~ 62 %
        fa[:, 0] = 0
        fa[:, 1] = 1
        c = fa[:, fos]
        mut = rng.integers(0, 256, (m, S), dtype=np.uint8) < 8
        c = np.where(mut, (rng.integers(0, 256, (m, S), dtype=np.uint8) % k[:, None]).astype(np.uint8), c)
        perm = np.argsort(rng.random((m, 4)), axis=1).astype(np.uint8)
        c = np.take_along_axis(perm, c.astype(np.int64), axis=1).astype(np.uint8)
        c[rng.integers(0, 65536, (m, S), dtype=np.uint16) < int(n_rate * 65536)] = 4
        out[lo:lo + m] = c
    return out
