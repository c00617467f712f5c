"""Small whole-genome fixture shared by tests/test_gpu_configs.py and the NCCL worker script."""

from __future__ import annotations

import numpy as np

HG19_MB = [249, 243, 198, 191, 181, 171, 159, 146, 141, 136, 135, 134, 115, 107, 103, 90, 81, 78, 59, 63, 48, 51]


def small_genome(total_sites=66_000, seed=400):
    """22 chromosomes with hg19-proportional site counts, ref 150 / tgt 100 / src 4 diploid,
    1 % missing: positions, int8 matrices and 50 kb / 10 kb windows per chromosome."""
    import synth
    from sai_b200.windows import split_genome

    pops = {"ref": {"REF": (150, 2)}, "tgt": {"TGT": (100, 2)}, "src": {"SRC": (4, 2)}}
    chroms = []
    for c, mb in enumerate(HG19_MB):
        n = max(200, int(total_sites * mb / sum(HG19_MB)))
        pos, mats = synth.make_populations(seed + c, n, pops, mean_gap=mb * 1e6 / n / 40.0, introgressed=0.03, missing=0.01)
        wins = split_genome([int(pos[0]), int(pos[-1])], 50_000, 10_000)
        chroms.append(dict(pos=pos, mats=[mats["ref"]["REF"], mats["tgt"]["TGT"], mats["src"]["SRC"]], wins=wins))
    return chroms


def score_rank(chroms, pieces, lay, job):
    """One rank's share: its pieces side by side, one genotype pass + one window launch."""
    from sai_b200.encode import pack_populations
    from sai_b200.genome import GenomeBatch, piece_site_range

    ranges = [piece_site_range(chroms[p.chrom]["pos"], chroms[p.chrom]["wins"], p) for p in pieces]
    batch = GenomeBatch(lay, [hi - lo for lo, hi in ranges], [chroms[p.chrom]["wins"][p.win_lo : p.win_hi] for p in pieces], 1,
                        cap_u=1 << 18, cap_q=1 << 19)
    for k, (p, (lo, hi)) in enumerate(zip(pieces, ranges)):
        ch = chroms[p.chrom]
        pg = pack_populations([m[lo:hi] for m in ch["mats"]], [2, 2, 2], ch["pos"][lo:hi], bits=[2, 2, 2])
        batch.load_piece(k, pg.packed, pg.pos)
    batch.score([job])
    return batch, batch.results()


def rank_rows(res, batch, pieces):
    """(chrom, window index) -> (nsnps, U, Q bits, U candidates, Q candidates)."""
    out = {}
    for p, sl in zip(pieces, batch.piece_slices()):
        for k, i in enumerate(range(sl.start, sl.stop)):
            out[(p.chrom, p.win_lo + k)] = (int(res.nsnps[0, i]), int(res.u[0, i]), float(res.q[0, i]).hex(),
                                            res.u_positions(0, i).tolist(), res.q_positions(0, i).tolist())
    return out
