"""Generates the golden fixtures in this directory by running the REFERENCE
itself (imported from /root/reference in the build container).

    python tests/golden/make_golden.py

The reference's arithmetic (``sai.stats``), window extraction
(``WindowGenerator._window_generator``), dispatch (``FeaturePreprocessor.run``)
and text output (``FeaturePreprocessor.process_items``) run unmodified.  Three
third-party modules that are not installed here are stubbed at import
(``allel`` -- imported but unused on this path --, ``pysam``, ``natsort``) and
``WindowGenerator.__init__`` (which calls ``allel.read_vcf``) is bypassed by
filling the attributes ``_window_generator`` reads.

Outputs (committed; /root/reference does not exist on the GPU box):
  stat_cases.npz / stat_cases.json   random stat-class level cases -> U, Q, cdd_pos
  pipe_<name>.npz / pipe_<name>.json in-memory chunks -> items + TSV / log text
  vcf_<name>.vcf + .json             normalised copies of the reference's test
                                     VCFs (same positions / alleles / genotypes,
                                     header and INFO stripped) + expected items
"""

import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# ---- stubs, then the reference ---------------------------------------------
allel = types.ModuleType("allel")
allel.GenotypeVector = allel.GenotypeArray = object
sys.modules["allel"] = allel
sys.modules["pysam"] = types.ModuleType("pysam")
ns = types.ModuleType("natsort")
ns.natsorted = sorted
sys.modules["natsort"] = ns
sys.path.insert(0, REF)

import sai.stats  # noqa: E402,F401  (registers the statistics)
from sai.configs import PloidyConfig, StatConfig  # noqa: E402
from sai.generators.window_generator import WindowGenerator  # noqa: E402
from sai.preprocessors.feature_preprocessor import FeaturePreprocessor  # noqa: E402
from sai.stats import DdStatistic, QStatistic, UStatistic  # noqa: E402
from sai.utils import split_genome  # noqa: E402
from sai.utils.genomic_dataclasses import ChromosomeData  # noqa: E402

import synth  # noqa: E402
from sai_b200 import vcf as my_vcf  # noqa: E402  (host ingest only; no GPU needed)
from sai_b200.configs import PloidyConfig as MyPloidy  # noqa: E402


def fhex(x):
    x = float(x)
    return "nan" if x != x else x.hex()


# ---- 1. stat-class level random cases --------------------------------------
def stat_cases(n_cases=400, seed=20261018):
    rng = np.random.default_rng(seed)
    ops = ["=", "<", ">", "<=", ">="]
    arrays, meta = {}, []
    for c in range(n_cases):
        n_sites = int(rng.integers(1, 60))
        n_src = int(rng.integers(1, 4))
        ploidy = [int(rng.integers(1, 5)) for _ in range(2 + n_src)]
        n_ind = [int(rng.integers(1, 70)), int(rng.integers(1, 40))] + [int(rng.integers(1, 4)) for _ in range(n_src)]
        miss = float(rng.choice([0.0, 0.0, 0.05, 0.3]))
        mats = []
        style = int(rng.integers(0, 3))
        f = rng.beta(0.3, 0.6, size=n_sites)
        for k, (n, p) in enumerate(zip(n_ind, ploidy)):
            if k >= 2 and style != 0:  # src mostly fixed so that "=1"/"=0" hit
                fk = np.where(rng.random(n_sites) < 0.6, np.round(f), f)
            elif k == 0:
                fk = f * rng.choice([0.05, 0.5, 1.0])
            else:
                fk = f
            g = rng.binomial(p, fk[:, None], size=(n_sites, n)).astype(np.int64)
            if miss > 0:
                m = rng.random(g.shape) < miss
                g[m] = rng.choice([-1, -2], size=int(m.sum()))
            mats.append(g)
        pos = np.sort(rng.choice(np.arange(1, 100000), size=n_sites, replace=False)).astype(np.int64)
        ychoices = [0.0, 1.0, 0.5, 0.25, 0.8, 1 / 3]
        y_list = [(str(rng.choice(ops)) if rng.random() < 0.5 else "=", float(rng.choice(ychoices))) for _ in range(n_src)]
        w = float(rng.choice([0.01, 0.1, 0.3, 0.5, 1.0]))
        x = float(rng.choice([0.0, 0.01, 0.2, 0.5, 0.8]))
        q = float(rng.choice([0.0, 0.25, 0.5, 0.9, 0.95, 1.0]))
        anc = bool(rng.integers(0, 2))
        kw = dict(ref_gts=mats[0], tgt_gts=mats[1], src_gts_list=mats[2:], ref_ploidy=ploidy[0],
                  tgt_ploidy=ploidy[1], src_ploidy_list=ploidy[2:])
        ru = UStatistic(**kw).compute(pos=pos, w=w, x=x, y_list=y_list, anc_allele_available=anc)
        rq = QStatistic(**kw).compute(pos=pos, w=w, y_list=y_list, quantile=q, anc_allele_available=anc)
        rd = DdStatistic(**kw).compute()
        for k, m in enumerate(mats):
            arrays[f"c{c}_g{k}"] = m.astype(np.int8)
        arrays[f"c{c}_pos"] = pos.astype(np.int32)
        meta.append(dict(ploidy=ploidy, y_list=y_list, w=w, x=x, q=q, anc=anc, n_src=n_src,
                         U=int(ru["value"]), U_pos=[int(p) for p in ru["cdd_pos"]],
                         Q=fhex(rq["value"]), Q_pos=[int(p) for p in rq["cdd_pos"]],
                         DD=[fhex(v) for v in rd["value"]]))
    np.savez_compressed(os.path.join(HERE, "stat_cases.npz"), **arrays)
    with open(os.path.join(HERE, "stat_cases.json"), "w") as f:
        json.dump(meta, f)
    print("stat_cases:", n_cases)


# ---- 2./3. pipeline level: the reference's window loop ----------------------
def run_reference_pipeline(chr_name, start, end, win_len, win_step, data, ploidies, stats, anc):
    """data = {"ref": {pop: (POS, GT int64)}, "tgt": ..., "src": ...}"""
    wg = object.__new__(WindowGenerator)
    wg.win_len, wg.win_step, wg.chr_name = win_len, win_step, chr_name
    wg.num_src = len(data["src"])
    wg.ploidy_config = PloidyConfig(ploidies)
    mk = lambda d: {p: ChromosomeData(POS=np.asarray(pos), REF=None, ALT=None, GT=np.asarray(gt, dtype=np.int64)) for p, (pos, gt) in d.items()}
    wg.ref_data, wg.tgt_data, wg.src_data = mk(data["ref"]), mk(data["tgt"]), mk(data["src"])
    wg.out_data = mk(data["outgroup"]) if data.get("outgroup") else None
    names = lambda d: {p: [f"{p}_{i}" for i in range(np.asarray(gt).shape[1])] for p, (pos, gt) in d.items()}
    wg.ref_samples, wg.tgt_samples, wg.src_samples = names(data["ref"]), names(data["tgt"]), names(data["src"])
    wg.out_samples = names(data["outgroup"]) if data.get("outgroup") else None
    wg.src_combinations = [tuple(data["src"].keys())]
    wg.tgt_windows = {
        t: split_genome(
            pos=(wg.tgt_data[t].POS if (start is None and end is None) else [start, end - win_len + win_step]),
            window_size=win_len, step_size=win_step, start=start)
        for t in data["tgt"]
    }
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "scores.tsv")
        fp = FeaturePreprocessor(out, StatConfig(json.loads(json.dumps(stats))), anc_allele_available=anc)
        items = []
        for w in wg.get():
            items.extend(fp.run(**w))
        fp.process_items(items)
        text = {"tsv": open(out).read()}
        for key in ("U", "Q"):
            p = os.path.join(tmp, f"scores.{key}.log")
            if os.path.exists(p):
                text[key] = open(p).read()
    exp = []
    for it in items:
        e = {k: it[k] for k in ("chr_name", "start", "end", "ref_pop", "tgt_pop", "out_pop", "nsnps")}
        e["start"], e["end"] = int(e["start"]), int(e["end"])
        e["src_pop_list"] = list(it["src_pop_list"])
        for s in ("U", "Q"):
            if s in it:
                v = it[s]
                e[s] = fhex(v) if (s == "Q" or v != v) else int(v)
                e[s + "_pos"] = [int(p) for p in it["cdd_pos"][s]]
        for s in ("Danc", "Dplus", "df", "fd", "DD"):
            if s in it:
                v = it[s]
                e[s] = [fhex(x) for x in v] if isinstance(v, list) else fhex(v)
        exp.append(e)
    return exp, text


PIPE_CASES = {
    # name: (seed, n_sites, mean_gap, pops, kwargs, win_len, win_step, stats, anc, chunk)
    "diploid_basic": dict(
        seed=1, n_sites=4000, gap=150.0,
        pops={"ref": {"AFR": (60, 2)}, "tgt": {"EUR": (40, 2)}, "src": {"NEA": (2, 2)}},
        synth={}, win=(50000, 10000),
        stats={"U": {"ref": {"AFR": 0.05}, "tgt": {"EUR": 0.3}, "src": {"NEA": "=1"}},
               "Q": {"ref": {"AFR": 0.05}, "tgt": {"EUR": 0.95}, "src": {"NEA": "=1"}}},
        anc=True),
    "two_src_missing_noanc": dict(
        seed=2, n_sites=3000, gap=100.0,
        pops={"ref": {"AFR": (50, 2)}, "tgt": {"EAS": (30, 2)}, "src": {"NEA": (2, 2), "DEN": (1, 2)}},
        synth=dict(missing=0.02, src_all_missing=0.005), win=(40000, 40000),
        stats={"U": {"ref": {"AFR": 0.1}, "tgt": {"EAS": 0.2}, "src": {"NEA": "=1", "DEN": "=1"}},
               "Q": {"ref": {"AFR": 0.1}, "tgt": {"EAS": 0.95}, "src": {"NEA": "=1", "DEN": "=1"}}},
        anc=False),
    "haploid_ops": dict(
        seed=3, n_sites=2500, gap=80.0,
        pops={"ref": {"R": (40, 1)}, "tgt": {"T": (25, 1)}, "src": {"S1": (2, 1), "S2": (2, 1)}},
        synth=dict(missing=0.01), win=(20000, 5000),
        stats={"U": {"ref": {"R": 0.2}, "tgt": {"T": 0.1}, "src": {"S1": ">=0.5", "S2": "<=0.5"}},
               "Q": {"ref": {"R": 0.3}, "tgt": {"T": 0.9}, "src": {"S1": ">0.4", "S2": "<1"}}},
        anc=False),
    "mixed_ploidy_multi_pop": dict(
        seed=4, n_sites=2000, gap=120.0,
        pops={"ref": {"R1": (30, 2), "R2": (17, 4)}, "tgt": {"T1": (20, 4), "T2": (33, 2)}, "src": {"S": (3, 3)}},
        synth=dict(missing=0.03), win=(30000, 15000),
        stats={"U": {"ref": {"R1": 0.1, "R2": 0.2}, "tgt": {"T1": 0.3, "T2": 0.1}, "src": {"S": "=1"}},
               "Q": {"ref": {"R1": 0.1, "R2": 0.2}, "tgt": {"T1": 0.95, "T2": 0.5}, "src": {"S": "=1"}}},
        anc=True),
    "loose_big_windows": dict(
        seed=5, n_sites=6000, gap=20.0,
        pops={"ref": {"R": (40, 2)}, "tgt": {"T": (50, 2)}, "src": {"S": (2, 2)}},
        synth=dict(missing=0.01), win=(60000, 30000),
        stats={"U": {"ref": {"R": 0.5}, "tgt": {"T": 0.01}, "src": {"S": ">=0"}},
               "Q": {"ref": {"R": 1.0}, "tgt": {"T": 0.95}, "src": {"S": ">=0"}}},
        anc=False),
    "gaps_empty_windows": dict(
        seed=6, n_sites=300, gap=2000.0,
        pops={"ref": {"R": (20, 2)}, "tgt": {"T": (20, 2)}, "src": {"S": (1, 2)}},
        synth=dict(introgressed=0.2), win=(1000, 500),
        stats={"U": {"ref": {"R": 0.3}, "tgt": {"T": 0.2}, "src": {"S": "=1"}},
               "Q": {"ref": {"R": 0.3}, "tgt": {"T": 0.75}, "src": {"S": "=1"}}},
        anc=True),
    "fourpop_two_src": dict(
        seed=8, n_sites=2500, gap=120.0,
        pops={"ref": {"AFR": (60, 2)}, "tgt": {"EUR": (40, 2)}, "src": {"NEA": (2, 2), "DEN": (1, 2)}},
        synth=dict(missing=0.01, src_all_missing=0.004), win=(40000, 20000),
        stats={"Danc": True, "Dplus": True, "df": True, "fd": True,
               "U": {"ref": {"AFR": 0.05}, "tgt": {"EUR": 0.3}, "src": {"NEA": "=1", "DEN": "=1"}},
               "Q": {"ref": {"AFR": 0.05}, "tgt": {"EUR": 0.95}, "src": {"NEA": "=1", "DEN": "=1"}}},
        anc=True),
    "fourpop_outgroup": dict(
        seed=9, n_sites=2000, gap=100.0,
        pops={"ref": {"R": (50, 2)}, "tgt": {"T": (30, 4)}, "src": {"S": (3, 1)}, "outgroup": {"O": (2, 2)}},
        synth={}, win=(30000, 10000),
        stats={"fd": True, "df": True, "Danc": True, "Dplus": False,
               "U": {"ref": {"R": 0.1}, "tgt": {"T": 0.2}, "src": {"S": "=1"}}},
        anc=True),
    "dd_two_src_missing": dict(
        seed=10, n_sites=2500, gap=100.0,
        pops={"ref": {"AFR": (70, 2)}, "tgt": {"EUR": (45, 2)}, "src": {"NEA": (3, 2), "DEN": (1, 2)}},
        synth=dict(missing=0.02, src_all_missing=0.004), win=(40000, 20000),
        stats={"DD": True,
               "U": {"ref": {"AFR": 0.05}, "tgt": {"EUR": 0.3}, "src": {"NEA": "=1", "DEN": "=1"}}},
        anc=False),
    "dd_mixed_ploidy_many_src": dict(
        seed=11, n_sites=1500, gap=150.0,
        pops={"ref": {"R1": (33, 4), "R2": (20, 1)}, "tgt": {"T": (40, 3)}, "src": {"S": (11, 2), "S4": (2, 4)}},
        synth=dict(missing=0.03), win=(30000, 15000),
        stats={"fd": True, "DD": True, "Danc": False,
               "Q": {"ref": {"R1": 0.2, "R2": 0.2}, "tgt": {"T": 0.9}, "src": {"S": ">=0.5", "S4": ">=0.5"}}},
        anc=True),
    "u_only_chunked": dict(
        seed=7, n_sites=1500, gap=100.0,
        pops={"ref": {"R": (64, 2)}, "tgt": {"T": (32, 2)}, "src": {"S": (2, 2)}},
        synth={}, win=(20000, 10000),
        stats={"U": {"ref": {"R": 0.05}, "tgt": {"T": 0.5}, "src": {"S": "=1"}}},
        anc=True, chunk=(30001, 110000)),
}


def pipe_cases():
    for name, c in PIPE_CASES.items():
        pos, mats = synth.make_populations(c["seed"], c["n_sites"], c["pops"], mean_gap=c["gap"], **c["synth"])
        chunk = c.get("chunk")
        if chunk is not None:  # what a region read returns
            keep = (pos >= chunk[0]) & (pos <= chunk[1])
            pos = pos[keep]
            mats = {g: {p: m[keep] for p, m in d.items()} for g, d in mats.items()}
            start, end = chunk
        else:
            start, end = 1, int(pos[-1]) // c["win"][1] * c["win"][1] + c["win"][0]
        data = {g: {p: (pos, m) for p, m in d.items()} for g, d in mats.items()}
        ploidies = {g: {p: c["pops"][g][p][1] for p in c["pops"][g]} for g in c["pops"]}
        exp, text = run_reference_pipeline("1", start, end, c["win"][0], c["win"][1], data, ploidies, c["stats"], c["anc"])
        arrays = {"pos": pos.astype(np.int32)}
        for g, d in mats.items():
            for p, m in d.items():
                arrays[f"{g}__{p}"] = m.astype(np.int8)
        np.savez_compressed(os.path.join(HERE, f"pipe_{name}.npz"), **arrays)
        with open(os.path.join(HERE, f"pipe_{name}.json"), "w") as f:
            json.dump(dict(chr_name="1", start=start, end=end, win_len=c["win"][0], win_step=c["win"][1],
                           ploidies=ploidies, stats=c["stats"], anc=c["anc"], items=exp, text=text), f)
        nU = sum(1 for e in exp if isinstance(e.get("U"), int) and e["U"] > 0)
        nQ = sum(1 for e in exp if e.get("Q", "nan") != "nan")
        print(f"pipe_{name}: {len(exp)} items, {nU} with U>0, {nQ} with Q")


# ---- 3. the reference's own test VCFs --------------------------------------
def write_normalised_vcf(src, dst):
    """Same records; header reduced to fileformat + GT + #CHROM, INFO -> '.'."""
    import gzip

    opener = my_vcf._open_text
    with opener(src) as f, (gzip.open(dst, "wt") if dst.endswith(".gz") else open(dst, "w")) as out:
        out.write("##fileformat=VCFv4.1\n")
        out.write('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n')
        for line in f:
            if line.startswith("##"):
                continue
            if line.startswith("#CHROM"):
                out.write(line)
                continue
            c = line.rstrip("\n").split("\t")
            c[2], c[5], c[6], c[7] = ".", ".", ".", "."
            out.write("\t".join(c) + "\n")


VCF_CASES = {
    # test_sai.py:45-63 -> Q == 0.9
    "example_q": dict(vcf="example.vcf", chr="21", lists=("example.ref.ind.list", "example.tgt.ind.list", "example.src.ind.list"),
                      ploidies={"ref": {"AFR": 2}, "tgt": {"CHB": 2}, "src": {"Nean": 2}},
                      stats={"Q": {"ref": {"AFR": 0.3}, "tgt": {"CHB": 0.95}, "src": {"Nean": "=1"}}},
                      win=(6666, 6666), anc=None, first_last=True),
    # test_feature_preprocessor.py:196-223 -> U == 3
    "example_u": dict(vcf="example.vcf", chr="21", lists=("example.ref.ind.list", "example.tgt.ind.list", "example.src.ind.list"),
                      ploidies={"ref": {"AFR": 2}, "tgt": {"CHB": 2}, "src": {"Nean": 2}},
                      stats={"U": {"ref": {"AFR": 0.3}, "tgt": {"CHB": 0.5}, "src": {"Nean": "=1"}}},
                      win=(6666, 6666), anc=None, first_last=True),
    # test_sai.py:127-151 -> U rows [0, 1]
    "mixed_ploidy": dict(vcf="test.mixed.ploidy.data.vcf.gz", chr="21", lists=("test.ref.ind.list", "test.tgt.ind.list", "test.src.ind.list"),
                         ploidies={"ref": {"ref1": 2}, "tgt": {"tgt1": 4, "tgt2": 4}, "src": {"src1": 4, "src2": 4}},
                         stats={"U": {"ref": {"ref1": 0.3}, "tgt": {"tgt1": 0.8, "tgt2": 0.8}, "src": {"src1": "=1", "src2": "=1"}}},
                         win=(50000, 50000), anc="test.mixed.ploidy.data.anc.alleles", first_last=True),
    # test_sai.py:92-110 -> fd, df, Danc, Dplus of tests/data/test.with.outgroup.res.tsv
    "outgroup_stats": dict(vcf="test.with.outgroup.vcf.gz", chr="1", lists=("test.with.outgroup.ref.list", "test.with.outgroup.tgt.list", "test.with.outgroup.src.list"),
                           out_list="test.with.outgroup.out.list",
                           ploidies={"ref": {"ref": 2}, "tgt": {"tgt": 2}, "src": {"src": 2}, "outgroup": {"out": 2}},
                           stats={"fd": True, "df": True, "Danc": True, "Dplus": True},
                           win=(40000, 40000), anc="test.with.outgroup.anc.alleles", first_last=True),
    # realistic shape (SLiM, 1008 ref + 503 tgt + 1 src), goldens minted here
    "outgroup_shape": dict(vcf="test.with.outgroup.vcf.gz", chr="1", lists=("test.with.outgroup.ref.list", "test.with.outgroup.tgt.list", "test.with.outgroup.src.list"),
                           ploidies=None,
                           stats=None, win=(40000, 10000), anc="test.with.outgroup.anc.alleles", first_last=True),
}


def vcf_cases():
    for name, c in VCF_CASES.items():
        src = os.path.join(REF, "tests", "data", c["vcf"])
        dst_name = f"vcf_{name}.vcf" + (".gz" if os.path.getsize(src) > 20000 else "")
        write_normalised_vcf(src, os.path.join(HERE, dst_name))
        lists = {}
        for g, fn in zip(("ref", "tgt", "src"), c["lists"]):
            lists[g] = my_vcf.parse_ind_file(os.path.join(REF, "tests", "data", fn))
        ploidies, stats = c["ploidies"], c["stats"]
        if ploidies is None:
            ploidies = {g: {p: 2 for p in lists[g]} for g in lists}
            r, t, s = (list(lists[g])[0] for g in ("ref", "tgt", "src"))
            stats = {"U": {"ref": {r: 0.05}, "tgt": {t: 0.05}, "src": {s: ">=0.5"}},
                     "Q": {"ref": {r: 0.05}, "tgt": {t: 0.95}, "src": {s: ">=0.5"}}}
        anc_map = None
        if c["anc"]:
            anc_lines = [l.split() for l in open(os.path.join(REF, "tests", "data", c["anc"])) if l.strip()]
            with open(os.path.join(HERE, f"vcf_{name}.anc.bed"), "w") as f:
                for e in anc_lines:
                    f.write("\t".join(e[:4]) + "\n")
        # what ChunkGenerator does: first/last POS of the chromosome -> one chunk (sai.py:86-93)
        region = my_vcf.VcfRegion(os.path.join(HERE, dst_name), c["chr"])
        from sai_b200.windows import split_genome as my_split, split_windows_ranges
        wins = my_split([int(region.pos[0]), int(region.pos[-1])], c["win"][0], c["win"][1])
        start, end = split_windows_ranges(wins, 1)[0]
        out_path = None
        if c.get("out_list"):
            out_path = _write_list(name, "outgroup", my_vcf.parse_ind_file(os.path.join(REF, "tests", "data", c["out_list"])))
        groups = my_vcf.read_data(
            os.path.join(HERE, dst_name), c["chr"], MyPloidy(ploidies),
            *[_write_list(name, g, lists[g]) for g in ("ref", "tgt", "src")], out_path,
            os.path.join(HERE, f"vcf_{name}.anc.bed") if c["anc"] else None, start=start, end=end)
        data = {g: {p: (d.POS, d.GT.astype(np.int64)) for p, d in groups[g][0].items()}
                for g in ("ref", "tgt", "src", "outgroup") if groups[g][0] is not None}
        exp, text = run_reference_pipeline(c["chr"], start, end, c["win"][0], c["win"][1], data, ploidies, stats, c["anc"] is not None)
        with open(os.path.join(HERE, f"vcf_{name}.json"), "w") as f:
            json.dump(dict(vcf=dst_name, chr_name=c["chr"], start=start, end=end, win_len=c["win"][0], win_step=c["win"][1],
                           ploidies=ploidies, stats=stats, anc=c["anc"] is not None, items=exp, text=text), f)
        print(f"vcf_{name}: {len(exp)} items:", [(e.get('U'), e.get('Q')) for e in exp][:6])


def _write_list(case, group, pops):
    path = os.path.join(HERE, f"vcf_{case}.{group}.list")
    with open(path, "w") as f:
        for p, names in pops.items():
            for n in names:
                f.write(f"{p}\t{n}\n")
    return path


# ---- 4. the input of the reference's own ingest KATs -------------------------
def ingest_kat_inputs():
    """tests/utils/test_utils.py:204-209 (read_anc_allele), :269-318 (check_anc_allele) run on
    tests/data/test.data.vcf + test.anc.allele.bed + test.{ref,tgt,src}.ind.list: normalised
    copies, so that tests/test_cabi_host.py can assert the expected values the reference's tests
    hold (transcribed there with file:line) on the GPU box too."""
    write_normalised_vcf(os.path.join(REF, "tests", "data", "test.data.vcf"), os.path.join(HERE, "vcf_testdata.vcf"))
    with open(os.path.join(HERE, "vcf_testdata.anc.bed"), "w") as f:
        for l in open(os.path.join(REF, "tests", "data", "test.anc.allele.bed")):
            if l.strip():
                f.write("\t".join(l.split()[:4]) + "\n")
    for g in ("ref", "tgt", "src"):
        _write_list("testdata", g, my_vcf.parse_ind_file(os.path.join(REF, "tests", "data", f"test.{g}.ind.list")))
    print("vcf_testdata: inputs of the reference's ingest KATs")


if __name__ == "__main__":
    steps = {"stat": stat_cases, "pipe": pipe_cases, "vcf": vcf_cases, "ingest": ingest_kat_inputs}
    for name in (sys.argv[1:] or list(steps)):
        steps[name]()
