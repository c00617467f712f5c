"""Host-side VCF ingest for the scoring path (GT only).

Stays on the host by design (the north star keeps VCF parsing off the GPU).
Produces what ``read_data(... is_phased=False, filter_*=False,
filter_missing=False)`` hands to the reference's ``WindowGenerator``
(sai/generators/window_generator.py:102-118): per population the positions and
the per-individual allele sums.  Follows

* ``parse_ind_file``        sai/utils/utils.py:31-75
* ``read_geno_data``        sai/utils/utils.py:78-186  (scikit-allel ``read_vcf``
  with ``numbers={"GT": ploidy}``, ``alt_number=1``, ``region``: GT is cut or
  padded with -1 to ``ploidy`` alleles, ``.`` = -1)
* ``read_anc_allele``       sai/utils/utils.py:435-489
* ``check_anc_allele`` / ``flip_snps``   sai/utils/utils.py:492-555
  (sites without ancestral allele or with one that is neither REF nor ALT are
  dropped; where ALT is ancestral every allele a becomes ``abs(a - 1)``, so a
  missing allele -1 becomes 2)
* ``reshape_genotypes(is_phased=False)``  sai/utils/utils.py:405-410
  (sum over the ploidy axis)

scikit-allel itself is a third-party dependency of the reference
(scikit-allel==1.3.7, pyproject.toml) and is not available here; this reader
covers plain-text and gzip/bgzip VCF.
"""

from __future__ import annotations

import gzip
import warnings
from typing import Optional

import numpy as np

from .encode import PopData


def parse_ind_file(filename: str) -> dict[str, list[str]]:
    """``category sample`` lines -> ``{category: [samples]}``; lines that do
    not have exactly two fields are skipped."""
    try:
        samples: dict[str, list[str]] = {}
        with open(filename, "r") as f:
            for line in f:
                parts = line.strip().split()
                if len(parts) != 2:
                    continue
                samples.setdefault(parts[0], []).append(parts[1])
        if not samples:
            raise ValueError(f"No samples found in {filename}. Please check your data.")
    except FileNotFoundError:
        raise FileNotFoundError(f"File '{filename}' not found. Please check the file path.")
    return samples


def _open_text(path: str):
    with open(path, "rb") as f:
        magic = f.read(2)
    if magic == b"\x1f\x8b":
        return gzip.open(path, "rt")
    return open(path, "r")


def _parse_gt(field: str, ploidy: int, out: np.ndarray) -> None:
    gt = field.split(":", 1)[0]
    n = 0
    tok = ""
    for ch in gt:
        if ch == "|" or ch == "/":
            if n < ploidy:
                out[n] = -1 if tok in (".", "") else int(tok)
            n += 1
            tok = ""
        else:
            tok += ch
    if n < ploidy:
        out[n] = -1 if tok in (".", "") else int(tok)
    n += 1
    for k in range(n, ploidy):
        out[k] = -1


class VcfRegion:
    """All records of one chromosome region, parsed once; populations are then
    cut out by sample name with their own ploidy (the reference re-reads the
    file once per population, sai/utils/utils.py:716-734)."""

    def __init__(self, vcf_file: str, chr_name: str, start: Optional[int] = None, end: Optional[int] = None):
        self.samples: list[str] = []
        pos, ref, alt, rows = [], [], [], []
        try:
            with _open_text(vcf_file) as f:
                for line in f:
                    if line.startswith("##"):
                        continue
                    if line.startswith("#CHROM"):
                        self.samples = line.rstrip("\n").split("\t")[9:]
                        continue
                    tab = line.find("\t")
                    if line[:tab] != chr_name:
                        continue
                    cols = line.rstrip("\n").split("\t")
                    p = int(cols[1])
                    if start is not None and end is not None and not (start <= p <= end):
                        continue
                    pos.append(p)
                    ref.append(cols[3])
                    alt.append(cols[4].split(",")[0])
                    fmt = cols[8].split(":")
                    gi = fmt.index("GT") if "GT" in fmt else 0
                    rows.append([c.split(":")[gi] if gi else c for c in cols[9:]])
        except FileNotFoundError:
            raise
        except Exception as e:  # same wrapping as utils.py:136-137
            region = chr_name if start is None and end is None else f"{chr_name}:{start}-{end}"
            raise ValueError(f"Failed to read VCF file {vcf_file} from {region}: {e}") from e
        self.pos = np.asarray(pos, dtype=np.int32)
        self.ref = ref
        self.alt = alt
        self.rows = rows

    def genotypes(self, sample_names: list[str], ploidy: int) -> np.ndarray:
        """int8 ``(sites, individuals, ploidy)`` like ``calldata/GT``."""
        idx = [self.samples.index(s) for s in sample_names]
        gt = np.full((len(self.rows), len(idx), ploidy), -1, dtype=np.int8)
        for r, row in enumerate(self.rows):
            for c, i in enumerate(idx):
                _parse_gt(row[i], ploidy, gt[r, c])
        return gt


def read_anc_allele(anc_allele_file: str, chr_name: str, start=None, end=None) -> dict[int, str]:
    """BED ``chrom start end allele`` -> ``{end: allele}`` for one chromosome."""
    anc: dict[int, str] = {}
    try:
        with open(anc_allele_file, "r") as f:
            for line in f:
                e = line.rstrip().split()
                chrom, pos, allele = e[0], int(e[2]), e[3]
                if chrom != chr_name:
                    continue
                if (start is not None and pos < start) or (end is not None and pos > end):
                    continue
                anc[pos] = allele
    except FileNotFoundError as exc:
        raise FileNotFoundError(f"File {anc_allele_file} not found.") from exc
    if not anc:
        if start is not None or end is not None:
            raise ValueError(
                f"No ancestral allele is found for chromosome {chr_name} in the region {start}-{end}."
            )
        raise ValueError(f"No ancestral allele is found for chromosome {chr_name}.")
    return anc


def _polarise(pos: np.ndarray, ref: list[str], alt: list[str], gt: np.ndarray, anc: dict[int, str]):
    """Keeps sites whose ancestral allele is REF or ALT; flips those where it is
    ALT (``abs(a - 1)`` on every allele, missing included)."""
    keep = np.zeros(pos.shape[0], dtype=bool)
    flip = np.zeros(pos.shape[0], dtype=bool)
    for i, p in enumerate(pos.tolist()):
        a = anc.get(p)
        if a is None or a not in (ref[i], alt[i]):
            continue
        keep[i] = True
        flip[i] = a == alt[i]  # utils.py:523-524
    gt = gt[keep].copy()
    f = flip[keep]
    gt[f] = np.abs(gt[f].astype(np.int16) - 1).astype(np.int8)
    return pos[keep], gt


def load_population_group(
    region: VcfRegion,
    sample_file: Optional[str],
    group: str,
    ploidy_config,
    anc: Optional[dict[int, str]],
):
    """``(data, samples)`` of one group like ``_load_population_data``
    (sai/utils/utils.py:649-761) with ``is_phased=False`` and no filtering."""
    if sample_file is None:
        return None, None
    samples = parse_ind_file(sample_file)
    if group not in ploidy_config.root:
        raise ValueError(f"Ploidy configuration missing group '{group}'.")
    ploidies = ploidy_config.root[group]
    for population in ploidies:
        if population not in samples:
            raise ValueError(
                f"Population '{population}' in ploidy_config[{group}] not found in sample file: {sample_file}"
            )
    data: dict[str, PopData] = {}
    for population, names in samples.items():
        if population not in ploidies:
            warnings.warn(
                f"Population '{population}' found in sample file but not in ploidy_config[{group}]; skipping.",
                RuntimeWarning,
            )
            continue
        if region.pos.size == 0:
            continue
        try:
            gt = region.genotypes(names, ploidies[population])
        except Exception as e:
            raise ValueError(f"Failed to read VCF data for {sample_file}, population '{population}': {e}")
        pos = region.pos
        if anc is not None:
            pos, gt = _polarise(pos, region.ref, region.alt, gt, anc)
        data[population] = PopData(pos.copy(), gt.sum(axis=2, dtype=np.int64).astype(np.int8))
    if not data:
        return None, samples
    return data, samples


def read_data(
    vcf_file: str,
    chr_name: str,
    ploidy_config,
    ref_ind_file: Optional[str],
    tgt_ind_file: Optional[str],
    src_ind_file: Optional[str],
    out_ind_file: Optional[str],
    anc_allele_file: Optional[str],
    start: Optional[int] = None,
    end: Optional[int] = None,
):
    """``{"ref": (data, samples), "tgt": ..., "src": ..., "outgroup": ...}``
    (sai/utils/utils.py:215-356 with the WindowGenerator's fixed flags)."""
    region = VcfRegion(vcf_file, chr_name, start, end)
    anc = None
    if anc_allele_file:
        anc = read_anc_allele(anc_allele_file, chr_name, start, end)
    out = {}
    for group, ind_file in (("ref", ref_ind_file), ("tgt", tgt_ind_file), ("src", src_ind_file), ("outgroup", out_ind_file)):
        if ind_file is None or (group == "outgroup" and group not in ploidy_config.root):
            out[group] = (None, None)
            continue
        out[group] = load_population_group(region, ind_file, group, ploidy_config, anc)
    return out
