#!/usr/bin/env python
"""Host ingest throughput (N2; reported separately from the kernels).

    python tools/ingest_bench.py [--sites 20000] [--samples 2504]

Writes a synthetic VCF (diploid, phased GT only), then times
  1. the native one-pass parser (sai_vcf_parse_gt, all host threads) -> int8 allele sums,
  1b. the same from a bgzipped file (sai_bgzf_inflate: parallel block inflate, then the parser),
  2. the host packer (sai_pack_i8) -> tiled bit-planes,
  2b. the zt wire encoder (sai_zt_encode) -> zero-suppressed tiles,
  3. the pure-Python reader on a slice (the cross-check implementation),
and prints one JSON line.
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sai_b200.configs import PloidyConfig  # noqa: E402
from sai_b200.encode import compress, pack_populations  # noqa: E402
from sai_b200.vcf import read_data, write_bgzf  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=20000)
    ap.add_argument("--samples", type=int, default=2504)
    a = ap.parse_args()
    rng = np.random.default_rng(1)
    n_ref, n_tgt = int(a.samples * 0.6), int(a.samples * 0.4) - 4
    n_src = a.samples - n_ref - n_tgt
    with tempfile.TemporaryDirectory() as tmp:
        vcf = os.path.join(tmp, "s.vcf")
        tokens = np.array(["0|0", "0|1", "1|0", "1|1"])
        with open(vcf, "w") as f:
            f.write("##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"i{k}" for k in range(a.samples)) + "\n")
            pos = np.cumsum(rng.geometric(1 / 40.0, size=a.sites))
            for p in pos:
                fq = rng.beta(0.2, 2.0)
                g = (rng.random(a.samples) < fq).astype(np.int8) * 2 + (rng.random(a.samples) < fq)
                f.write(f"1\t{p}\t.\tA\tT\t.\tPASS\t.\tGT\t" + "\t".join(tokens[g]) + "\n")
        size = os.path.getsize(vcf)
        for g, (lo, hi, name) in {"ref": (0, n_ref, "REF"), "tgt": (n_ref, n_ref + n_tgt, "TGT"), "src": (n_ref + n_tgt, a.samples, "SRC")}.items():
            with open(os.path.join(tmp, f"{g}.list"), "w") as f:
                f.writelines(f"{name}\ti{k}\n" for k in range(lo, hi))
        pc = PloidyConfig({"ref": {"REF": 2}, "tgt": {"TGT": 2}, "src": {"SRC": 2}})
        lists = [os.path.join(tmp, f"{g}.list") for g in ("ref", "tgt", "src")]
        read_data(vcf, "1", pc, *lists, None, None)  # warm the page cache
        t0 = time.perf_counter()
        d = read_data(vcf, "1", pc, *lists, None, None)
        t_parse = time.perf_counter() - t0
        # the same file bgzipped: blocks inflated in parallel by the native library
        bgz = os.path.join(tmp, "s.vcf.gz")
        write_bgzf(bgz, open(vcf, "rb").read(), level=6)
        bgz_size = os.path.getsize(bgz)
        read_data(bgz, "1", pc, *lists, None, None)
        t0 = time.perf_counter()
        d2 = read_data(bgz, "1", pc, *lists, None, None)
        t_bgz = time.perf_counter() - t0
        assert np.array_equal(d2["ref"][0]["REF"].GT, d["ref"][0]["REF"].GT)
        mats = [d["ref"][0]["REF"].GT, d["tgt"][0]["TGT"].GT, d["src"][0]["SRC"].GT]
        t0 = time.perf_counter()
        pg = pack_populations(mats, [2, 2, 2], d["ref"][0]["REF"].POS)
        t_pack = time.perf_counter() - t0
        compress(pg)
        t0 = time.perf_counter()
        zt = compress(pg)
        t_zt = time.perf_counter() - t0
        # the Python reader on the first 300 records
        small = os.path.join(tmp, "small.vcf")
        with open(vcf) as f, open(small, "w") as o:
            for i, line in enumerate(f):
                if i >= 302:
                    break
                o.write(line)
        t0 = time.perf_counter()
        read_data(small, "1", pc, *lists, None, None, native=False)
        t_py = time.perf_counter() - t0
    n_gt = a.sites * a.samples
    print(json.dumps({
        "sites": a.sites, "samples": a.samples, "vcf_bytes": size, "host_threads": os.cpu_count(),
        "native_parse_s": t_parse, "native_parse_MBps": size / t_parse / 1e6, "native_parse_Mgenotypes_per_s": n_gt / t_parse / 1e6,
        "bgzf_bytes": bgz_size, "bgzf_parse_s": t_bgz, "bgzf_parse_text_MBps": size / t_bgz / 1e6,
        "bgzf_parse_Mgenotypes_per_s": n_gt / t_bgz / 1e6,
        "pack_s": t_pack, "pack_Mgenotypes_per_s": n_gt / t_pack / 1e6, "packed_bytes": pg.nbytes,
        "zt_encode_s": t_zt, "zt_encode_MBps": pg.nbytes / t_zt / 1e6, "zt_bytes": zt.nbytes,
        "zt_ratio": pg.nbytes / max(1, zt.nbytes),
        "python_reader_Mgenotypes_per_s": 300 * a.samples / t_py / 1e6,
    }))


if __name__ == "__main__":
    main()
