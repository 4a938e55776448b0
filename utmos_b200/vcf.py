"""Host-side VCF genotype reader: text -> (samples, int8 GT[V, S, 2]).

Stands in for ``allel.read_vcf(fields=["calldata/GT", "samples"])`` (utmos/convert.py:50-53, scikit-allel
1.3.5): genotypes are parsed to a fixed ploidy-2 int8 tensor, a missing allele ('.', or an absent second
allele of a haploid call) is -1, phasing is ignored.  Text parsing stays on the host (SURVEY.md section 2
row 3); the numeric work on the tensor happens in the K1 kernel (csrc/convert.cu).
"""
import gzip

import numpy as np


def _open(path):
    if path.endswith(".gz"):
        return gzip.open(path, "rt")
    return open(path, "r")


def _parse_gt(token):
    """'0|1' / '1/2' / './.' / '1' -> (a0, a1)."""
    sep = token.find("|")
    if sep < 0:
        sep = token.find("/")
    if sep < 0:
        first, second = token, "."
    else:
        first, second = token[:sep], token[sep + 1:]
        nxt = second.find("|")
        if nxt < 0:
            nxt = second.find("/")
        if nxt >= 0:
            second = second[:nxt]                   # ploidy > 2 is truncated like allel's numbers=2
    a0 = int(first) if first.isdigit() else -1
    a1 = int(second) if second.isdigit() else -1
    return a0, a1


def read_vcf_genotypes(path, chunk_length=2000):
    """Yield (samples ndarray[str], int8 GT [n<=chunk_length, S, 2]) blocks in file order."""
    samples = None
    rows = []
    cache = {}
    with _open(path) as fh:
        for line in fh:
            if line.startswith("##"):
                continue
            if line.startswith("#"):
                samples = np.array(line.rstrip("\n").split("\t")[9:])
                continue
            fields = line.rstrip("\n").split("\t")
            fmt = fields[8].split(":")
            try:
                gi = fmt.index("GT")
            except ValueError:
                gi = -1
            row = np.full((len(samples), 2), -1, dtype=np.int8)
            if gi >= 0:
                for s, call in enumerate(fields[9:]):
                    tok = call if gi == 0 and ":" not in call else call.split(":")[gi]
                    got = cache.get(tok)
                    if got is None:
                        got = _parse_gt(tok)
                        cache[tok] = got
                    row[s, 0] = got[0]
                    row[s, 1] = got[1]
            rows.append(row)
            if len(rows) >= chunk_length:
                yield samples, np.stack(rows)
                rows = []
    if samples is None:
        raise ValueError(f"{path}: no #CHROM header line")
    if rows:
        yield samples, np.stack(rows)
    elif samples is not None and not rows:
        yield samples, np.zeros((0, len(samples), 2), dtype=np.int8)
