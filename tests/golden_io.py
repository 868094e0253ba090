"""Loader of the committed golden fixtures (outputs of the unmodified reference, see tests/golden/make_golden.py)."""
import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def parse_fasta_bytes(b):
    out = []
    lines = b.split(b"\n")
    for i in range(0, len(lines) - 1, 2):
        head = lines[i][1:].split(None, 1)
        out.append((head[0].decode(), head[1].decode() if len(head) > 1 else "", lines[i + 1]))
    return out


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["records"] = parse_fasta_bytes(g["fasta"].tobytes())
    g["similarity"] = float(g["similarity"])
    g["relax"] = bool(int(g["relax"]))
    g["single_end"] = bool(int(g["single_end"]))
    g["iters"] = int(g["iters"])
    g["aligned"] = int(g["aligned"])
    if g["single_end"]:
        g["reads2"] = None
    g["uniq_seq"] = [s.encode() if isinstance(s, str) else bytes(s) for s in g["uniq_seq"].tolist()]
    return g


def uniq_overlaps(g, i):
    return g["uniq_ov"][g["uniq_ptr"][i]:g["uniq_ptr"][i + 1]]


def frag_rows(g, i):
    return g["frag_as"][g["frag_ptr"][i]:g["frag_ptr"][i + 1]]


def load_hla_scale():
    """Outputs of the unmodified reference on the bench configuration (tests/golden/make_golden_hla.py); the 30,000-allele
    reference itself is regenerated from bench.make_workload(n_pairs, seed)."""
    z = np.load(os.path.join(GOLDEN, "hla_scale", "reference_outputs.npz"))
    g = {k: z[k] for k in z.files}
    g["uniq_seq"] = [s.encode() if isinstance(s, str) else bytes(s) for s in g["uniq_seq"].tolist()]
    for k in ("n_pairs", "seed", "iters", "aligned", "n_groups", "n_ec"):
        g[k] = int(g[k])
    return g
