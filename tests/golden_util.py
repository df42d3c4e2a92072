"""Golden fixtures: outputs of the UNMODIFIED reference (oracle/_ref) on small seeded cases, committed under
tests/golden/ together with their inputs so the checks do not depend on /root/reference, on oracle/_ref or on
numpy's random streams.  Regenerate with `python tests/golden/make_golden.py` (needs oracle/_ref)."""
from __future__ import annotations

import os
import pickle
import zlib

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

GOLDEN_CASES = {
    "pe_nw": dict(seed=3, n_pairs=600, genome_len=40000),
    "pe_ksw2": dict(seed=4, n_pairs=500, genome_len=40000, alg_ksw2=1, indel_rate=0.003),
    "se_nw": dict(seed=7, n_pairs=800, genome_len=40000, paired=0),
    "pe_multi": dict(seed=5, n_pairs=1500, genome_len=60000, contigs=3, n_rate=0.004, sv=4.0),
}


def path(name: str) -> str:
    return os.path.join(HERE, name + ".golden")


def save(name: str, case: dict, ref: dict) -> None:
    payload = dict(contigs=[(n, s.tobytes()) for n, s in case["contigs"]], seq=case["seq"].tobytes(), off=case["off"].tolist(), params=case["params"],
                   ref=dict(ref, profile=ref["profile"].astype(np.int16).tobytes(), profile_shape=ref["profile"].shape))
    with open(path(name), "wb") as fh:
        fh.write(zlib.compress(pickle.dumps(payload, protocol=4), 9))


def load(name: str):
    with open(path(name), "rb") as fh:
        payload = pickle.loads(zlib.decompress(fh.read()))
    contigs = [(n, np.frombuffer(b, dtype=np.uint8).copy()) for n, b in payload["contigs"]]
    case = dict(contigs=contigs, ref=np.concatenate([s for _, s in contigs]), seq=np.frombuffer(payload["seq"], dtype=np.uint8).copy(),
                off=np.asarray(payload["off"], dtype=np.int64), params=payload["params"])
    ref = dict(payload["ref"])
    ref["profile"] = np.frombuffer(ref["profile"], dtype=np.int16).reshape(ref.pop("profile_shape")).astype(np.int32)
    return case, ref


def load_vc(name: str):
    """(parameter sets, [(variant dicts, BlockDepthArr)]) of tests/golden/make_golden_vc.py for golden case `name`."""
    with open(path("vc_" + name), "rb") as fh:
        payload = pickle.loads(zlib.decompress(fh.read()))
    return payload["sets"], [(v, np.asarray(d, dtype=np.int32)) for v, d in payload["vc"]]


def with_mates(case):
    """Adds the r1 / r2 read matrices (fixed read length) that the FASTQ writers want to a loaded golden case."""
    off = case["off"]
    L = int(off[1] - off[0])
    assert (np.diff(off) == L).all()
    rows = case["seq"].reshape(-1, L)
    if case["params"]["paired"]:
        case["r1"], case["r2"] = rows[0::2].copy(), rows[1::2].copy()
    else:
        case["r1"] = case["r2"] = rows.copy()
    return case


def load_sam(name: str):
    """SAM records (bytes, header lines dropped) of the reference CLI for golden case `name` (tests/golden/make_golden_sam.py)."""
    with open(path("sam_" + name), "rb") as fh:
        return zlib.decompress(fh.read()).split(b"\n")


def available(name: str) -> bool:
    return os.path.exists(path(name))
