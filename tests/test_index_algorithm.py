"""The algorithm of csrc/index_gpu.cu restated with numpy (test infrastructure; the CUDA code itself is checked file against file in
tests/test_index_gpu.py): suffixes keyed by their first K0 symbols above min(K0, symbols left), sorted; ties refined inside their
group in rounds of KR symbols with the same end rule; rows cut into chunks by a histogram of the first symbols.  Must reproduce
the plain suffix order, in which a suffix that runs out is smaller than one that goes on."""
import numpy as np

K0, KR, KB = 5, 3, 2     # symbols of the first key / of a refinement round / of the chunking histogram (29 / 14 / 12 on the device)


def _window(t, i, k):
    """The k symbols at i packed base 4, zero padded past the end, and how many of them exist."""
    left = max(0, min(k, len(t) - i))
    v = 0
    for j in range(k):
        v = v * 4 + (int(t[i + j]) if i + j < len(t) else 0)
    return v, left


def chunked_suffix_order(t, chunk_max):
    n = len(t)
    bins = np.zeros(4 ** KB, dtype=np.int64)
    for i in range(n):
        bins[_window(t, i, KB)[0]] += 1
    order, b = [], 0
    while b < len(bins):
        b1, m = b, 0
        while b1 < len(bins) and m + bins[b1] <= chunk_max:
            m += bins[b1]; b1 += 1
        assert b1 > b, "a bin larger than a chunk"
        sel = [i for i in range(n) if b <= _window(t, i, KB)[0] < b1]                      # select pass (any order)
        keyed = sorted((_window(t, i, K0)[0] * 8 + _window(t, i, K0)[1], i) for i in sel)  # first sort
        slots = [p for _, p in keyed]; keys = [k for k, _ in keyed]
        tied = [c for c in range(len(keys)) if (c > 0 and keys[c - 1] == keys[c]) or (c + 1 < len(keys) and keys[c + 1] == keys[c])]
        gid = {c: next(d for d in range(c, -1, -1) if d == 0 or keys[d - 1] != keys[d]) for c in tied}
        depth = K0
        while tied:                                                                         # refinement rounds: slots stay put
            items = sorted(((gid[c], _window(t, slots[c] + depth, KR)[0] * 4 + _window(t, slots[c] + depth, KR)[1]), slots[c]) for c in tied)
            for c, (_, p) in zip(tied, items):
                slots[c] = p
            k2 = {c: key for c, (key, _) in zip(tied, items)}
            nxt = [c for j, c in enumerate(tied) if (j > 0 and k2[tied[j - 1]] == k2[c]) or (j + 1 < len(tied) and k2[tied[j + 1]] == k2[c])]
            heads = {}
            for j, c in enumerate(tied):
                heads[c] = c if j == 0 or k2[tied[j - 1]] != k2[c] else heads[tied[j - 1]]
            gid = {c: heads[c] for c in nxt}
            tied, depth = nxt, depth + KR
        order += slots
        b = b1
    return order


def plain_suffix_order(t):
    s = bytes(int(x) + 1 for x in t)          # byte order with "shorter prefix first" is exactly the rule
    return sorted(range(len(t)), key=lambda i: s[i:])


def test_chunked_refinement_reproduces_the_suffix_order():
    rng = np.random.default_rng(5)
    texts = [rng.integers(0, 4, size=300), np.zeros(90, dtype=np.int64), np.tile(np.array([0, 1, 2]), 40), np.array([2, 0, 3, 3, 0, 1, 0, 2, 0, 3, 3, 0, 1, 0]),
             np.concatenate([rng.integers(0, 4, size=60)] * 3 + [rng.integers(0, 2, size=80)])]
    for t in texts:
        want = plain_suffix_order(t)
        for chunk_max in (len(t), max(int(np.bincount([_window(t, i, KB)[0] for i in range(len(t))]).max()), len(t) // 3)):
            assert chunked_suffix_order(t, chunk_max) == want
