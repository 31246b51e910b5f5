"""numpy model of the dictionary suffix sort of pfp_pfbwt.cu (CPU; the CUDA kernels follow it step
by step): prefix doubling where every suffix ENDS AT ITS WORD'S TERMINATOR -- the second key half
is the rank h bytes on, or 0 past the terminator; rank = the slot where a suffix's group starts; a
suffix leaves the active list when it is alone in its group or compared to its end.  Claim checked
here against a plain sort of the byte strings: the final order is the order of the strings, and
the final groups (equal rank) are exactly the sets of EQUAL suffixes -- what pfbwt.cpp:208-219
finds through lcp >= suffixLen."""
import numpy as np
import pytest

from test_oracle_golden import _pfbwt_golden


def doubling_model(d: np.ndarray, h0: int = 2):
    N = d.size
    term = d <= 1
    wid = np.concatenate([[0], np.cumsum(term)[:-1]])                 # terminators before t
    wend = np.flatnonzero(term)
    lim = wend[wid]
    # first round: h0 bytes, cut behind the terminator
    key = np.zeros(N, dtype=object)
    live = ~term
    for i in range(h0):
        idx = np.minimum(np.arange(N) + i, N - 1)
        c = np.where(live & (np.arange(N) + i < N), d[idx], 0)
        key = key * 256 + c
        live = live & ~(c <= 1)
    order = np.lexsort((np.arange(N), key.astype(np.float64) if h0 <= 6 else key))
    sa = np.zeros(N, dtype=np.int64)
    rank = np.zeros(N, dtype=np.int64)
    slot = np.arange(N)
    val = np.arange(N)[order]
    k = np.array([key[v] for v in val], dtype=object)
    h, rounds = h0, 1
    while True:
        flag = np.concatenate([[True], k[1:] != k[:-1]])
        head = slot[np.flatnonzero(flag)][np.cumsum(flag) - 1]         # slot where each element's group starts
        single = flag & np.concatenate([flag[1:], [True]])
        done = lim[val] - val + 1 <= h
        rank[val] = head
        leave = single | done
        sa[slot[leave]] = val[leave]
        keep = ~leave
        if not keep.any():
            return sa, rank, rounds
        slot, val, rs = slot[keep], val[keep], head[keep]
        j = val + h
        r2 = np.where(j <= lim[val], rank[np.minimum(j, N - 1)], 0)
        k2 = rs * (N + 1) + r2
        o = np.argsort(k2, kind="stable")
        val, k = val[o], k2[o]                                         # back into the same slots, sorted
        h, rounds = 2 * h, rounds + 1


@pytest.mark.parametrize("name", ["short_w4_p10", "low_complexity_w4_p11", "identical_copies_w10_p50"])
def test_doubling_with_word_bounded_keys_sorts_and_groups_the_dictionary_suffixes(name):
    c = _pfbwt_golden()[name]
    d = np.frombuffer(c["dict"], dtype=np.uint8).astype(np.int64)
    sa, rank, rounds = doubling_model(d)
    term = d <= 1
    ends = np.flatnonzero(term)
    wid = np.concatenate([[0], np.cumsum(term)[:-1]])
    strings = {t: bytes(d[t:ends[wid[t]] + 1].astype(np.uint8)) for t in range(d.size) if not term[t]}
    live = [int(t) for t in sa if not term[t]]
    assert sorted(live) == sorted(strings)                              # a permutation of the suffixes
    srt = [strings[t] for t in live]
    assert srt == sorted(srt)                                           # in the order of the strings
    for a, b in zip(live[:-1], live[1:]):                               # groups = equal strings, nothing else
        assert (rank[a] == rank[b]) == (strings[a] == strings[b])
    assert rounds <= 16
