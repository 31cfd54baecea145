"""FM-index construction for synthetic references (test / bench data).

Produces the reference's `.bwt` / `.rbwt` payload bit-for-bit (checked against
`ibwa index` output in tests/test_fmbuild.py) so that synthetic indexes of any
size can be made on a box that has no reference tree and no time for BWT-SW
(~40 min for 3.1 Gbp with the reference; §8f N3).  Index
construction is NOT on the product path: the engine only ever consumes the
reference's file layout (bwtio.c:51-70, bwtmisc.c:122-144).

  build_bwt_numpy(text)  any text, CPU, prefix doubling           (tests)
  build_bwt_torch(text)  near-random text, GPU radix sort of 31-mers (bench, 3.1 Gbp)
"""
from __future__ import annotations

import numpy as np

from .bwtio import Bwt


def _suffix_array_numpy(t: np.ndarray) -> np.ndarray:
    """Suffix array of t (values 0..3) with an implicit smallest sentinel, by prefix doubling."""
    n = len(t)
    rank = t.astype(np.int64) + 1            # 0 is reserved for "past the end"
    sa = np.argsort(rank, kind="stable")
    k = 1
    while True:
        nxt = np.zeros(n, dtype=np.int64)
        nxt[:n - k] = rank[k:]
        key = rank * (n + 2) + nxt
        sa = np.argsort(key, kind="stable")
        sk = key[sa]
        newrank = np.empty(n, dtype=np.int64)
        newrank[sa] = np.concatenate([[0], np.cumsum(sk[1:] != sk[:-1])]) + 1
        rank = newrank
        if rank.max() == n:
            return sa
        k *= 2


def pack_reference_layout(b0: np.ndarray, primary: int, counts: np.ndarray) -> Bwt:
    """b0: the sentinel-free BWT string (uint8 0..3).  Lays it out as bwt_bwtupdate_core does
    (bwtmisc.c:122-144): per 128 bases 4 cumulative counts + 8 words, last block truncated,
    one trailing count block."""
    n = len(b0)
    nblk = (n + 127) // 128
    pad = np.zeros(nblk * 128, dtype=np.uint8)
    pad[:n] = b0
    sh = (30 - 2 * np.arange(16)).astype(np.uint32)
    words = (pad.reshape(-1, 16).astype(np.uint32) << sh[None, :]).sum(axis=1, dtype=np.uint64).astype(np.uint32)
    words = words.reshape(nblk, 8)
    onehot = np.zeros((nblk * 128, 4), dtype=np.uint8)
    onehot[np.arange(n), b0] = 1
    per_blk = onehot.reshape(nblk, 128, 4).sum(axis=1, dtype=np.int64)
    before = np.zeros((nblk + 1, 4), dtype=np.int64)
    np.cumsum(per_blk, axis=0, out=before[1:])
    out = np.empty((nblk, 12), dtype=np.uint32)
    out[:, :4] = before[:nblk].astype(np.uint32)
    out[:, 4:] = words
    flat = out.reshape(-1)
    n_words_bwt = (n + 15) // 16
    keep = 4 * nblk + n_words_bwt             # drop the zero words of the truncated last block
    flat = flat[:keep]
    payload = np.concatenate([flat, before[nblk].astype(np.uint32)])
    L2 = np.zeros(5, dtype=np.uint32)
    L2[1:] = np.cumsum(counts).astype(np.uint32)
    return Bwt(primary=int(primary), L2=L2, seq_len=n, bwt=np.ascontiguousarray(payload))


def build_bwt_numpy(text: np.ndarray) -> Bwt:
    t = np.ascontiguousarray(text, dtype=np.uint8)
    n = len(t)
    sa = _suffix_array_numpy(t)               # n suffixes; the sentinel suffix is row 0
    primary = int(np.nonzero(sa == 0)[0][0]) + 1
    prev = t[sa - 1]                          # sa == 0 wraps to t[n-1]; that row is dropped below
    keep = sa != 0
    b0 = np.concatenate([[t[n - 1]], prev[keep]]).astype(np.uint8)
    counts = np.bincount(t, minlength=4)[:4]
    return pack_reference_layout(b0, primary, counts)


def build_bwt_sa_numpy(text: np.ndarray, sa_intv: int = 32):
    """(Bwt, Sa) of a text: the .bwt payload and the sampled suffix array the reference's bwt_cal_sa /
    `bwt2sa` produce (bwt.c:48-67: SA of every sa_intv-th row, row 0 = the sentinel suffix)."""
    from .bwtio import Sa
    t = np.ascontiguousarray(text, dtype=np.uint8)
    n = len(t)
    sa = _suffix_array_numpy(t)
    primary = int(np.nonzero(sa == 0)[0][0]) + 1
    prev = t[sa - 1]
    keep = sa != 0
    b0 = np.concatenate([[t[n - 1]], prev[keep]]).astype(np.uint8)
    counts = np.bincount(t, minlength=4)[:4]
    bwt = pack_reference_layout(b0, primary, counts)
    rows = np.concatenate([[n], sa]).astype(np.uint64)          # SA over rows 0..n
    samples = rows[::sa_intv].astype(np.uint32)
    samples[0] = 0xFFFFFFFF
    return bwt, Sa(primary=primary, L2=bwt.L2.copy(), seq_len=n, sa_intv=sa_intv, sa=samples)


def build_index_numpy(text: np.ndarray):
    """(.bwt, .rbwt) of a text: the BWT of the text and of the reversed text (bwtindex.c:102-140)."""
    return build_bwt_numpy(text), build_bwt_numpy(np.ascontiguousarray(text[::-1]))


# --------------------------------------------------------------------- GPU ----

def _kmer31_chunk(t, s: int, e: int, n: int):
    """62-bit keys of the 31-mers starting at positions [s, e) (zero padded past the end)."""
    import torch
    m = e - s
    span = min(n, e + 31) - s
    x = torch.zeros(m + 31, dtype=torch.int64, device=t.device)
    x[:span] = t[s:s + span]
    k2 = (x[:-1] << 2) | x[1:]                       # length m+30
    k4 = (k2[:-2] << 4) | k2[2:]                     # m+28
    k8 = (k4[:-4] << 8) | k4[4:]                     # m+24
    k16 = (k8[:-8] << 16) | k8[8:]                   # m+16
    # 31 = 16 + 8 + 4 + 2 + 1
    key = (k16[:m] << 30) | (k8[16:16 + m] << 14) | (k4[24:24 + m] << 6) | (k2[28:28 + m] << 2) | x[30:30 + m]
    return key


def build_bwt_torch(t, bucket_bits: int = 4, chunk: int = 1 << 26, sa_intv: int = 0):
    """BWT of a near-random text on the GPU: sort suffixes by their first 31 bases (62-bit keys,
    one radix sort per leading-bases bucket), resolve the few remaining ties exactly.
    t: torch uint8 tensor (values 0..3), normally on a CUDA device."""
    import torch

    dev = t.device
    n = t.numel()
    assert 2 <= bucket_bits <= 8 and bucket_bits % 2 == 0
    k31 = torch.empty(n, dtype=torch.int64, device=dev)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        k31[s:e] = _kmer31_chunk(t, s, e, n)
    counts = torch.bincount(t[: min(n, 1 << 30)].to(torch.int64), minlength=4)[:4]
    for s in range(1 << 30, n, 1 << 30):
        counts = counts + torch.bincount(t[s:min(n, s + (1 << 30))].to(torch.int64), minlength=4)[:4]
    counts = counts.cpu().numpy()
    sa = torch.empty(n, dtype=torch.int64, device=dev)
    at = 0
    shift = 62 - bucket_bits
    for b in range(1 << bucket_bits):
        idx = torch.nonzero((k31 >> shift) == b).reshape(-1)
        if idx.numel() == 0:
            continue
        skeys, order = torch.sort(k31[idx])
        seg = idx[order]
        del idx, order
        seg = _fix_ties(t, seg, skeys, n)
        sa[at:at + seg.numel()] = seg
        at += seg.numel()
        del skeys, seg
    del k31
    assert at == n
    primary = int(torch.nonzero(sa == 0).reshape(-1)[0].item()) + 1
    sa_obj = None
    if sa_intv:                                   # sampled suffix array like bwt_cal_sa (bwt.c:48-67)
        from .bwtio import Sa
        samples = torch.cat([torch.tensor([0xFFFFFFFF], dtype=torch.int64, device=dev), sa[sa_intv - 1::sa_intv]])
        sa_obj = Sa(primary=primary, L2=None, seq_len=n, sa_intv=sa_intv,
                    sa=samples.cpu().numpy().astype(np.uint32))
    sa -= 1
    sa[primary - 1] = n - 1                       # placeholder for the dropped row
    prev = t[sa]
    keep = torch.ones(n, dtype=torch.bool, device=dev)
    keep[primary - 1] = False
    b0 = torch.cat([t[n - 1:n], prev[keep]])
    del sa, prev, keep
    bwt = _pack_reference_layout_torch(b0, primary, counts)
    if sa_obj is not None:
        sa_obj.L2 = bwt.L2.copy()
        return bwt, sa_obj
    return bwt


def _kmer31_at(t, pos, n: int, chunk: int = 1 << 24):
    """62-bit keys of the 31-mers starting at arbitrary positions `pos` (int64 tensor, all pos + 31 <= n)."""
    import torch
    out = torch.empty(pos.numel(), dtype=torch.int64, device=t.device)
    ar = torch.arange(31, device=t.device)
    sh = (2 * (30 - ar)).to(torch.int64)
    for s in range(0, pos.numel(), chunk):
        p = pos[s:s + chunk]
        out[s:s + chunk] = (t[p[:, None] + ar[None, :]].to(torch.int64) << sh[None, :]).sum(dim=1)
    return out


def _fix_ties_exact(t, seg, a: int, b: int, n: int):
    """Sorts seg[a:b] (suffix starts) exactly on the host: for the handful of runs the vector path leaves."""
    import torch

    def suffix_key(s):
        # +1 so that the end of the text (shorter suffix) sorts first
        return bytes((t[s:min(n, s + 65536)] + 1).cpu().numpy())

    members = [int(x) for x in seg[a:b].cpu().numpy()]
    keyed = sorted((suffix_key(s), s) for s in members)
    for (ka, _), (kb, _) in zip(keyed, keyed[1:]):
        assert ka != kb, "text too repetitive for the 31-mer builder; use build_bwt_numpy"
    seg[a:b] = torch.tensor([s for _, s in keyed], dtype=seg.dtype, device=seg.device)


def _fix_ties(t, seg, skeys, n, max_depth: int = 31 * 4096):
    """Orders the members of every run of equal 31-mer keys.  On the device: the tied suffixes are compared on
    their next 31 bases (one stable sort by key, one by run), the runs that are still tied go one more round, 31
    bases deeper — repeats resolve within a few rounds of their length / 31.  Runs that touch the end of the text
    (a suffix shorter than the compared depth sorts first) and whatever is left at max_depth are ordered exactly
    on the host, a handful of suffixes."""
    import torch
    m = seg.numel()
    if m < 2:
        return seg
    dev = seg.device
    start = torch.ones(m, dtype=torch.bool, device=dev)
    start[1:] = skeys[1:] != skeys[:-1]
    gid = torch.cumsum(start.to(torch.int64), 0) - 1            # run number of every element
    tie = ~start
    in_tie = tie.clone()
    in_tie[:-1] |= tie[1:]
    near_end = seg >= n - 31                                     # zero-padded keys: equal to a real key of A's
    host_runs = set(int(g) for g in gid[near_end].cpu().numpy())
    # a run with a near-end suffix must also pull in the run whose key equals the padded key: same gid by construction
    P = torch.nonzero(in_tie).reshape(-1)                        # slots of seg that take part, ascending
    depth = 31
    while P.numel() > 0:
        pos = seg[P]
        g = gid[P]
        over = pos + depth + 31 > n                              # would read past the end: to the host path
        if bool(over.any()) or depth >= max_depth:
            bad = torch.unique(g if depth >= max_depth else g[over])
            host_runs.update(int(x) for x in bad.cpu().numpy())
            keep = ~torch.isin(g, bad)
            P, pos, g = P[keep], pos[keep], g[keep]
            if P.numel() == 0:
                break
        key = _kmer31_at(t, pos + depth, n)
        o1 = torch.argsort(key, stable=True)
        o2 = torch.argsort(g[o1], stable=True)
        perm = o1[o2]                                            # by (run, next 31 bases)
        pos_s, key_s, g_s = pos[perm], key[perm], g[perm]        # g_s == g: P is grouped by run already
        seg[P] = pos_s
        same = (g_s[1:] == g_s[:-1]) & (key_s[1:] == key_s[:-1])
        still = torch.zeros(P.numel(), dtype=torch.bool, device=dev)
        still[1:] |= same
        still[:-1] |= same
        # new run numbers: a run splits wherever the next 31 bases differ
        brk = torch.ones(P.numel(), dtype=torch.bool, device=dev)
        brk[1:] = ~same
        new_g = torch.cumsum(brk.to(torch.int64), 0) - 1
        keep_host = torch.isin(g_s, torch.tensor(sorted(host_runs), dtype=torch.int64, device=dev)) if host_runs else None
        gid_new = new_g + (int(gid.max().item()) + 1)            # fresh numbers, disjoint from the old ones
        if keep_host is not None and bool(keep_host.any()):
            gid_new = torch.where(keep_host, g_s, gid_new)       # runs destined for the host keep their identity
            still = still & ~keep_host
        gid[P] = gid_new
        P = P[still]
        depth += 31
    if host_runs:
        hr = torch.tensor(sorted(host_runs), dtype=torch.int64, device=dev)
        sel = torch.nonzero(torch.isin(gid, hr)).reshape(-1)
        if sel.numel():
            gs = gid[sel].cpu().numpy()
            sl = sel.cpu().numpy()
            # the slots of one run are contiguous
            cut = np.nonzero(np.diff(gs) != 0)[0] + 1
            for grp in np.split(np.arange(len(sl)), cut):
                a, b = int(sl[grp[0]]), int(sl[grp[-1]]) + 1
                if b - a >= 2:
                    _fix_ties_exact(t, seg, a, b, n)
    return seg


def _pack_reference_layout_torch(b0, primary: int, counts) -> Bwt:
    import torch
    n = b0.numel()
    dev = b0.device
    nblk = (n + 127) // 128
    pad = torch.zeros(nblk * 128, dtype=torch.uint8, device=dev)
    pad[:n] = b0
    del b0
    sh = (30 - 2 * torch.arange(16, device=dev, dtype=torch.int64))
    words = torch.empty(nblk * 8, dtype=torch.int64, device=dev)
    CH = 1 << 23
    v = pad.reshape(-1, 16)
    for s in range(0, nblk * 8, CH):
        e = min(nblk * 8, s + CH)
        words[s:e] = (v[s:e].to(torch.int64) << sh[None, :]).sum(dim=1)
    words = words.reshape(nblk, 8)
    per_blk = torch.empty((nblk, 4), dtype=torch.int64, device=dev)
    pb = pad.reshape(nblk, 128)
    CB = 1 << 21
    for s in range(0, nblk, CB):
        e = min(nblk, s + CB)
        for c in range(4):
            per_blk[s:e, c] = (pb[s:e] == c).sum(dim=1)
    per_blk[nblk - 1, 0] -= nblk * 128 - n          # zero padding of the last block is not text
    before = torch.zeros((nblk + 1, 4), dtype=torch.int64, device=dev)
    before[1:] = torch.cumsum(per_blk, dim=0)
    out = torch.empty((nblk, 12), dtype=torch.int64, device=dev)
    out[:, :4] = before[:nblk]
    out[:, 4:] = words
    flat = out.reshape(-1)
    keep = 4 * nblk + (n + 15) // 16
    payload = torch.cat([flat[:keep], before[nblk]]).to(torch.int32)   # wraps to the uint32 bit pattern
    host = payload.cpu().numpy().view(np.uint32)
    L2 = np.zeros(5, dtype=np.uint32)
    L2[1:] = np.cumsum(np.asarray(counts, dtype=np.int64)).astype(np.uint32)
    return Bwt(primary=int(primary), L2=L2, seq_len=n, bwt=np.ascontiguousarray(host))


def build_index_torch(t):
    """(.bwt, .rbwt) on the GPU for a near-random text (torch uint8 CUDA tensor)."""
    import torch
    fwd = build_bwt_torch(t)
    rev = build_bwt_torch(torch.flip(t, dims=[0]))
    return fwd, rev
