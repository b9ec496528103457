"""The index arithmetic of the peer-memory all-reduce (sigma/diffsharp_b200/csrc/dsc_comm.cu, allreduce_p2p_kernel),
restated with NumPy and run for every rank count 2..8 and ragged buffer lists — the GPU tests cover 2 and 8 ranks.

Restated here: the 16-byte aligned packed layout (P2PTable), the slice owned by each rank, the scatter of slice q
into region A of rank q at slot [source rank] (with the rotated walk order), the rank-ordered reduction of the
slots, the broadcast into region B of every rank, and the unpack.  Checked: every region word that is read was
written in this call (or is alignment padding, which is zero by construction), the result equals the float32 sum
added in rank order, bit for bit, and it is identical on every rank."""
import numpy as np
import pytest

W = 4  # float32 per 16-byte vector


def allreduce_sim(bufs_per_rank):
    nranks = len(bufs_per_rank)
    lens = [len(b) for b in bufs_per_rank[0]]
    start = [0]
    for n in lens:
        start.append(start[-1] + (n + W - 1) // W * W)
    total = start[-1]
    nvec = total // W
    per = (nvec + nranks - 1) // nranks
    region_a = [np.zeros(nranks * per * W, np.float32) for _ in range(nranks)]      # zeroed once at setup
    written_a = [np.zeros(nranks * per * W, bool) for _ in range(nranks)]
    region_b = [np.zeros(max(total, nranks * per * W), np.float32) for _ in range(nranks)]
    written_b = [np.zeros_like(region_b[0], dtype=bool) for _ in range(nranks)]
    # 0. scatter (the walk order is rotated per rank; the destination of an element does not depend on it)
    for rank in range(nranks):
        for s, n in enumerate(lens):
            p = bufs_per_rank[rank][s]
            v_seg = start[s] // W
            nv = n // W
            rot = ((rank + 1) % nranks) * per - v_seg
            rot = 0 if rot < 0 or rot >= nv else rot
            seen = np.zeros(nv, bool)
            for j in range(nv):
                i = j + rot if j + rot < nv else j + rot - nv
                assert not seen[i]
                seen[i] = True
                v = v_seg + i
                q = v // per
                assert q < nranks
                dst = (rank * per + (v - q * per)) * W
                region_a[q][dst:dst + W] = p[i * W:(i + 1) * W]
                written_a[q][dst:dst + W] = True
            assert seen.all()
            for i in range(nv * W, n):   # tail of the segment, element by element
                e = start[s] + i
                v = e // W
                q = v // per
                dst = (rank * per + (v - q * per)) * W + (e - v * W)
                region_a[q][dst] = p[i]
                written_a[q][dst] = True
    # 2. reduce this rank's slice in rank order, broadcast to region B of every rank
    valid = np.zeros(total, bool)
    for s, n in enumerate(lens):
        valid[start[s]:start[s] + n] = True
    for rank in range(nranks):
        v0 = per * rank
        mine = max(0, min(per, nvec - v0))
        for i in range(mine):
            acc = None
            for r in range(nranks):
                x = region_a[rank][(r * per + i) * W:(r * per + i + 1) * W]
                wr = written_a[rank][(r * per + i) * W:(r * per + i + 1) * W]
                # a word that was not written in this call must be alignment padding (never written: stays zero)
                assert np.all(wr | ~valid[(v0 + i) * W:(v0 + i + 1) * W])
                acc = x.copy() if acc is None else (acc + x).astype(np.float32)
            for q in range(nranks):
                region_b[q][(v0 + i) * W:(v0 + i + 1) * W] = acc
                written_b[q][(v0 + i) * W:(v0 + i + 1) * W] = True
    # 4. unpack
    out = []
    for rank in range(nranks):
        res = []
        for s, n in enumerate(lens):
            assert written_b[rank][start[s]:start[s] + n].all()
            res.append(region_b[rank][start[s]:start[s] + n].copy())
        out.append(res)
    return out


@pytest.mark.parametrize("nranks", [2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("lens", [(1000, 37, 4096 * 3 + 1), (1,), (3, 3, 3), (16,), (5, 1024, 2, 777), (4, 8, 12)])
def test_layout_and_order(nranks, lens):
    rng = np.random.default_rng(nranks * 100 + len(lens))
    bufs = [[rng.standard_normal(n).astype(np.float32) for n in lens] for _ in range(nranks)]
    out = allreduce_sim(bufs)
    for s in range(len(lens)):
        want = bufs[0][s].copy()
        for r in range(1, nranks):
            want = (want + bufs[r][s]).astype(np.float32)
        for rank in range(nranks):
            assert np.array_equal(out[rank][s], want)


# ---- the flag-in-data kernel (allreduce_ll_kernel): 128-byte lines of 15 payload words + 1 flag word ----------------------------
def allreduce_ll_sim(bufs_per_rank, dtype, seq=7, stale_seq=6):
    """The line arithmetic of allreduce_ll_kernel restated: the packed space in 8-byte words (2 float32 / 1 float64), lines of
    15 payload words, slices of whole lines, the per-lane word assignment (lane group of 8: lanes 0..6 carry two payload words,
    lane 7 one payload word and the flag), the rotated walk, region A slots, the rank-ordered sum and region B.  The regions
    start out holding lines of an EARLIER call (flag = stale_seq): a reader may only use a line whose flag is `seq`."""
    nranks = len(bufs_per_rank)
    epw = 8 // np.dtype(dtype).itemsize
    w16 = 16 // np.dtype(dtype).itemsize
    lens = [len(b) for b in bufs_per_rank[0]]
    start = [0]
    for n in lens:
        start.append(start[-1] + (n + w16 - 1) // w16 * w16)
    total_e = start[-1]
    total_words = total_e // epw
    TL = (total_words + 14) // 15
    lpr = (TL + nranks - 1) // nranks
    span = lpr * nranks

    def gather(rank, w):          # packed word w of this rank's buffers as `epw` elements (zeros in gaps / past the end)
        out = np.zeros(epw, dtype)
        for k in range(epw):
            e = w * epw + k
            if e >= total_e:
                continue
            s = max(i for i in range(len(lens)) if start[i] <= e)
            i = e - start[s]
            if i < lens[s]:
                out[k] = bufs_per_rank[rank][s][i]
        return out

    # a line = 16 words: payload words 0..14 as elements, word 15 = flag
    def new_region(lines):
        return np.full((lines, 15 * epw), np.nan, dtype), np.full(lines, stale_seq, np.int64)
    region_a = [new_region(span) for _ in range(nranks)]
    region_b = [new_region(max(TL, 1)) for _ in range(nranks)]
    # 0. scatter
    for rank in range(nranks):
        seen = np.zeros(span, bool)
        for gi in range(span):
            g = gi + ((rank + 1) % nranks) * lpr
            if g >= span:
                g -= span
            assert not seen[g]
            seen[g] = True
            if g >= TL:
                continue
            q = g // lpr
            line = np.concatenate([gather(rank, g * 15 + sub * 2 + k) for sub in range(8) for k in range(2) if not (sub == 7 and k == 1)])
            data, flags = region_a[q]
            slot_line = rank * lpr + (g - q * lpr)
            assert flags[slot_line] == stale_seq   # every slot line is written once per call
            data[slot_line] = line
            flags[slot_line] = seq
        assert seen.all()
    # 2. reduce the own slice in rank order, broadcast
    for rank in range(nranks):
        first = lpr * rank
        mine = max(0, min(lpr, TL - first))
        data, flags = region_a[rank]
        for loc in range(mine):
            acc = None
            for r in range(nranks):
                assert flags[r * lpr + loc] == seq           # the line every reader waits for does arrive
                x = data[r * lpr + loc]
                acc = x.copy() if acc is None else (acc + x).astype(dtype)
            for q in range(nranks):
                bd, bf = region_b[q]
                assert bf[first + loc] == stale_seq
                bd[first + loc] = acc
                bf[first + loc] = seq
    # 4. unpack
    out = []
    for rank in range(nranks):
        bd, bf = region_b[rank]
        res = [np.zeros(n, dtype) for n in lens]
        for g in range(TL):
            assert bf[g] == seq
            for wi in range(15):
                for k in range(epw):
                    e = (g * 15 + wi) * epw + k
                    if e >= total_e:
                        continue
                    s = max(i for i in range(len(lens)) if start[i] <= e)
                    i = e - start[s]
                    if i < lens[s]:
                        res[s][i] = bd[g][wi * epw + k]
        out.append(res)
    region_bytes_needed = span * 128
    return out, region_bytes_needed, total_e * np.dtype(dtype).itemsize


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("nranks", [2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("lens", [(1000, 37, 4096 + 1), (1,), (3, 3, 3), (16,), (5, 1024, 2, 777), (30, 60), (29, 31, 1)])
def test_flag_in_data_layout_and_order(dtype, nranks, lens):
    rng = np.random.default_rng(nranks * 1000 + sum(lens))
    bufs = [[rng.standard_normal(n).astype(dtype) for n in lens] for _ in range(nranks)]
    out, need, payload = allreduce_ll_sim(bufs, dtype)
    for s in range(len(lens)):
        want = bufs[0][s].copy()
        for r in range(1, nranks):
            want = (want + bufs[r][s]).astype(dtype)
        for rank in range(nranks):
            assert np.array_equal(out[rank][s], want)
    # the capacity test of the library (p2p_fits): payload + payload / 15 + 128 * (max ranks + 2) covers the slot lines
    assert need <= payload + payload // 15 + 128 * (8 + 2)
