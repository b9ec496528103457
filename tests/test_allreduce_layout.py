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
