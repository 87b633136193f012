"""CPU model check of the packed FPS chain's position recovery (csrc/fps.cu, fps_chain_packed / ScanOrder).

The kernel tracks only the VALUE of a thread's running maximum and recovers its position afterwards: first group of
scan-order neighbours whose maximum equals the thread's maximum, then the first equal leaf inside it; the tie-break
word is OR-ed together from a per-thread part, a per-group constant and a per-leaf constant.  This test restates that
logic in Python for every (points/thread, threads) pair the library instantiates and requires, on tie-heavy inputs with
every padding pattern, the very key the plain chain's strict '>' scan in the reference's order produces
(tf_sampling_g.cu:131-160: slot k mod 512 ascending, then k ascending).  The kernel itself is checked on the GPU
(tests/test_parity_gpu.py, forced chains -1 / -2)."""
import numpy as np
import pytest

PACKED_INSTANCES = [(8, 128), (16, 128), (32, 128), (8, 256), (16, 256), (32, 256), (8, 512), (16, 512), (8, 1024)]


def tb_encode(k):
    return (((k & 511) << 23) | (k >> 9)) & 0xFFFFFFFF


def tb_decode(tb):
    return (((tb & 0x7FFFFF) << 9) | (tb >> 23)) & 0xFFFFFFFF


class ScanOrder:
    def __init__(self, p, t):
        self.P, self.T = p, t
        d = 1 if t >= 512 else 512 // t
        self.DD = d if d < p else p
        self.Q = p // self.DD
        self.GS = 4 if self.Q >= 4 else self.Q
        self.G = p // self.GS

    def j_of(self, e):
        return (e % self.Q) * self.DD + e // self.Q

    def tbj(self, e):
        jt = self.j_of(e) * self.T
        return (((jt & 511) << 23) | (jt >> 9)) & 0xFFFFFFFF


@pytest.mark.parametrize("p,t", PACKED_INSTANCES)
def test_scan_order_is_the_reference_tie_break_order_and_constants_separate(p, t):
    so = ScanOrder(p, t)
    # the order fps_step visits a thread's points in: slot residue ascending, then k ascending
    classic = [j for r in range(so.DD) for j in range(r, p, so.DD)]
    assert classic == [so.j_of(e) for e in range(p)]
    for tid in (0, 1, t // 2, t - 1):
        words = [tb_encode(tid + so.j_of(e) * t) for e in range(p)]
        assert words == sorted(words), "scan order must be ascending in the tie-break word"
        for a in range(so.G):
            for u in range(so.GS):
                e = so.GS * a + u
                assert so.tbj(e) == so.tbj(so.GS * a) | so.tbj(u)
                assert u == 0 or so.tbj(so.GS * a) & so.tbj(u) == 0
                assert tb_encode(tid + so.j_of(e) * t) == tb_encode(tid) | so.tbj(e)
                assert tb_encode(tid) & so.tbj(e) == 0
                assert tb_decode(tb_encode(tid) | so.tbj(e)) == tid + so.j_of(e) * t


@pytest.mark.parametrize("p,t", PACKED_INSTANCES)
def test_grouped_equality_search_selects_what_the_strict_scan_selects(p, t):
    so = ScanOrder(p, t)
    rs = np.random.RandomState(p * 4096 + t)
    levels = np.array([0.0, 0.5, 1.0, 2.0, 1e38], np.float32)
    for _ in range(1500):
        tid = int(rs.randint(0, t))
        nreal = int(rs.randint(0, p + 1))  # points j < nreal exist (k = tid + j*T < n is a prefix in j)
        td_j = np.where(np.arange(p) < nreal, rs.choice(levels, size=p), np.float32(-1)).astype(np.float32)
        # plain chain: strict '>' from -1 in scan order; (value bits, ~word) reduced as max-then-max
        best, bj = np.float32(-1), 0
        for r in range(so.DD):
            for j in range(r, p, so.DD):
                if td_j[j] > best:
                    best, bj = td_j[j], j
        want = (int(best.view(np.uint32)), tb_encode(tid + bj * t)) if best >= 0 else (0, 0xFFFFFFFF)
        # packed chain: value-only maximum with floor 0, then first group / first leaf that equals it
        td = [td_j[so.j_of(e)] for e in range(p)]
        g = [max(td[so.GS * a:so.GS * a + so.GS]) for a in range(so.G)]
        mx = max([np.float32(0)] + g)
        s = [td[so.GS * (so.G - 1) + u] for u in range(so.GS)]
        cg = so.tbj(so.GS * (so.G - 1))
        for a in range(so.G - 2, -1, -1):
            if g[a] == mx:
                for u in range(so.GS - 1):
                    s[u] = td[so.GS * a + u]
                cg = so.tbj(so.GS * a)
        cu = so.tbj(so.GS - 1)
        for u in range(so.GS - 2, -1, -1):
            if s[u] == mx:
                cu = so.tbj(u)
        tp = tb_encode(tid) if nreal > 0 else 0xFFFFFFFF
        got = (int(np.float32(mx).view(np.uint32)), tp | cg | cu)
        assert got == want, (p, t, tid, nreal, td_j.tolist())
