"""Lane-level numpy emulation of scan_pq_il2_kernel (distributed_faiss_b200/csrc/dfx_scan_il2.cu).

Test infrastructure: the kernel cannot run without a GPU, so its ALGORITHM -- block layout 2, wide
table, rotated lane-per-vector lookups, the halving tree over the rotated index, the
register-resident top-k with its shuffle networks, the lazy id gather, the CTA-shared bound and
the final merge -- is restated here one warp-instruction at a time (arrays of 32 lanes) and
compared with the oracle on CPU.  The block layout comes from the library itself
(dfx_debug_il_byte), everything else mirrors the CUDA source statement by statement.
"""
import ctypes

import numpy as np

F = np.float32
NONE = np.uint64(0xFFFFFFFFFFFFFFFF)
LANES = np.arange(32)
NW = 8
QCAP = 64
INF = F(np.inf)


def f2key(v):
    v = (np.asarray(v, dtype=F) + F(0.0)).astype(F)
    u = v.view(np.uint32)
    return np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000)).astype(np.uint32)


def key2f(k):
    k = np.asarray(k, dtype=np.uint32)
    u = np.where(k & np.uint32(0x80000000), k & np.uint32(0x7FFFFFFF), ~k).astype(np.uint32)
    return u.view(F)


def il_byte_table(lib, layout):
    """[v][m] -> byte offset inside the 1 KB block, from the library's own layout function"""
    lib.dfx_debug_il_byte.restype = ctypes.c_int
    return np.array([[lib.dfx_debug_il_byte(layout, v, m) for m in range(32)] for v in range(32)], dtype=np.int64)


def rm_to_il(byte_of, list_off, codes, tvals, ids):
    """pq_rm_to_il_kernel: row-major list-sorted arrays -> padded 32-vector blocks"""
    nlist = len(list_off) - 1
    blk_off = np.zeros(nlist + 1, dtype=np.int64)
    for l in range(nlist):
        blk_off[l + 1] = blk_off[l] + (list_off[l + 1] - list_off[l] + 31) // 32
    nblk = int(blk_off[-1])
    il_codes = np.zeros((max(nblk, 1), 1024), dtype=np.uint8)
    il_tvals = np.full((max(nblk, 1), 32), np.inf, dtype=F)
    il_ids = np.full((max(nblk, 1), 32), -1, dtype=np.int32)
    for l in range(nlist):
        for b in range(int(blk_off[l]), int(blk_off[l + 1])):
            base = int(list_off[l]) + (b - int(blk_off[l])) * 32
            n = min(32, int(list_off[l + 1]) - base)
            for v in range(n):
                il_codes[b, byte_of[v]] = codes[base + v]
            il_tvals[b, :n] = tvals[base:base + n]
            il_ids[b, :n] = ids[base:base + n]
    return blk_off, il_codes, il_tvals, il_ids


def wide_table(lut):
    """pq_prep_kernel mode 2: lut [32][256] -> [256][64], column c holds m = c & 31"""
    return np.ascontiguousarray(lut.T[:, np.arange(64) & 31], dtype=F)


# ------------------------------------------------------------------ warp primitives
def shfl_xor(x, s):
    return x[LANES ^ s]


def sort32_asc(x):
    for size in (2, 4, 8, 16, 32):
        stride = size >> 1
        while stride >= 1:
            o = shfl_xor(x, stride)
            up = (LANES & size) == 0
            lower = (LANES & stride) == 0
            mn, mx = np.minimum(x, o), np.maximum(x, o)
            x = np.where(lower == up, mn, mx)
            stride >>= 1
    return x


def merge_sorted(kept, x_asc):
    xr = x_asc[31 - LANES]
    y = np.minimum(kept, xr)
    for stride in (16, 8, 4, 2, 1):
        o = shfl_xor(y, stride)
        lower = (LANES & stride) == 0
        y = np.where(lower, np.minimum(y, o), np.maximum(y, o))
    return y


def block_values(lutW, block_codes, tv, d0):
    """the 32 lookups + halving tree of one block; returns v[lane] (fp32, kernel order)"""
    y = np.empty((32, 32), dtype=F)  # [t][lane]
    for t in range(32):
        code = block_codes[(t >> 4) * 512 + LANES * 16 + (t & 15)].astype(np.int64)
        y[t] = lutW[code, LANES + t]
    lo = y[0::2].copy()  # p2[i].lo = y[2i], p2[i].hi = y[2i+1]
    hi = y[1::2].copy()
    off = 8
    while off >= 1:
        for i in range(off):
            lo[i] = (lo[i] + lo[i + off]).astype(F)
            hi[i] = (hi[i] + hi[i + off]).astype(F)
        off >>= 1
    s = (lo[0] + hi[0]).astype(F)
    return (F(d0) + (tv + s).astype(F)).astype(F)


class Warp:
    def __init__(self, cta, warp, k):
        self.cta, self.warp, self.k = cta, warp, k
        self.kept = np.full(32, NONE, dtype=np.uint64)
        self.queue = np.zeros(QCAP, dtype=np.uint64)
        self.cnt = 0
        self.thr, self.thr_sec, self.bnd = INF, np.uint32(0xFFFFFFFF), INF
        self.flushes = 0

    def flush(self):
        """il2_flush: merge the first min(cnt, 32) entries, move the rest to the front"""
        cta = self.cta
        n = min(self.cnt, 32)
        x = np.full(32, NONE, dtype=np.uint64)
        live = LANES < n
        c = self.queue[LANES]
        ids = cta.il_ids.reshape(-1)[(c & np.uint64(0xFFFFFFFF)).astype(np.int64) % cta.il_ids.size].astype(np.int64)
        comp = (c & np.uint64(0xFFFFFFFF00000000)) | (ids & 0xFFFFFFFF).astype(np.uint64)
        x = np.where(live, comp, x)
        rest = self.cnt - n
        self.queue[:rest] = self.queue[32:32 + rest].copy()
        self.kept = merge_sorted(self.kept, sort32_asc(x))
        self.cnt = rest
        self.flushes += 1
        kth = self.kept[self.k - 1]
        if kth != NONE:
            self.thr = key2f(np.uint32(kth >> np.uint64(32)))[()]
            self.thr_sec = np.uint32(kth & np.uint64(0xFFFFFFFF))
            cta.cta_key = min(cta.cta_key, int(kth >> np.uint64(32)))

    def process(self, pos, d0):
        cta = self.cta
        v = block_values(cta.lutW, cta.il_codes[pos], cta.il_tvals[pos], d0)
        with np.errstate(invalid="ignore"):
            pas = v <= self.bnd
        if not pas.any():
            return
        self.bnd = key2f(np.uint32(cta.cta_key))[()]
        with np.errstate(invalid="ignore"):
            pas = v <= self.bnd
            ids = cta.il_ids[pos].astype(np.int64).astype(np.uint32)
            want = pas & ((v < self.thr) | ((v == self.thr) & (ids < self.thr_sec)))
        if not want.any():
            return
        slot = self.cnt + np.cumsum(want) - want
        comp = (f2key(v).astype(np.uint64) << np.uint64(32)) | (np.uint64(pos) * np.uint64(32) + LANES.astype(np.uint64))
        self.queue[slot[want]] = comp[want]
        self.cnt += int(want.sum())
        if self.cnt >= 32:
            self.flush()

    def stream(self):
        """generator over this warp's blocks: (pos, d0), lists in probe order, blocks lb+warp, +NW, ..."""
        for lb, le, d0 in self.cta.lists:
            for b in range(lb + self.warp, le, NW):
                yield b, d0


class Cta:
    """one (query, probe group): emulates scan_pq_il2_kernel<REG = true>"""

    def __init__(self, lutW, lists, il_codes, il_tvals, il_ids, k, rng=None):
        assert 1 <= k <= 32
        self.lutW, self.lists = lutW, lists
        self.il_codes, self.il_tvals, self.il_ids = il_codes, il_tvals, il_ids
        self.k = k
        self.cta_key = 0xFF800000
        self.warps = [Warp(self, w, k) for w in range(NW)]
        self.rng = rng

    def run(self):
        streams = [w.stream() for w in self.warps]
        live = list(range(NW))
        while live:  # any interleaving of the warps is a valid schedule
            i = live[0] if self.rng is None else live[self.rng.randint(len(live))]
            try:
                pos, d0 = next(streams[i])
            except StopIteration:
                live.remove(i)
                continue
            self.warps[i].process(pos, d0)
        for w in self.warps:
            if w.cnt > 0:
                w.flush()
        kept = self.warps[0].kept
        for w in self.warps[1:]:
            kept = merge_sorted(kept, w.kept)
        return kept[:self.k]


def search(orc_index, lib, xq, k, nprobe, G=None, rng=None, layout=2):
    """whole IVF-PQ search through the emulated kernel; returns (D, I) like the oracle"""
    from oracle import oracle as O

    st = orc_index.get_state()
    byte_of = il_byte_table(lib, layout)
    blk_off, il_codes, il_tvals, il_ids = rm_to_il(byte_of, st["list_off"], st["codes"], st["tvals"],
                                                   st["ids"].astype(np.int32))
    keys, _ = O.coarse(st["coarse_metric"], st["centroids"], xq, nprobe)
    nq = xq.shape[0]
    G = nprobe if G is None else G
    D = np.full((nq, k), np.inf, dtype=F)
    I = np.full((nq, k), -1, dtype=np.int64)
    stats = {"flushes": 0}
    for q in range(nq):
        lutW = wide_table(orc_index.query_lut(xq[q]))
        parts = []
        for g0 in range(0, nprobe, G):
            lists = []
            for p in range(g0, min(nprobe, g0 + G)):
                l = int(keys[q, p])
                if l < 0:
                    lists.append((0, 0, F(0)))
                else:
                    d0 = F(O.warp_dot(xq[q], st["centroids"][l], 1))
                    lists.append((int(blk_off[l]), int(blk_off[l + 1]), d0))
            cta = Cta(lutW, lists, il_codes, il_tvals, il_ids, k, rng)
            parts.append(cta.run())
            stats["flushes"] += sum(w.flushes for w in cta.warps)
        allc = np.sort(np.concatenate(parts))[:k]  # the downstream selection kernel (unchanged)
        for j, c in enumerate(allc):
            if c != NONE:
                D[q, j] = key2f(np.uint32(c >> np.uint64(32)))[()]
                I[q, j] = int(c & np.uint64(0xFFFFFFFF))
    return D, I, stats
