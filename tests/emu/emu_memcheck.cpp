// emu_memcheck.cpp -- the emulated kernels under AddressSanitizer + UBSan: random shards (empty,
// short and ragged lists), every block layout, both table kernels and every scan kernel, with
// all global arrays and the dynamic shared memory allocated at their exact sizes, so that an
// out-of-bounds or misaligned access of a kernel aborts the run.  Values are not checked here
// (tests/test_emu_kernels.py does that); this is the CPU stand-in for compute-sanitizer memcheck.
#include "emu_pq.cpp"
#include <random>

int main() {
    std::mt19937 rng(7);
    auto rnd = [&](int lo, int hi) { return lo + (int)(rng() % (unsigned)(hi - lo + 1)); };
    const int d = 128, M = 32, ksub = 256, dsub = 4;
    long launches = 0;
    for (int trial = 0; trial < 8; trial++) {
        const int nlist = rnd(3, 9);
        std::vector<int64_t> list_off(nlist + 1, 0), blk_off(nlist + 1, 0);
        for (int l = 0; l < nlist; l++) {
            const int len = (trial == 0 && l == 1) ? 0 : rnd(0, 3) == 0 ? rnd(0, 5) : rnd(20, 400);
            list_off[l + 1] = list_off[l] + len;
            blk_off[l + 1] = blk_off[l] + (len + 31) / 32;
        }
        const int64_t n = list_off[nlist], nblk = blk_off[nlist];
        std::vector<uint8_t> codes((size_t)n * 32), il_codes((size_t)nblk * 1024), codes2((size_t)n * 32);
        std::vector<float> tvals(n), il_tvals((size_t)nblk * 32), tvals2(n);
        std::vector<int32_t> ids(n), il_ids((size_t)nblk * 32), ids2(n);
        for (auto& c : codes) c = (uint8_t)rng();
        for (auto& t : tvals) t = (float)(rng() % 1000) * 0.01f;
        for (int64_t i = 0; i < n; i++) ids[i] = (int32_t)i;
        std::vector<float> cb((size_t)M * ksub * dsub), cent((size_t)nlist * d);
        for (auto& v : cb) v = (float)((int)(rng() % 2001) - 1000) * 1e-3f;
        for (auto& v : cent) v = (float)((int)(rng() % 2001) - 1000) * 1e-3f;
        const int nq = rnd(1, 11), nprobe = rnd(1, nlist);
        std::vector<float> Q((size_t)nq * d);
        for (auto& v : Q) v = (float)((int)(rng() % 2001) - 1000) * 1e-3f;
        std::vector<int32_t> keys((size_t)nq * nprobe);
        for (int q = 0; q < nq; q++)
            for (int p = 0; p < nprobe; p++) keys[(size_t)q * nprobe + p] = (p == 1 && trial == 2) ? -1 : (int32_t)((q + p) % nlist);
        emu_set_seed(trial % 2 ? 0x9e3779b97f4a7c15ull + trial : 0);
        for (int layout = 1; layout <= 3; layout++) {
            if (nblk > 0) {
                emu_rm_to_il(layout, nlist, list_off.data(), blk_off.data(), codes.data(), tvals.data(), ids.data(), nblk,
                             il_codes.data(), il_tvals.data(), il_ids.data());
                emu_il_to_rm(layout, nlist, list_off.data(), blk_off.data(), il_codes.data(), il_tvals.data(),
                             il_ids.data(), nblk, codes2.data(), tvals2.data(), ids2.data());
                if (codes2 != codes || ids2 != ids) { fprintf(stderr, "layout %d round trip failed\n", layout); return 1; }
                launches += 2;
            }
            const int wide = layout == 2;
            std::vector<float> lut((size_t)nq * 256 * (wide ? 64 : 32)), dis0((size_t)nq * nprobe);
            if (trial % 2 == 0)
                emu_pq_prep(Q.data(), nq, d, M, ksub, dsub, cb.data(), cent.data(), keys.data(), nprobe, lut.data(),
                            dis0.data(), wide ? 2 : 1);
            else
                emu_pq_prep2(Q.data(), nq, d, cb.data(), cent.data(), keys.data(), nprobe, lut.data(), dis0.data(), wide);
            launches++;
            for (int k : {1, 10, 32, 33, 100}) {
                int G = rnd(1, std::min(nprobe, 16));
                const int ngroups = (nprobe + G - 1) / G;
                int KP = 32;
                while (KP < k) KP *= 2;
                const int cap = 2 * KP;
                std::vector<uint64_t> part((size_t)nq * ngroups * k);
                if (nblk == 0) continue;
                if (layout == 2)
                    emu_scan_v3(lut.data(), dis0.data(), keys.data(), nq, nprobe, G, ngroups, blk_off.data(),
                                il_codes.data(), il_tvals.data(), il_ids.data(), k, cap, part.data());
                else
                    emu_scan_v2(lut.data(), dis0.data(), keys.data(), nq, nprobe, G, ngroups, blk_off.data(),
                                il_codes.data(), il_tvals.data(), il_ids.data(), k, cap, part.data(), layout == 3);
                launches++;
            }
        }
    }
    printf("memcheck ok: %ld kernel launches\n", launches);
    return 0;
}
