// host_streamk.cpp -- CPU build of the matvec kernel's stream-K plan (streamk_plan.cuh) plus a replay of the
// bookkeeping the kernel derives from it (producer chunk order, consumer segment walk, fix-up contributor ranges), for
// tests/test_streamk_plan.py.  Test infrastructure only.
#include "../streamk_plan.cuh"

#include <vector>

using namespace b200q;

// Replays one launch with G CTAs over T tiles x KC chunks.  Returns 0 when every invariant holds, else a code:
//  1 chunk not covered exactly once            2 a CTA's segments do not add up to its range
//  3 the consumer walk flushes a wrong tile    4 fix-up contributor range / count / slot mismatch
//  5 a tile flagged "full" is not owned by one CTA      6 the 32-bit plan differs from the 64-bit plan
//  7 the wide reducer's CTA (contributor gf + 1 of a tile with >= 3 contributors) owns chunks outside that tile
extern "C" int streamk_check(int64_t T, int64_t KC, int64_t G_in, int64_t* n_split_tiles, int64_t* max_contrib) {
    const int64_t C = T * KC;
    int64_t G = G_in > C ? C : G_in;   // matvec_plan clamps the grid to the chunk count
    std::vector<int> covered((size_t)C, 0);
    // contributions[tile] = list of (cta, slot) that publish a partial; full_owner[tile] = cta writing y directly
    std::vector<std::vector<std::pair<int, int>>> contrib((size_t)T);
    std::vector<int> full_owner((size_t)T, -1);
    for (int64_t g = 0; g < G; g++) {
        const int64_t c0 = sk_begin(g, C, G), c1 = sk_begin(g + 1, C, G);
        if (c1 <= c0) return 2;
        const SkPlan sp = sk_plan(c0, c1, KC);
        if ((C + 1) * G < (1ll << 32)) {  // the kernel's 32-bit fast path must agree with the 64-bit definition
            const uint32_t a0 = sk_begin32((uint32_t)g, (uint32_t)C, (uint32_t)G), a1 = sk_begin32((uint32_t)g + 1, (uint32_t)C, (uint32_t)G);
            const SkPlan s32 = sk_plan32(a0, a1, (uint32_t)KC);
            if (a0 != c0 || a1 != c1 || s32.nH != sp.nH || s32.nT != sp.nT || s32.nF != sp.nF || s32.kcH != sp.kcH || s32.tH != sp.tH || s32.tT != sp.tT ||
                s32.tF != sp.tF)
                return 6;
        }
        const int n = (int)(c1 - c0);
        if (sp.nH + sp.nT + sp.nF != n || sp.nH < 0 || sp.nT < 0 || sp.nF < 0 || sp.nF % KC != 0) return 2;
        // producer order (matvec_impl.cuh chunk_vc): head, tail, full
        for (int j = 0; j < n; j++) {
            int64_t vc;
            if (j < sp.nH) vc = c0 + j;
            else if (j < sp.nH + sp.nT) vc = c0 + sp.nH + sp.nF + (j - sp.nH);
            else vc = c0 + sp.nH + (j - sp.nH - sp.nT);
            if (vc < c0 || vc >= c1) return 1;
            covered[(size_t)vc]++;
            // k-chunk index the producer uses for the activation record
            int kc;
            if (j < sp.nH) kc = sp.kcH + j;
            else if (j < sp.nH + sp.nT) kc = j - sp.nH;
            else kc = (j - sp.nH - sp.nT) % (int)KC;
            if (kc != (int)(vc % KC)) return 3;
        }
        // consumer walk (segment / tile bookkeeping of matvec_impl.cuh)
        int seg = sp.nH > 0 ? 0 : (sp.nT > 0 ? 1 : 2);
        int seg_left = seg == 0 ? sp.nH : (seg == 1 ? sp.nT : (int)KC);
        int t = seg == 0 ? sp.tH : (seg == 1 ? sp.tT : sp.tF);
        for (int j = 0; j < n; j++) {
            int64_t vc;
            if (j < sp.nH) vc = c0 + j;
            else if (j < sp.nH + sp.nT) vc = c0 + sp.nH + sp.nF + (j - sp.nH);
            else vc = c0 + sp.nH + (j - sp.nH - sp.nT);
            if ((int)(vc / KC) != t) return 3;   // the chunk being accumulated belongs to the tile that will be flushed
            if (--seg_left > 0) continue;
            if (seg == 2) {
                if (full_owner[(size_t)t] != -1) return 5;
                full_owner[(size_t)t] = (int)g;
            } else {
                contrib[(size_t)t].push_back({(int)g, seg});
            }
            if (seg == 0 && sp.nT > 0) { seg = 1; seg_left = sp.nT; t = sp.tT; }
            else if (seg != 2) { seg = 2; seg_left = (int)KC; t = sp.tF; }
            else { seg_left = (int)KC; t++; }
        }
    }
    for (int64_t c = 0; c < C; c++)
        if (covered[(size_t)c] != 1) return 1;
    int64_t nsplit = 0, maxc = 0;
    for (int64_t tq = 0; tq < T; tq++) {
        const auto& cs = contrib[(size_t)tq];
        if (cs.empty()) {
            if (full_owner[(size_t)tq] < 0) return 5;
            continue;
        }
        if (full_owner[(size_t)tq] >= 0) return 5;
        nsplit++;
        if ((int64_t)cs.size() > maxc) maxc = (int64_t)cs.size();
        // fix-up bookkeeping (matvec_impl.cuh): contributors are CTAs gf..gl, nc = gl - gf + 1, the first uses slot
        // sgf (0 when its range starts exactly at the tile, else 1 = its tail segment), all later ones slot 0 (head)
        const int64_t gf = sk_owner(tq * KC, C, G), gl = sk_owner((tq + 1) * KC - 1, C, G);
        const int nc = (int)(gl - gf + 1);
        const int sgf = (sk_begin(gf, C, G) == tq * KC) ? 0 : 1;
        if (nc != (int)cs.size()) return 4;
        for (int idx = 0; idx < nc; idx++) {
            const int g = (int)gf + idx;
            const int want_slot = (idx == 0) ? sgf : 0;
            bool found = false;
            for (auto& pr : cs)
                if (pr.first == g && pr.second == want_slot) found = true;
            if (!found) return 4;
        }
        // wide reducer (matvec_impl.cuh, MV_WIDE_MIN = 3): with three or more contributors the tile is reduced by ALL consumer
        // warps of contributor gf + 1 after its last chunk -- valid only if that CTA's whole chunk range lies inside this tile
        // (its head segment is all it owns: nothing else to compute, the weight ring is dead and can hold the partials)
        if (nc >= 3) {
            const int64_t g = gf + 1;
            const int64_t c0 = sk_begin(g, C, G), c1 = sk_begin(g + 1, C, G);
            const SkPlan sp = sk_plan(c0, c1, KC);
            if (c0 < tq * KC || c1 > (tq + 1) * KC) return 7;
            if (sp.nH != (int)(c1 - c0) || sp.nT != 0 || sp.nF != 0 || sp.tH != (int)tq) return 7;
        }
    }
    if (n_split_tiles) *n_split_tiles = nsplit;
    if (max_contrib) *max_contrib = maxc;
    return 0;
}
