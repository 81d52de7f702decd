// plan.cu - host-side planning of the first-stage launches (no device code).
//
// Both 16-bit halves of the packed words must be busy for the DPX kernel to run at full rate, and
// they must see the same database residue.  So the queries are dealt onto two TRACKS (one per
// half), balanced by rows; on a track the queries follow each other, each starting on a lane
// boundary.  A pass covers the next G*R rows of both tracks.  Compared with pairing queries
// one-to-one, nothing is lost when the two queries of a pair differ in length, and short
// queries ride in the same full-width (G = 32, R = 40) launches as long ones.
#include "osw_internal.h"
#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {

const int kG[4] = {4, 8, 16, 32};
const int kNR = 10;
const int kR[kNR] = {8, 12, 16, 20, 24, 28, 32, 36, 40, 44};
const int kRmaxDefault = 40;        // rows per lane of the full-height passes

struct Track { std::vector<int> q; uint64_t rows = 0; };

// Fills one pass of one half: lanes [0,G) get the next rows of the track.  cur/done: index of the
// track's current query and rows of it already placed.  Returns lanes used.
int fill_half(const Track &tr, const uint32_t *q_len, size_t &cur, uint32_t &done, int G, int R, OswLaneDesc *lane,
              bool *continues_in, bool *continues_out) {
    int t = 0;
    *continues_in = cur < tr.q.size() && done > 0;
    while (t < G && cur < tr.q.size()) {
        const int q = tr.q[cur];
        const uint32_t len = q_len[q];
        while (t < G && done < len) {
            lane[t].query = (uint32_t)q; lane[t].q_len = len; lane[t].row0 = done;
            lane[t].flags = done == 0 ? OSW_LANE_START : 0u;
            done += (uint32_t)R;
            ++t;
        }
        if (done >= len) { lane[t - 1].flags |= OSW_LANE_EMIT; ++cur; done = 0; }
    }
    *continues_out = cur < tr.q.size() && done > 0;
    if (*continues_out) lane[G - 1].flags |= OSW_LANE_EMIT;        // partial maximum of a query that goes on
    const int used = t;
    for (; t < G; ++t) { lane[t].query = 0xffffffffu; lane[t].q_len = 0; lane[t].row0 = 0; lane[t].flags = OSW_LANE_START; }
    return used;
}

int lanes_needed(const Track &tr, const uint32_t *q_len, size_t cur, uint32_t done, int R) {
    int n = 0;
    for (size_t k = cur; k < tr.q.size(); ++k) {
        uint32_t rem = q_len[tr.q[k]] - (k == cur ? done : 0u);
        n += (int)((rem + R - 1) / R);
    }
    return n;
}

}  // namespace

// Pair-database mode keeps two score tables in shared memory: with 32 lanes at most 28 rows each.
static int pd_rmax(int G) { return G == 32 ? 28 : 40; }

extern "C" int osw_plan_passes(const uint32_t *q_len, int nq, OswPass *out, int max_passes, int mode, int min_g) {
    return osw_plan_passes_ex(q_len, nq, out, max_passes, mode, min_g, 0);
}

int osw_plan_passes_ex(const uint32_t *q_len, int nq, OswPass *out, int max_passes, int mode, int min_g, int rmax) {
    if (!q_len || nq < 1 || !out || max_passes < 1) return -1;
    const int kRmax = rmax >= 16 && rmax <= 44 && rmax % 4 == 0 ? rmax : kRmaxDefault;     // (experiments: OSW_RMAX)
    // longest first onto the lighter track (rows rounded up to whole lanes of the widest geometry)
    std::vector<int> order(nq);
    for (int i = 0; i < nq; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return q_len[a] > q_len[b]; });
    Track tr[2];
    for (int q : order) {
        if (q_len[q] == 0) continue;                   // an empty query scores 0 everywhere: nothing to run
        Track &t = tr[0].rows <= tr[1].rows ? tr[0] : tr[1];
        t.q.push_back(q);
        t.rows += (q_len[q] + kRmax - 1) / kRmax * kRmax;
    }
    // Two tracks need two comparable loads.  With a single query, or one much longer than the
    // rest together, half of every word would idle: then all queries go on ONE track and the two
    // halves score two different database sequences instead (pair-database mode).  That mode runs at
    // about 78 % of the two-track rate per cell (two table reads per row), so it pays when the
    // lighter track holds less than 56 % of the heavier one's rows.
    const bool pair_db = mode == OSW_PLAN_PAIR_DB || (mode == OSW_PLAN_AUTO && tr[1].rows * 100 < tr[0].rows * 56);
    if (pair_db) {
        tr[0].q.clear(); tr[1].q.clear();
        for (int q : order) if (q_len[q]) tr[0].q.push_back(q);
    }
    size_t cur[2] = {0, 0};
    uint32_t done[2] = {0, 0};
    int n = 0;
    // Several passes: as few as the tallest geometry allows, then the smallest R that still fits
    // that many passes (1000 rows are two passes of 32 x 16, not 32 x 28 + 32 x 16).
    const int r_cap = pair_db ? std::min(kRmax, pd_rmax(32)) : kRmax;
    int r_full = r_cap;
    {
        const int need_cap = std::max(lanes_needed(tr[0], q_len, 0, 0, r_cap), lanes_needed(tr[1], q_len, 0, 0, r_cap));
        const int n_min = (need_cap + 31) / 32;
        for (int ri = kNR - 1; ri >= 0; --ri) {
            if (kR[ri] > r_cap) continue;
            const int need = std::max(lanes_needed(tr[0], q_len, 0, 0, kR[ri]), lanes_needed(tr[1], q_len, 0, 0, kR[ri]));
            if ((need + 31) / 32 <= n_min) r_full = kR[ri];
        }
    }
    for (;;) {
        if (cur[0] >= tr[0].q.size() && cur[1] >= tr[1].q.size()) break;
        if (n >= max_passes) return -1;
        // smallest geometry that finishes both tracks in this pass, if there is one
        int G = 32, R = r_full;
        uint64_t best = ~0ull;
        for (int gi = 0; gi < 4; ++gi)
            for (int ri = 0; ri < kNR; ++ri) {
                if (kR[ri] > kRmax || (pair_db && kR[ri] > pd_rmax(kG[gi]))) continue;
                const int need = std::max(lanes_needed(tr[0], q_len, cur[0], done[0], kR[ri]),
                                          lanes_needed(tr[1], q_len, cur[1], done[1], kR[ri]));
                if (need > kG[gi] || kG[gi] < min_g) continue;
                if (n > 0 && kG[gi] != 32) continue;            // a continued query keeps the 32-lane array
                const uint64_t cost = (uint64_t)kG[gi] * (kR[ri] + 3);
                if (cost < best) { best = cost; G = kG[gi]; R = kR[ri]; }
            }
        OswPass &p = out[n];
        memset(&p, 0, sizeof p);
        p.G = G; p.R = R;
        bool in0, in1, out0, out1;
        fill_half(tr[0], q_len, cur[0], done[0], G, R, p.lane[0], &in0, &out0);
        fill_half(tr[1], q_len, cur[1], done[1], G, R, p.lane[1], &in1, &out1);
        p.has_in = in0 || in1;
        p.has_out = out0 || out1;
        p.pair_db = pair_db ? 1 : 0;
        if (pair_db) memcpy(p.lane[1], p.lane[0], sizeof p.lane[0]);     // both halves: the same query rows
        ++n;
    }
    return n;
}
