#include "../../despawnengine_b200/csrc/dsp_api.cu"
#include <chrono>
#include <random>
// CPU harness (no CUDA call is made; the library source is included as one translation unit so that its internal functions are in
// reach): resolve_positions -- the batch calls' position -> slot resolution with its "continue the previous answer" loop -- against
// acquire_slot + stamp_slot entry by entry and against find_slot, through orders, duplicates, holes and box / map boundaries; then its speed.
static dsp_ctx* fresh_ctx(bool box, bool both) {
    dsp_ctx* ctx = new dsp_ctx();
    ctx->max_chunks = 20000;
    ctx->resident.assign(20000, 0);
    if (both) {  // a few map slots first
        for (int k = 0; k < 5; ++k) { int32_t p[3] = {100 + k, 0, 0}; uint32_t s; acquire_slot(ctx, p, &s); }
    }
    if (box || both) {
        Region r{{0, 0, 0}, {16, 16, 16}, static_cast<uint32_t>(ctx->next_unused)};
        ctx->regions.push_back(r);
        std::fill(ctx->resident.begin() + ctx->next_unused, ctx->resident.begin() + ctx->next_unused + 4096, 1);
        ctx->next_unused += 4096; ctx->n_resident += 4096;
        Region r2{{-3, 5, 40}, {3, 2, 2}, static_cast<uint32_t>(ctx->next_unused)};
        ctx->regions.push_back(r2);
        std::fill(ctx->resident.begin() + ctx->next_unused, ctx->resident.begin() + ctx->next_unused + 12, 1);
        ctx->next_unused += 12; ctx->n_resident += 12;
    }
    return ctx;
}
static int compare(dsp_ctx* A, dsp_ctx* B, const std::vector<int32_t>& pos, const char* what) {
    const size_t n = pos.size() / 3;
    std::vector<uint32_t> sa(n, 0xdead), sb(n, 0xbeef), fa, fb;
    std::vector<int4> pa(n), pb(n);
    new_stamp(A); new_stamp(B);
    // A: in three ragged pieces, B: entry by entry
    int ra = DSP_OK;
    const size_t cuts[4] = {0, n / 5, n / 2 + 3 > n ? n : n / 2 + 3, n};
    for (int k = 0; k < 3 && ra == DSP_OK; ++k) ra = resolve_positions(A, pos.data(), cuts[k], cuts[k + 1], sa.data(), pa.data(), fa);
    int rb = DSP_OK;
    for (size_t i = 0; i < n && rb == DSP_OK; ++i) {
        bool fresh = false;
        rb = acquire_slot(B, &pos[3 * i], &sb[i], &fresh);
        if (rb != DSP_OK) break;
        if (fresh) fb.push_back(sb[i]);
        if (!stamp_slot(B, sb[i])) { rb = DSP_ERR_BAD_ARG; break; }
        pb[i] = make_int4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], 0);
    }
    if (ra != rb) { printf("FAIL %s: rc %d vs %d\n", what, ra, rb); return 1; }
    if (ra == DSP_OK) {
        for (size_t i = 0; i < n; ++i)
            if (sa[i] != sb[i] || pa[i].x != pb[i].x || pa[i].y != pb[i].y || pa[i].z != pb[i].z) { printf("FAIL %s: entry %zu slot %u vs %u\n", what, i, sa[i], sb[i]); return 1; }
        // and against the map / box list alone
        for (size_t i = 0; i < n; ++i) { uint32_t s; if (!find_slot(A, &pos[3 * i], &s) || s != sa[i]) { printf("FAIL %s: find_slot entry %zu\n", what, i); return 1; } }
    }
    if (fa != fb) { printf("FAIL %s: fresh lists differ\n", what); return 1; }
    printf("ok   %-44s rc %d, %zu entries, %zu fresh\n", what, ra, n, fa.size());
    return 0;
}
int main() {
    int bad = 0;
    std::mt19937 rng(5);
    for (int kind = 0; kind < 3; ++kind) {
        dsp_ctx* A = fresh_ctx(kind == 1, kind == 2);
        dsp_ctx* B = fresh_ctx(kind == 1, kind == 2);
        std::vector<int32_t> box;
        for (int c = 0; c < 16; ++c) for (int b = 0; b < 16; ++b) for (int a = 0; a < 16; ++a) { box.push_back(a); box.push_back(b); box.push_back(c); }
        auto pick = [&](const std::vector<size_t>& idx) { std::vector<int32_t> o; for (size_t i : idx) { o.push_back(box[3 * i]); o.push_back(box[3 * i + 1]); o.push_back(box[3 * i + 2]); } return o; };
        std::vector<size_t> fwd(4096), rev(4096), shuf(4096), gaps, runs;
        for (size_t i = 0; i < 4096; ++i) { fwd[i] = i; rev[i] = 4095 - i; shuf[i] = i; }
        std::shuffle(shuf.begin(), shuf.end(), rng);
        for (size_t i = 0; i < 4096; i += 3) gaps.push_back(i);
        for (size_t i = 2000; i < 2600; ++i) runs.push_back(i);
        for (size_t i = 0; i < 700; ++i) runs.push_back(i);
        for (size_t i = 4000; i < 4096; ++i) runs.push_back(i);
        char w[96];
        snprintf(w, 96, "kind %d: first load, box order", kind); bad += compare(A, B, pick(fwd), w);
        snprintf(w, 96, "kind %d: again", kind); bad += compare(A, B, pick(fwd), w);
        snprintf(w, 96, "kind %d: reversed", kind); bad += compare(A, B, pick(rev), w);
        snprintf(w, 96, "kind %d: shuffled", kind); bad += compare(A, B, pick(shuf), w);
        snprintf(w, 96, "kind %d: every third", kind); bad += compare(A, B, pick(gaps), w);
        snprintf(w, 96, "kind %d: runs with jumps", kind); bad += compare(A, B, pick(runs), w);
        std::vector<int32_t> dup = pick(fwd); dup[3 * 300] = dup[3 * 299]; dup[3 * 300 + 1] = dup[3 * 299 + 1]; dup[3 * 300 + 2] = dup[3 * 299 + 2];
        snprintf(w, 96, "kind %d: duplicate of the previous entry", kind); bad += compare(A, B, dup, w);
        dup = pick(fwd); dup[3 * 900] = dup[3 * 17]; dup[3 * 900 + 1] = dup[3 * 17 + 1]; dup[3 * 900 + 2] = dup[3 * 17 + 2];
        snprintf(w, 96, "kind %d: duplicate of an early entry", kind); bad += compare(A, B, dup, w);
        // holes: unload slots 40..44 and 1000 on both (host bookkeeping of dsp_unload_chunks)
        for (dsp_ctx* c : {A, B}) {
            uint32_t base; int32_t p[3] = {0, 0, 0}; find_slot(c, p, &base);
            for (uint32_t s : {base + 40u, base + 41u, base + 42u, base + 43u, base + 44u, base + 1000u}) {
                if (!slot_in_box_region(c, s)) { c->map.erase(c->slot_pos[s]); c->free_slots.push_back(s); }
                c->resident[s] = 0; c->n_resident--;
            }
        }
        snprintf(w, 96, "kind %d: with holes (reloaded)", kind); bad += compare(A, B, pick(fwd), w);
        snprintf(w, 96, "kind %d: with holes, again", kind); bad += compare(A, B, pick(fwd), w);
        snprintf(w, 96, "kind %d: with holes, shuffled", kind); bad += compare(A, B, pick(shuf), w);
        // a run that leaves the big box for the small one and map slots
        std::vector<int32_t> mixed = pick({4090, 4091, 4092, 4093, 4094, 4095});
        for (int c = 0; c < 2; ++c) for (int b = 0; b < 2; ++b) for (int a = 0; a < 3; ++a) { mixed.push_back(-3 + a); mixed.push_back(5 + b); mixed.push_back(40 + c); }
        for (int k = 0; k < 7; ++k) { mixed.push_back(100 + k); mixed.push_back(0); mixed.push_back(0); }
        for (int k = 0; k < 3; ++k) { mixed.push_back(k); mixed.push_back(0); mixed.push_back(0); }
        snprintf(w, 96, "kind %d: box end -> small box -> map -> box", kind); bad += compare(A, B, mixed, w);
        snprintf(w, 96, "kind %d: the same again", kind); bad += compare(A, B, mixed, w);
    }
    // speed
    for (int kind = 0; kind < 2; ++kind) {
        dsp_ctx* ctx = fresh_ctx(kind == 1, false);
        std::vector<int32_t> pos;
        for (int c = 0; c < 16; ++c) for (int b = 0; b < 16; ++b) for (int a = 0; a < 16; ++a) { pos.push_back(a); pos.push_back(b); pos.push_back(c); }
        std::vector<uint32_t> slots(4096), fresh;
        std::vector<int4> hp(4096);
        for (int rep = 0; rep < 5; ++rep) {
            new_stamp(ctx);
            auto t0 = std::chrono::steady_clock::now();
            resolve_positions(ctx, pos.data(), 0, 4096, slots.data(), hp.data(), fresh);
            auto t1 = std::chrono::steady_clock::now();
            printf("%s rep %d: %.1f us for 4096 chunks\n", kind ? "box slots" : "map slots", rep, std::chrono::duration<double, std::micro>(t1 - t0).count());
        }
    }
    printf(bad ? "FAILED\n" : "ALL OK\n");
    return bad;
}
