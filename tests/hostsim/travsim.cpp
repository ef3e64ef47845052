// travsim.cpp — DESIGN EXPLORATION ONLY (not product, not a test): a lock-step SIMT replay of the
// secondary megakernel (render_kernel<SPLIT>) over real C2 path segments, under different BVH
// traversal algorithms, to decide which ones are worth GPU minutes ("next round" item 1):
//
//   bvh2        today's while-while traversal of the binary tree (calibration against ncu)
//   bvh2-ls     "leaf start": a secondary ray starts on a known primitive, so its leaf and the chain of
//               ancestors are known; their boxes contain the origin and need no test — the walk tests
//               only the siblings along the path, bottom-up, two per step
//   bvh4, bvh8  wide trees collapsed from the SAH binary tree (largest-area child expanded first),
//               one node = K box tests, hits sorted front to back
//   bvh4-ls, bvh8-ls  wide tree + leaf start
//   coop8       8 lanes per ray on the 8-wide tree (4 rays per warp, one child box per lane)
//
// Cost model (issue slots per warp-level step, from the SASS of the round-1 kernel: 50 per pair
// step, 35 per sphere test + 10 per leaf; wide steps scaled by their arithmetic): see COST_* below.
// Built by `make -C tests/hostsim libtravsim.so`, driven by scripts/sim_traversal_variants.py.
#include "hostsim.cpp"

#include <cstdio>
#include <map>
#include <numeric>

namespace {

struct SRay {
    V3 o, d;
    uint32_t start_slot; // sphere slot the ray leaves from (RT_REF_NONE for camera rays)
    int blocks;          // Philox blocks consumed by the shading that follows this segment
    bool hit;
};
struct SPath {
    std::vector<SRay> rays; // secondary segments only
};

struct Tree {
    const HostScene *s;
    std::vector<int> parent_node;  // per pair index/2: BVH2 node whose ref is that pair (-1: root pair)
    std::vector<int> slot_leaf;    // sphere slot -> BVH2 node index of its leaf
    std::vector<int> node_depth;   // per node
    // wide tree for K: per pair index, the child entries (BVH2 node indices)
    std::map<int, std::vector<int>> wide[9];
    std::vector<int> wide_home[9]; // per BVH2 node: pair index of the wide node that lists it as a child (-1: none)

    uint32_t ref_of(int node) const { return as_uint(s->bvh.dev_nodes[2 * (size_t)node].w); }
    float area(int node) const {
        const F4 &h = s->bvh.dev_nodes[2 * (size_t)node + 1];
        return h.x * h.y + h.y * h.z + h.z * h.x;
    }
    void build() {
        const size_t n_nodes = s->bvh.nodes.size() / 2; // the tree's own nodes (dev_nodes also holds the product's walk pairs)
        parent_node.assign(n_nodes / 2 + 1, -1);
        slot_leaf.assign(s->bvh.sph.size(), -1);
        node_depth.assign(n_nodes, 0);
        for (size_t i = 0; i < n_nodes; i++) {
            const uint32_t ref = ref_of((int)i);
            if (ref & RT_LEAF) {
                const uint32_t first = (ref & RT_LEAF_SLOT_MASK) >> 3, count = (ref & 7u) + 1;
                if (!(ref & RT_LEAF_QUAD))
                    for (uint32_t q = first; q < first + count; q++) slot_leaf[q] = (int)i;
            } else {
                parent_node[ref / 2] = (int)i;
            }
        }
        for (size_t i = 2; i < n_nodes; i++) node_depth[i] = node_depth[parent_node[i / 2]] + 1; // pre-order: parents first
        for (int K : {4, 8}) {
            wide_home[K].assign(n_nodes, -1);
            std::vector<int> todo{(int)s->bvh.root_ref};
            while (!todo.empty()) {
                const int pair = todo.back();
                todo.pop_back();
                std::vector<int> ch{pair, pair + 1};
                for (;;) {
                    if ((int)ch.size() >= K) break;
                    int bi = -1;
                    float ba = -1;
                    for (size_t c = 0; c < ch.size(); c++)
                        if (!(ref_of(ch[c]) & RT_LEAF) && area(ch[c]) > ba) ba = area(ch[c]), bi = (int)c;
                    if (bi < 0) break;
                    const int q = (int)ref_of(ch[bi]);
                    ch[bi] = q;
                    ch.push_back(q + 1);
                }
                for (int c : ch) {
                    wide_home[K][c] = pair;
                    if (!(ref_of(c) & RT_LEAF)) todo.push_back((int)ref_of(c));
                }
                wide[K][pair] = ch;
            }
        }
    }
};

const int COST_PAIR = 50, COST_SPH = 35, COST_LEAF = 10, COST_REGEN = 30, COST_SHADE = 150, COST_PHILOX = 55;
int cost_wide(int K) { return K == 4 ? 100 : 215; } // K box tests (14 each) + loads + sorting network + pushes

struct Counters {
    double slots = 0;        // warp issue slots
    double inner_slots = 0;  // ... spent in inner steps
    double inner_lane = 0;   // sum over inner steps of active lanes
    double inner_steps = 0;  // warp-level inner steps
    double box = 0, sph = 0; // thread-level tests
    double thread_steps = 0;
    double rays = 0;
    double cls_steps[2] = {0, 0}, cls_rays[2] = {0, 0}, cls_sq[2] = {0, 0}; // [0] starts on a small sphere, [1] on the ground
};

// per-lane traversal state shared by the variants
struct Lane {
    bool alive = false;
    size_t path = 0, seg = 0;
    V3 o, d, inv, noi, ainv;
    float a, tbest;
    uint32_t best, ref;
    std::vector<uint32_t> stack;
    int up; // leaf-start walk: BVH2 node whose siblings are tested next (-1: done)
};

struct Variant {
    const Tree *t;
    int K;       // 2, 4, 8
    bool ls;     // leaf start
    bool walk1 = false; // binary leaf start: one sibling per walk step instead of two
    bool nested = false; // walk steps run in their own loop after the leaf phase (no bookkeeping in the descend loop)
    bool near = false;   // rays that leave a huge primitive start at the leaf of the small sphere nearest to their origin
    int extra = 0;       // extra issue slots per merged inner step (walk/descend selection)
    const F4 *nodes() const { return t->s->bvh.dev_nodes.data(); }
    bool box(const Lane &l, int node, float &tn) const {
        return box_test(nodes()[2 * (size_t)node], nodes()[2 * (size_t)node + 1], l.inv, l.noi, l.ainv, 0.001f, l.tbest, tn);
    }
    void init(Lane &l, const SRay &r) const {
        l.o = r.o, l.d = r.d;
        l.inv = v3(cull_rcp(r.d.x), cull_rcp(r.d.y), cull_rcp(r.d.z));
        l.noi = v3(-(r.o.x * l.inv.x), -(r.o.y * l.inv.y), -(r.o.z * l.inv.z));
        l.ainv = v3(fabsf(l.inv.x), fabsf(l.inv.y), fabsf(l.inv.z));
        l.a = lensq(r.d), l.tbest = INFINITY, l.best = RT_REF_NONE;
        l.stack.clear();
        l.up = -1;
        if (ls && r.start_slot != RT_REF_NONE) {
            uint32_t slot = r.start_slot;
            if (near && fabsf(t->s->bvh.sph[slot].w) > 100) {
                float bd = INFINITY;
                for (size_t q = 0; q < t->s->bvh.sph.size(); q++) {
                    const F4 &sp = t->s->bvh.sph[q];
                    if (fabsf(sp.w) > 100) continue;
                    const float dx = sp.x - r.o.x, dy = sp.y - r.o.y, dz = sp.z - r.o.z, dd = dx * dx + dy * dy + dz * dz;
                    if (dd < bd) bd = dd, slot = (uint32_t)q;
                }
            }
            l.up = t->slot_leaf[slot];
            l.ref = t->ref_of(l.up);
        } else {
            l.ref = t->s->bvh.root_ref;
        }
    }
    bool done(const Lane &l) const { return l.ref == RT_REF_NONE && l.up < 0; }
    bool want_inner(const Lane &l) const { return !(l.ref & RT_LEAF) || (!nested && l.ref == RT_REF_NONE && l.up >= 0); }
    bool want_walk(const Lane &l) const { return nested && l.ref == RT_REF_NONE && l.up >= 0; }
    bool want_leaf(const Lane &l) const { return l.ref != RT_REF_NONE && (l.ref & RT_LEAF); }
    uint32_t pop(Lane &l) const {
        if (l.stack.empty()) return RT_REF_NONE;
        const uint32_t r = l.stack.back();
        l.stack.pop_back();
        return r;
    }
    // test the candidate nodes, go to the nearest hit, push the rest farthest first
    void dispatch(Lane &l, const int *cand, int n, Counters &c) const {
        std::pair<float, uint32_t> hits[8];
        int nh = 0;
        for (int i = 0; i < n; i++) {
            float tn;
            c.box += 1;
            if (box(l, cand[i], tn)) hits[nh++] = {tn, t->ref_of(cand[i])};
        }
        std::stable_sort(hits, hits + nh, [](const auto &x, const auto &y) { return x.first < y.first; });
        for (int i = nh - 1; i >= 1; i--) l.stack.push_back(hits[i].second);
        l.ref = nh ? hits[0].second : pop(l);
    }
    void inner(Lane &l, Counters &c) const {
        c.thread_steps += 1;
        int cand[8], n = 0;
        if (!(l.ref & RT_LEAF)) {
            if (K == 2) cand[0] = (int)l.ref, cand[1] = (int)l.ref + 1, n = 2;
            else
                for (int ch : t->wide[K].at((int)l.ref)) cand[n++] = ch;
        } else if (K == 2) { // walk step: siblings of `up` and of its parent
            const int p = t->parent_node[l.up / 2];
            cand[n++] = l.up ^ 1;
            if (walk1) {
                l.up = p;
            } else {
                if (p >= 0) cand[n++] = p ^ 1;
                l.up = p >= 0 ? t->parent_node[p / 2] : -1;
            }
        } else { // wide walk step: the other children of the wide node that lists `up`
            const int home = t->wide_home[K][l.up];
            for (int ch : t->wide[K].at(home))
                if (ch != l.up) cand[n++] = ch;
            l.up = t->parent_node[home / 2];
        }
        dispatch(l, cand, n, c);
    }
    int leaf(Lane &l, Counters &c) const {
        const uint32_t first = (l.ref & RT_LEAF_SLOT_MASK) >> 3, count = (l.ref & 7u) + 1;
        for (uint32_t q = first; q < first + count; q++) {
            float tt;
            c.sph += 1;
            if (!sphere_candidate(t->s->bvh.sph[q], l.o, l.d, l.a, 0.001f, tt)) continue;
            if (tt < l.tbest) l.tbest = tt, l.best = q;
            else if (tt == l.tbest && l.best != RT_REF_NONE && t->s->bvh.meta[q].x < t->s->bvh.meta[l.best].x) l.best = q;
        }
        l.ref = pop(l);
        return (int)count;
    }
};

// megakernel replay: `width` lanes per warp (32, or 4 for the cooperative variant)
Counters replay(const Variant &v, const std::vector<SPath> &paths, int width, int step_cost, bool check,
                const std::vector<std::vector<uint32_t>> *expect) {
    Counters c;
    std::vector<Lane> lanes(width);
    size_t next = 0;
    for (;;) {
        for (auto &l : lanes)
            if (!l.alive && next < paths.size()) l.alive = true, l.path = next++, l.seg = 0;
        int alive = 0;
        for (auto &l : lanes) alive += l.alive;
        if (!alive) break;
        c.slots += COST_REGEN * width / 32.0; // cooperative variant: regeneration and shading stay 32 wide
        double steps0[64];
        int li = 0;
        for (auto &l : lanes) {
            steps0[li++] = c.thread_steps; // (filled below per lane)
            if (l.alive) v.init(l, paths[l.path].rays[l.seg]), c.rays += 1;
        }
        std::vector<double> lane_steps(width, 0.0);
        for (;;) { // while-while
            for (;;) {
                int n = 0;
                for (auto &l : lanes)
                    if (l.alive && v.want_inner(l)) n++;
                if (!n) break;
                for (int q = 0; q < width; q++)
                    if (lanes[q].alive && v.want_inner(lanes[q])) v.inner(lanes[q], c), lane_steps[q] += 1;
                c.slots += step_cost + v.extra, c.inner_slots += step_cost + v.extra, c.inner_lane += n, c.inner_steps += 1;
            }
            int mx = 0;
            for (auto &l : lanes)
                if (l.alive && v.want_leaf(l)) mx = std::max(mx, v.leaf(l, c));
            if (mx) c.slots += (width == 32 ? mx * COST_SPH : COST_SPH + 10) + COST_LEAF; // cooperative: one sphere per lane
            bool walked = false;
            for (;;) { // nested walk loop: lanes whose stack ran dry test the next siblings on their path
                int n = 0;
                for (int q = 0; q < width; q++)
                    if (lanes[q].alive && v.want_walk(lanes[q])) v.inner(lanes[q], c), lane_steps[q] += 1, n++;
                if (!n) break;
                walked = true;
                c.slots += 45, c.inner_slots += 45, c.inner_lane += n, c.inner_steps += 1;
            }
            if (!mx && !walked) break;
        }
        int blocks = 0;
        for (auto &l : lanes) {
            if (!l.alive) continue;
            const SRay &r = paths[l.path].rays[l.seg];
            if (check && l.best != (*expect)[l.path][l.seg]) {
                fprintf(stderr, "MISMATCH path %zu seg %zu: %u vs %u\n", l.path, l.seg, l.best, (*expect)[l.path][l.seg]);
                abort();
            }
            blocks = std::max(blocks, r.blocks);
            {
                const int cls = fabsf(v.t->s->bvh.sph[r.start_slot].w) > 100 ? 1 : 0;
                const double st = lane_steps[&l - &lanes[0]];
                c.cls_steps[cls] += st, c.cls_rays[cls] += 1, c.cls_sq[cls] += st * st;
            }
            if (++l.seg >= paths[l.path].rays.size()) l.alive = false;
        }
        c.slots += (COST_SHADE + COST_PHILOX * blocks) * width / 32.0;
    }
    return c;
}

} // namespace

extern "C" int ts_run(const rt_scene_desc *d, const rt_camera *cam, uint64_t seed, int32_t spp, int32_t row_stride,
                      int max_leaf) {
    HostScene s;
    load(d, max_leaf, 0, &s);
    DevCamera c = make_dev_camera(*cam);
    std::vector<SPath> paths;
    std::vector<std::vector<uint32_t>> expect;
    double n_primary = 0, prim_box = 0;
    for (int row = row_stride / 2; row < cam->height; row += row_stride)
        for (int col = 0; col < cam->width; col++) {
            const int64_t pix = (int64_t)row * cam->width + col;
            for (int k = 0; k < spp; k++) {
                PathRng rng;
                rng.init(seed, (uint32_t)pix, (uint32_t)k);
                V3 o, dir;
                generate_ray(c, rng, col, row, o, dir);
                SPath p;
                std::vector<uint32_t> ex;
                uint32_t start = RT_REF_NONE;
                for (int depth = 0; depth < c.max_depth;) {
                    LocalStack<64> stack;
                    HitRec h;
                    WorkCounters wc{0, 0};
                    trace_closest<LocalStack<64>, true, false>(s.bvh.dev_nodes.data(), s.bvh.sph.data(), s.bvh.meta.data(),
                                                               s.bvh.root_ref, o, dir, 0.001f, INFINITY, stack, h, &wc);
                    if (depth == 0) n_primary += 1, prim_box += (double)wc.box_tests;
                    SRay r;
                    r.o = o, r.d = dir, r.start_slot = start, r.blocks = 0, r.hit = h.slot != RT_REF_NONE;
                    bool go_on = false;
                    if (r.hit) {
                        const uint32_t b0 = rng.block;
                        V3 atten, emitted;
                        go_on = shade_at(s, h, rng, o, dir, atten, emitted);
                        r.blocks = (int)(rng.block - b0);
                        start = h.slot;
                    }
                    if (depth > 0) p.rays.push_back(r), ex.push_back(h.slot);
                    if (!go_on) break;
                    depth++;
                }
                if (!p.rays.empty()) paths.push_back(std::move(p)), expect.push_back(std::move(ex));
            }
        }
    Tree t;
    t.s = &s;
    t.build();
    double n_sec = 0, ground = 0, depth_sum = 0;
    for (auto &p : paths)
        for (auto &r : p.rays) {
            n_sec += 1;
            const int leaf = t.slot_leaf[r.start_slot];
            depth_sum += t.node_depth[leaf];
            if (fabsf(s.bvh.sph[r.start_slot].w) > 100) ground += 1;
        }
    printf("primary paths %.0f (%.2f box tests/ray), surviving paths %zu, secondary rays %.0f (%.2f per survivor)\n", n_primary,
           prim_box / n_primary, paths.size(), n_sec, n_sec / paths.size());
    printf("secondary rays that start on the ground sphere: %.1f %%; mean depth of the start leaf: %.2f; tree height %u, %zu nodes\n",
           100 * ground / n_sec, depth_sum / n_sec, s.bvh.max_depth, s.bvh.dev_nodes.size() / 2);
    for (int K : {4, 8}) {
        double ch = 0;
        for (auto &w : t.wide[K]) ch += w.second.size();
        printf("bvh%d: %zu wide nodes, %.2f children per node\n", K, t.wide[K].size(), ch / t.wide[K].size());
    }
    printf("%-10s %9s %9s %9s %9s %9s %9s %9s\n", "variant", "slots/ray", "vs bvh2", "steps/ray", "box/ray", "sph/ray", "lanes", "inner%");
    double base = 0;
    struct Cfg {
        const char *name;
        int K;
        bool ls;
        int width;
        bool walk1;
        bool nested = false;
        int extra = 0;
        bool near = false;
    } cfgs[] = {{"bvh2", 2, false, 32, false},  {"bvh2-ls", 2, true, 32, false}, {"bvh2-ls1", 2, true, 32, true},
                {"bvh2-ls+12", 2, true, 32, false, false, 12}, {"bvh2-ls-nest", 2, true, 32, false, true, 0},
                {"bvh2-ls-near", 2, true, 32, false, false, 0, true}, {"bvh2-ls-near+3", 2, true, 32, false, false, 3, true},
                {"bvh4", 4, false, 32, false},  {"bvh4-ls", 4, true, 32, false}, {"bvh8", 8, false, 32, false},
                {"bvh8-ls", 8, true, 32, false}, {"coop8", 8, false, 4, false},  {"coop8-ls", 8, true, 4, false}};
    for (auto &cf : cfgs) {
        Variant v{&t, cf.K, cf.ls, cf.walk1, cf.nested, cf.near, cf.extra};
        const int step = cf.width == 4 ? 45 : (cf.K == 2 ? COST_PAIR : cost_wide(cf.K));
        // (a start leaf far from the origin may add a candidate the padded boxes cull: outside the envelope)
        Counters r = replay(v, paths, cf.width, step, !cf.near, &expect);
        const double slots_per_ray = r.slots / r.rays; // warp issue slots per ray
        if (base == 0) base = slots_per_ray;
        printf("%-10s %9.1f %9.3f %9.2f %9.2f %9.2f %9.2f %8.1f%%\n", cf.name, slots_per_ray, slots_per_ray / base,
               r.thread_steps / r.rays, r.box / r.rays, r.sph / r.rays, r.inner_lane / r.inner_steps, 100 * r.inner_slots / r.slots);
        for (int q = 0; q < 2; q++) {
            const double m = r.cls_steps[q] / r.cls_rays[q];
            printf("    %s: %.0f rays, %.2f +- %.2f steps/ray\n", q ? "ground start" : "sphere start", r.cls_rays[q], m,
                   sqrt(r.cls_sq[q] / r.cls_rays[q] - m * m));
        }
    }
    return 0;
}
