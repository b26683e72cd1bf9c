/* Scheduling study, second version (developer tool, CPU only): a timed model of the vote-scheduled kernel.
 *
 * Records for every ray the steps the reference order takes -- inner-node visits with their hit mask and the
 * children that became the new top, Tri4 packets, cull pops -- and replays them through a model of the kernel in
 * rodent_b200/csrc/traverse_sched.cuh: 148 SMs x 20 resident warps, one ray per lane, a global ray counter, the
 * majority vote, node-step streaks, and a per-block instruction cost taken from the ncu source view
 * (profiles/r01_traverse_vote_streaks_blocks.txt).  Time: every warp iteration issues `cost` warp instructions at
 * min(R_WARP, R_SM / active warps on its SM) per cycle.  The model answers "what if" questions (refill threshold,
 * a separate sort phase, handing the last rays of a launch to a second kernel) before GPU time is spent on them.
 *
 * build: gcc -O2 -march=x86-64-v3 -ffp-contract=off -o /tmp/sim2 scripts/sim_sched2.c -lpthread -lm
 * usage: sim2 BVH8_FILE RAYS_FILE tmin tmax [max_rays [mortonB]]   (mortonB: rays reordered by a 3B-bit Morton key of the origin)
 */
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
typedef struct { uint8_t kind, mask, tops, culls; } Step;
static Step* g_steps; static size_t g_len, g_cap;
static void put_step(int kind) {
    if (g_len == g_cap) { g_cap = g_cap ? g_cap * 2 : (1u << 24); g_steps = realloc(g_steps, g_cap * sizeof(Step)); }
    g_steps[g_len].kind = (uint8_t)kind; g_steps[g_len].mask = 0; g_steps[g_len].tops = 0; g_steps[g_len].culls = 0; g_len++;
}
static size_t g_ray_first;
static void trace(int c) {
    if (c == 'N' || c == 'L') put_step(c);
    else if (c == 'c' && g_len > g_ray_first) g_steps[g_len - 1].culls++;
}
#include <stdlib.h>
#define ORACLE_TRACE(kind) trace(kind)
#define ORACLE_TRACE_PUSHES(m, t) do { g_steps[g_len - 1].mask = (uint8_t)(m); g_steps[g_len - 1].tops = (uint8_t)(t); } while (0)
#include "../oracle/traversal_oracle.c"

static void* read_file(const char* path, size_t* size) {
    FILE* f = fopen(path, "rb"); if (!f) { perror(path); exit(1); }
    fseek(f, 0, SEEK_END); *size = (size_t)ftell(f); fseek(f, 0, SEEK_SET);
    void* p = malloc(*size); if (fread(p, 1, *size, f) != *size) exit(1); fclose(f); return p;
}

/* ---- model parameters ---- */
typedef struct {
    int refill_min, streak_min;
    int sort_min;          /* sort phase runs when at least this many lanes wait for it (0: plain majority) */
    int sort_phase;        /* 1: the sort of a node step is a step kind of its own in the vote */
    int chain_push;        /* 1: pushes as a predicated compare-exchange chain over the slots any lane hit */
    int orphan_k;          /* > 0: once the ray queue is drained, a warp with <= k live rays hands them to a second launch */
    int orphan_quad;       /* second launch runs with this many lanes per ray (1 or 4) */
    int age_limit;         /* > 0: a ray older than this many steps is handed to the orphan queue at once, and refills
                              take orphans first (old rays travel together) */
    double r_sm, r_warp;   /* issue rates, warp instructions per cycle */
    int warps_per_sm;
    double k_push, k_sort, k_refill, k_loop;   /* what-if scales of the push, sort, refill and loop-overhead costs (1: as measured) */
    int coop_sort;         /* > 0: when at most this many lanes of a node step need a sort, the warp sorts them together (what if) */
    int postpone;          /* > 0: a lane that reaches a leaf keeps it pending and goes on with node steps until it reaches the
                              next leaf (Aila-Laine's speculative traversal); a leaf step runs once this many lanes hold one.
                              Optimistic: the same steps in another order, the extra node visits under the stale tmax not counted */
} Policy;

enum { SMS = 148 };
enum { C_LOOP = 60, C_STREAK = 12, C_REFILL = 45, C_INIT = 130, C_NODE = 212, C_PUSH = 15, C_CHAIN_SLOT = 9, C_CHAIN_BASE = 16,
       C_SORT4 = 63, C_SORT8 = 153, C_COOP_SORT = 95, C_CULL = 9, C_LEAF = 195, C_DUMP = 90, C_RESTORE = 110, C_VOTE_SORT = 8 };

typedef struct { int ray; size_t pos, end; int need_sort; int age; size_t ppos, pend; } Lane;   /* [ppos, pend): the pending leaf run */
typedef struct { Lane l[32]; int sm; int alive; int drained; double t; } Warp;

typedef struct { double cat[8]; double time_us, drained_us, k1_us; double warp_inst, thread_inst; double node_exec, node_lanes, leaf_exec, leaf_lanes, sort_exec, sort_lanes; long orphans; } Result;

/* binary heap of warps by time */
static int* g_heap; static int g_hn; static Warp* g_w;
static void heap_push(int i) { int k = g_hn++; g_heap[k] = i; while (k > 0) { int p = (k - 1) / 2; if (g_w[g_heap[p]].t <= g_w[g_heap[k]].t) break; int x = g_heap[p]; g_heap[p] = g_heap[k]; g_heap[k] = x; k = p; } }
static int heap_pop(void) {
    int top = g_heap[0]; g_heap[0] = g_heap[--g_hn]; int k = 0;
    for (;;) { int a = 2 * k + 1, b = a + 1, m = k; if (a < g_hn && g_w[g_heap[a]].t < g_w[g_heap[m]].t) m = a; if (b < g_hn && g_w[g_heap[b]].t < g_w[g_heap[m]].t) m = b; if (m == k) break; int x = g_heap[m]; g_heap[m] = g_heap[k]; g_heap[k] = x; k = m; }
    return top;
}

typedef struct { int ray; size_t pos, end; int age; } Orphan;

static double node_cost(const Policy* P, Warp* w, unsigned go, Result* R, int* any_sort) {
    /* executes one node step for the lanes in `go`; returns its cost */
    double c = C_NODE;
    unsigned un = 0; unsigned blocks = 0; int s4 = 0, s8 = 0, maxcull = 0, lanes = 0;
    for (int i = 0; i < 32; i++) if (go & (1u << i)) {
        Lane* q = &w->l[i]; const Step* s = &g_steps[q->pos];
        un |= s->mask;
        blocks |= (unsigned)s->tops | ((unsigned)(s->mask & ~s->tops) << 8);
        const int n = __builtin_popcount(s->mask);
        if (n >= 3) { if (P->sort_phase) q->need_sort = n; else if (n <= 4) s4++; else s8++; }
        if (s->culls > maxcull) maxcull = s->culls;
        q->pos++; q->age++; lanes++;
    }
    if (P->chain_push) c += C_CHAIN_BASE + C_CHAIN_SLOT * __builtin_popcount(un);
    else c += P->k_push * C_PUSH * __builtin_popcount(blocks);
    if (P->coop_sort > 0 && s4 + s8 > 0 && s4 + s8 <= P->coop_sort) {
        /* the warp sorts up to eight lanes' entries together, four lanes per job, one comparator per lane and layer */
        c += C_COOP_SORT * ((s4 + s8 + 7) / 8); R->sort_exec++; R->sort_lanes += 4 * (s4 + s8); R->thread_inst += C_COOP_SORT * 4 * (s4 + s8);
    } else {
    if (s4) { c += P->k_sort * C_SORT4; R->sort_exec++; R->sort_lanes += s4; R->thread_inst += C_SORT4 * s4; }
    if (s8) { c += P->k_sort * C_SORT8; R->sort_exec++; R->sort_lanes += s8; R->thread_inst += C_SORT8 * s8; }
    }
    c += C_CULL * maxcull;
    R->node_exec++; R->node_lanes += lanes; R->thread_inst += (double)C_NODE * lanes;
    *any_sort = s4 + s8;
    return c;
}

static Result simulate(const size_t* start, int num_rays, const Policy* P) {
    Result R = {0};
    const int nw = SMS * P->warps_per_sm;
    Warp* W = calloc((size_t)nw, sizeof(Warp)); g_w = W;
    g_heap = malloc(sizeof(int) * (size_t)nw); g_hn = 0;
    int active[SMS] = {0};
    for (int i = 0; i < nw; i++) { W[i].sm = i % SMS; W[i].alive = 1; for (int l = 0; l < 32; l++) W[i].l[l].ray = -1; active[W[i].sm]++; }
    for (int i = 0; i < nw; i++) heap_push(i);
    int next_ray = 0;
    Orphan* orph = malloc(sizeof(Orphan) * (size_t)num_rays); long n_orph = 0, orph_taken = 0;
    int phase2 = 0;
    double t_end = 0;
    for (;;) {
        if (g_hn == 0) {
            if (!phase2 && n_orph > orph_taken) {
                /* second launch over the orphan queue */
                phase2 = 1; R.k1_us = t_end / 1965.0; R.orphans = n_orph;
                const double t0 = t_end + 4.0 * 1965.0;          /* 4 us launch gap */
                const int lanes_per_ray = P->orphan_quad > 1 ? P->orphan_quad : 1;
                const int rays_per_warp = 32 / lanes_per_ray;
                long need = (n_orph - orph_taken + rays_per_warp - 1) / rays_per_warp;
                if (need > nw) need = nw;
                for (int s = 0; s < SMS; s++) active[s] = 0;
                for (int i = 0; i < need; i++) { W[i].alive = 1; W[i].drained = 0; W[i].t = t0; W[i].sm = i % SMS; active[W[i].sm]++; for (int l = 0; l < 32; l++) W[i].l[l].ray = -1; heap_push(i); }
                continue;
            }
            break;
        }
        const int wi = heap_pop();
        Warp* w = &W[wi];
        double cost = 0;
        const int quad = phase2 && P->orphan_quad > 1;
        const int slots = quad ? 32 / P->orphan_quad : 32;
        /* finished rays leave */
        int idle = 0, live = 0;
        for (int i = 0; i < slots; i++) { Lane* q = &w->l[i]; if (q->ray >= 0 && q->pos == q->end && !q->need_sort && q->ppos == q->pend) q->ray = -1; if (q->ray < 0) idle++; else live++; }
        /* hand old rays over */
        if (!phase2 && P->age_limit > 0) {
            int moved = 0;
            for (int i = 0; i < 32; i++) { Lane* q = &w->l[i]; if (q->ray >= 0 && q->age > P->age_limit && q->age < 1000000) { orph[n_orph++] = (Orphan){q->ray, q->pos, q->end, 1000000}; q->ray = -1; idle++; live--; moved++; } }
            if (moved) cost += C_DUMP;
        }
        /* refill */
        if (!w->drained && (idle >= P->refill_min || idle == slots)) {
            cost += P->k_refill * C_REFILL;
            int took = 0, took_orph = 0;
            for (int i = 0; i < slots; i++) { Lane* q = &w->l[i]; if (q->ray >= 0) continue;
                if (phase2 || (P->age_limit > 0 && orph_taken < n_orph)) {
                    if (orph_taken < n_orph) { Orphan o = orph[orph_taken++]; q->ray = o.ray; q->pos = o.pos; q->end = o.end; q->age = o.age; q->need_sort = 0; took_orph++; live++; continue; }
                    if (phase2) continue;
                }
                if (next_ray < num_rays) { q->ray = next_ray; q->pos = start[next_ray]; q->end = start[next_ray + 1]; q->age = 0; q->need_sort = 0; q->ppos = q->pend = 0; next_ray++; took++; live++; }
            }
            if (took) cost += P->k_refill * C_INIT;
            if (took_orph) cost += C_RESTORE;
            if (phase2 ? orph_taken >= n_orph : next_ray >= num_rays) w->drained = 1;
            if (!phase2 && next_ray >= num_rays && R.drained_us == 0) R.drained_us = w->t / 1965.0;
        }
        /* the last rays of a launch go to the orphan queue */
        if (!phase2 && P->orphan_k > 0 && w->drained && live > 0 && live <= P->orphan_k && (P->age_limit == 0 || orph_taken >= n_orph)) {
            for (int i = 0; i < 32; i++) { Lane* q = &w->l[i]; if (q->ray >= 0) { if (q->pos < q->end || q->need_sort) orph[n_orph++] = (Orphan){q->ray, q->pos, q->end, q->age}; q->ray = -1; } }
            cost += C_DUMP; live = 0;
        }
        /* vote */
        unsigned bn = 0, bl = 0, bs = 0;
        if (P->postpone > 0) {
            unsigned must = 0;                                     /* lanes that can only go on with a leaf step */
            for (int i = 0; i < slots; i++) { Lane* q = &w->l[i]; if (q->ray < 0) continue;
                if (q->ppos == q->pend && q->pos < q->end && g_steps[q->pos].kind == 'L') {      /* reached a leaf: keep it pending, look past it */
                    q->ppos = q->pos; while (q->pos < q->end && g_steps[q->pos].kind == 'L') q->pos++; q->pend = q->pos;
                }
                const int can_node = q->pos < q->end && g_steps[q->pos].kind == 'N';
                if (q->ppos < q->pend) { bl |= 1u << i; if (!can_node) must |= 1u << i; }
                if (can_node) bn |= 1u << i;
            }
            /* leaf step when enough lanes hold one, or more lanes are stuck behind theirs than can take a node step */
            const int cl = __builtin_popcount(bl), cn = __builtin_popcount(bn), cm = __builtin_popcount(must);
            if (cl > 0 && (cl >= P->postpone || cn == 0 || cm > cn)) {
                int maxcull = 0;
                for (int i = 0; i < 32; i++) if (bl & (1u << i)) { Lane* q = &w->l[i]; if (g_steps[q->ppos].culls > maxcull) maxcull = g_steps[q->ppos].culls; q->ppos++; q->age++; }
                cost += P->k_loop * C_LOOP + C_LEAF + C_CULL * maxcull;
                R.leaf_exec++; R.leaf_lanes += cl; R.thread_inst += (double)C_LEAF * cl;
                bn = bl = 0;
            } else if (cn > 0) {
                unsigned go = bn; int first = 1;
                cost += P->k_loop * C_LOOP;
                do {
                    int any_sort;
                    cost += node_cost(P, w, go, &R, &any_sort) + (first ? 0 : C_STREAK);
                    first = 0; go = 0;
                    for (int i = 0; i < slots; i++) { Lane* q = &w->l[i]; if (q->ray >= 0 && q->pos < q->end && g_steps[q->pos].kind == 'N') go |= 1u << i; }
                } while (__builtin_popcount(go) >= P->streak_min);
                bn = bl = 0;
            } else if (live == 0 && w->drained) {
                w->alive = 0; active[w->sm]--; t_end = w->t > t_end ? w->t : t_end;
                R.warp_inst += cost;
                continue;
            } else {
                cost += P->k_loop * C_LOOP;
            }
            R.warp_inst += cost;
            double rate2 = P->r_sm / active[w->sm]; if (rate2 > P->r_warp) rate2 = P->r_warp;
            w->t += cost / rate2;
            heap_push(wi);
            continue;
        }
        for (int i = 0; i < slots; i++) { Lane* q = &w->l[i]; if (q->ray < 0) continue;
            if (q->need_sort) bs |= 1u << i; else if (q->pos < q->end) { if (g_steps[q->pos].kind == 'N') bn |= 1u << i; else bl |= 1u << i; } }
        cost += P->k_loop * C_LOOP + (P->sort_phase ? C_VOTE_SORT : 0);
        const double lane_scale = quad ? 0.45 : 1.0;     /* a quad step: a quarter of the slab tests per lane, shuffles for pushes and sort */
        if ((bn | bl | bs) == 0) {
            if (live == 0 && (w->drained)) {
                /* in the aged-orphan mode a drained warp with nothing to do must wait for orphans still in flight elsewhere: model as exit */
                w->alive = 0; active[w->sm]--; t_end = w->t > t_end ? w->t : t_end;
                R.warp_inst += cost;
                continue;
            }
        } else {
            const int cn = __builtin_popcount(bn), cl = __builtin_popcount(bl), cs = __builtin_popcount(bs);
            if (cs > 0 && (P->sort_min > 0 ? (cs >= P->sort_min || cn + cl == 0) : (cs >= cn && cs >= cl))) {
                int s4 = 0, s8 = 0;
                for (int i = 0; i < 32; i++) if (bs & (1u << i)) { if (w->l[i].need_sort <= 4) s4++; else s8++; w->l[i].need_sort = 0; }
                if (s4) { cost += C_SORT4; R.thread_inst += C_SORT4 * s4; }
                if (s8) { cost += C_SORT8; R.thread_inst += C_SORT8 * s8; }
                R.sort_exec++; R.sort_lanes += cs;
            } else if (cn >= cl) {
                unsigned go = bn; int first = 1;
                do {
                    int any_sort;
                    cost += (node_cost(P, w, go, &R, &any_sort) + (first ? 0 : C_STREAK)) * lane_scale;
                    first = 0;
                    go = 0;
                    for (int i = 0; i < slots; i++) { Lane* q = &w->l[i]; if (q->ray >= 0 && !q->need_sort && q->pos < q->end && g_steps[q->pos].kind == 'N') go |= 1u << i; }
                } while (__builtin_popcount(go) >= P->streak_min);
            } else {
                int maxcull = 0;
                for (int i = 0; i < 32; i++) if (bl & (1u << i)) { Lane* q = &w->l[i]; if (g_steps[q->pos].culls > maxcull) maxcull = g_steps[q->pos].culls; q->pos++; q->age++; }
                cost += (C_LEAF + C_CULL * maxcull) * lane_scale;
                R.leaf_exec++; R.leaf_lanes += cl; R.thread_inst += (double)C_LEAF * cl;
            }
        }
        R.warp_inst += cost;
        double rate = P->r_sm / active[w->sm]; if (rate > P->r_warp) rate = P->r_warp;
        w->t += cost / rate;
        heap_push(wi);
    }
    R.time_us = t_end / 1965.0;
    free(W); free(g_heap); free(orph);
    return R;
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: sim2 BVH8 RAYS tmin tmax [max_rays]\n"); return 1; }
    size_t bs, rs;
    const char* bvh = read_file(argv[1], &bs);
    const float* rf = read_file(argv[2], &rs);
    const unsigned* hdr = (const unsigned*)(bvh + 4 + 8);
    const unsigned nn = hdr[1];
    const Node8* nodes = (const Node8*)(bvh + 4 + 8 + 12);
    const Tri4* tris = (const Tri4*)((const char*)nodes + (size_t)nn * sizeof(Node8));
    int num_rays = (int)(rs / 24);
    if (argc > 5 && atoi(argv[5]) < num_rays) num_rays = atoi(argv[5]);
    pthread_once(&g_net_once, init_networks);
    int* perm = malloc(sizeof(int) * (size_t)num_rays);
    for (int i = 0; i < num_rays; i++) perm[i] = i;
    if (argc > 6) {                                        /* "mortonB": B bits per axis of the ray origin; stable */
        const int bits = atoi(argv[6] + 6);
        float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
        for (int i = 0; i < num_rays; i++) for (int c = 0; c < 3; c++) { const float v = rf[6 * i + c]; if (v < lo[c]) lo[c] = v; if (v > hi[c]) hi[c] = v; }
        unsigned* key = malloc(sizeof(unsigned) * (size_t)num_rays);
        for (int i = 0; i < num_rays; i++) {
            unsigned k = 0;
            unsigned q[3];
            for (int c = 0; c < 3; c++) { float f = hi[c] > lo[c] ? (rf[6 * i + c] - lo[c]) / (hi[c] - lo[c]) : 0; q[c] = (unsigned)(f * ((1u << bits) - 1)); }
            for (int b = bits - 1; b >= 0; b--) for (int c = 0; c < 3; c++) k = (k << 1) | ((q[c] >> b) & 1);
            const unsigned oct = (rf[6 * i + 3] > 0) | ((rf[6 * i + 4] > 0) << 1) | ((rf[6 * i + 5] > 0) << 2);
            if (strstr(argv[6], "+octlo")) k = (k << 3) | oct;            /* octant as the least significant bits */
            if (strstr(argv[6], "+octhi")) k |= oct << (3 * bits);         /* ... or the most significant ones */
            key[i] = k;
        }
        /* stable counting sort on the key, 11 bits at a time */
        int* tmp = malloc(sizeof(int) * (size_t)num_rays);
        for (int shift = 0; shift < 3 * bits + 3; shift += 11) {
            static int cnt[2049];
            memset(cnt, 0, sizeof cnt);
            for (int i = 0; i < num_rays; i++) cnt[((key[perm[i]] >> shift) & 2047) + 1]++;
            for (int b = 0; b < 2048; b++) cnt[b + 1] += cnt[b];
            for (int i = 0; i < num_rays; i++) tmp[cnt[(key[perm[i]] >> shift) & 2047]++] = perm[i];
            memcpy(perm, tmp, sizeof(int) * (size_t)num_rays);
        }
        fprintf(stderr, "rays reordered by a %d-bit Morton key of the origin\n", 3 * bits);
    }
    size_t* start = malloc(((size_t)num_rays + 1) * sizeof(size_t));
    const float tmin = (float)atof(argv[3]), tmax = (float)atof(argv[4]);
    for (int j = 0; j < num_rays; j++) {
        const int i = perm[j];
        Ray1 r = {{rf[6 * i], rf[6 * i + 1], rf[6 * i + 2]}, tmin, {rf[6 * i + 3], rf[6 * i + 4], rf[6 * i + 5]}, tmax};
        Hit1 h;
        start[j] = g_len; g_ray_first = g_len;
        traverse_single(8, 0, nodes, tris, &r, &h, NULL, NULL);
    }
    start[num_rays] = g_len;
    fprintf(stderr, "rays %d steps %zu (%.2f per ray)\n", num_rays, g_len, (double)g_len / num_rays);
    const Policy base = {24, 8, 0, 0, 0, 0, 1, 0, 1.95, 0.40, 20, 1, 1, 1, 1, 0, 0};
    struct { const char* name; Policy p; } cfg[64]; int nc = 0;
    cfg[nc].name = "current (refill 24, streak 8)"; cfg[nc++].p = base;
    { Policy p = base; p.refill_min = 16; cfg[nc].name = "refill 16"; cfg[nc++].p = p; }
    { Policy p = base; p.refill_min = 8; cfg[nc].name = "refill 8"; cfg[nc++].p = p; }
    { Policy p = base; p.sort_phase = 1; cfg[nc].name = "sort phase"; cfg[nc++].p = p; }
    { Policy p = base; p.chain_push = 1; cfg[nc].name = "chain push"; cfg[nc++].p = p; }
    { Policy p = base; p.chain_push = 1; p.sort_phase = 1; cfg[nc].name = "chain push + sort phase"; cfg[nc++].p = p; }
    { Policy p = base; p.orphan_k = 8; cfg[nc].name = "orphans k=8, 2nd launch thread per ray"; cfg[nc++].p = p; }
    { Policy p = base; p.orphan_k = 16; cfg[nc].name = "orphans k=16, 2nd launch thread per ray"; cfg[nc++].p = p; }
    { Policy p = base; p.orphan_k = 8; p.orphan_quad = 4; cfg[nc].name = "orphans k=8, 2nd launch quad per ray"; cfg[nc++].p = p; }
    { Policy p = base; p.orphan_k = 16; p.orphan_quad = 4; cfg[nc].name = "orphans k=16, 2nd launch quad per ray"; cfg[nc++].p = p; }
    { Policy p = base; p.orphan_k = 8; p.age_limit = 48; cfg[nc].name = "age 48 -> orphan queue, k=8"; cfg[nc++].p = p; }
    { Policy p = base; p.orphan_k = 8; p.age_limit = 32; cfg[nc].name = "age 32 -> orphan queue, k=8"; cfg[nc++].p = p; }
    { Policy p = base; p.chain_push = 1; p.sort_phase = 1; p.orphan_k = 8; p.orphan_quad = 4; cfg[nc].name = "chain + sort phase + orphans k=8 quad"; cfg[nc++].p = p; }
    { Policy p = base; p.k_push = 0; p.k_sort = 0; cfg[nc].name = "what if: pushes and sorts free"; cfg[nc++].p = p; }
    { Policy p = base; p.k_refill = 0; p.refill_min = 1; cfg[nc].name = "what if: refill free and immediate"; cfg[nc++].p = p; }
    { Policy p = base; p.k_refill = 0; p.refill_min = 1; p.k_push = 0; p.k_sort = 0; cfg[nc].name = "what if: both"; cfg[nc++].p = p; }
    { Policy p = base; p.refill_min = 12; p.streak_min = 8; cfg[nc].name = "refill 12"; cfg[nc++].p = p; }
    { Policy p = base; p.refill_min = 16; p.streak_min = 4; cfg[nc].name = "refill 16 streak 4"; cfg[nc++].p = p; }
    { Policy p = base; p.refill_min = 16; p.streak_min = 12; cfg[nc].name = "refill 16 streak 12"; cfg[nc++].p = p; }
    { Policy p = base; p.warps_per_sm = 24; p.r_sm = 2.1; cfg[nc].name = "24 warps per SM (r_sm 2.1)"; cfg[nc++].p = p; }
    { Policy p = base; p.coop_sort = 8; cfg[nc].name = "cooperative sort when <= 8 lanes need one"; cfg[nc++].p = p; }
    { Policy p = base; p.coop_sort = 16; cfg[nc].name = "cooperative sort when <= 16 lanes need one"; cfg[nc++].p = p; }
    { Policy p = base; p.coop_sort = 8; p.chain_push = 1; cfg[nc].name = "cooperative sort <= 8 + chain push"; cfg[nc++].p = p; }
    for (int m = 8; m <= 32; m += 8) { Policy p = base; p.postpone = m; char* nm = malloc(64); sprintf(nm, "postponed leaves, leaf step at >= %d lanes", m); cfg[nc].name = nm; cfg[nc++].p = p; }
    for (int m = 4; m <= 12; m += 2) { Policy p = base; p.sort_phase = 1; p.sort_min = m; p.chain_push = 1; char* nm = malloc(64); sprintf(nm, "chain push + sort phase at >= %d lanes", m); cfg[nc].name = nm; cfg[nc++].p = p; }
    for (int c = 0; c < nc; c++) {
        Result r = simulate(start, num_rays, &cfg[c].p);
        printf("%-44s %7.1f us (queue drained %6.1f, 1st launch %6.1f, orphans %6ld)  %6.1f Mrays/s  warp-inst %6.1f M  node %4.1f lanes  leaf %4.1f  sort %4.1f\n",
               cfg[c].name, r.time_us, r.drained_us, r.k1_us, r.orphans, num_rays / r.time_us, r.warp_inst / 1e6,
               r.node_lanes / r.node_exec, r.leaf_lanes / r.leaf_exec, r.sort_exec ? r.sort_lanes / r.sort_exec : 0.0);
    }
    return 0;
}
