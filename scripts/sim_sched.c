/* Scheduling study for the traversal kernel (developer tool, CPU only).
 *
 * Records, for every ray of a ray set, the sequence of traversal steps the reference order
 * takes ('N' = one inner-node visit, 'L' = one Tri4 packet), then replays those sequences
 * through warp-scheduling policies to predict SIMT lane utilisation:
 *   K     rays held per lane (lane-private slots; K = 1 is thread-per-ray)
 *   policy 0: while-while (nodes until no lane has one, then leaves until none)
 *   policy 1: majority vote between the node and the leaf step
 *   policy 2: node step unless >= TL lanes wait with a leaf or < TN lanes have a node
 * build: gcc -O2 -march=x86-64-v3 -ffp-contract=off -o /tmp/sim_sched scripts/sim_sched.c -lpthread -lm
 * usage: sim_sched BVH8_FILE RAYS_FILE tmin tmax
 */
#include <stdio.h>
static unsigned char* g_trace; static size_t g_trace_len, g_trace_cap;
static void trace_put(int c);
#define ORACLE_TRACE(kind) trace_put(kind)
#include "../oracle/traversal_oracle.c"

static void trace_put(int c) {
    if (g_trace_len == g_trace_cap) { g_trace_cap = g_trace_cap ? g_trace_cap * 2 : (1u << 24); g_trace = realloc(g_trace, g_trace_cap); }
    g_trace[g_trace_len++] = (unsigned char)c;
}

static void* read_file(const char* path, size_t* size) {
    FILE* f = fopen(path, "rb"); if (!f) { perror(path); exit(1); }
    fseek(f, 0, SEEK_END); *size = (size_t)ftell(f); fseek(f, 0, SEEK_SET);
    void* p = malloc(*size); if (fread(p, 1, *size, f) != *size) exit(1); fclose(f); return p;
}

enum { COST_N = 270, COST_L = 250, COST_FETCH = 60 };

typedef struct { double issued, useful, n_steps, l_steps, n_lanes, l_lanes; } SimResult;

static SimResult simulate(const size_t* start, int num_rays, int K, int policy, int TN, int TL, int num_warps) {
    SimResult r = {0};
    typedef struct { size_t pos, end; } Slot;
    Slot* slots = calloc((size_t)num_warps * 32 * K, sizeof(Slot));
    int next_ray = 0, live_warps = num_warps;
    char* mode = calloc(num_warps, 1);            /* policy 0: current phase of the warp */
    char* dead = calloc(num_warps, 1);
    while (live_warps > 0) {
        for (int w = 0; w < num_warps; w++) {
            if (dead[w]) continue;
            Slot* s = slots + (size_t)w * 32 * K;
            /* refill empty slots */
            int empties = 0, occupied = 0;
            for (int i = 0; i < 32 * K; i++) { if (s[i].pos == s[i].end) empties++; else occupied++; }
            if (empties && next_ray < num_rays) {
                for (int i = 0; i < 32 * K && next_ray < num_rays; i++)
                    if (s[i].pos == s[i].end) {
                        s[i].pos = start[next_ray]; s[i].end = start[next_ray + 1]; next_ray++;
                        if (s[i].pos != s[i].end) occupied++;
                    }
                r.issued += COST_FETCH;
            }
            if (!occupied) { dead[w] = 1; live_warps--; continue; }
            int cn = 0, cl = 0;
            for (int l = 0; l < 32; l++) {
                int hn = 0, hl = 0;
                for (int k = 0; k < K; k++) { const Slot* q = &s[l * K + k]; if (q->pos != q->end) { if (g_trace[q->pos] == 'N') hn = 1; else hl = 1; } }
                cn += hn; cl += hl;
            }
            int phase;
            if (policy == 0) { if (mode[w] == 0 && cn == 0) mode[w] = 1; else if (mode[w] == 1 && cl == 0) mode[w] = 0; phase = mode[w]; }
            else if (policy == 1) phase = cl > cn;
            else phase = (cl >= TL || cn < TN) && cl > 0;
            if (phase == 0 && cn == 0) phase = 1;
            if (phase == 1 && cl == 0) phase = 0;
            const unsigned char want = phase ? 'L' : 'N';
            int act = 0;
            for (int l = 0; l < 32; l++)
                for (int k = 0; k < K; k++) { Slot* q = &s[l * K + k]; if (q->pos != q->end && g_trace[q->pos] == want) { q->pos++; act++; break; } }
            const double cost = phase ? COST_L : COST_N;
            r.issued += cost; r.useful += cost * act / 32.0;
            if (phase) { r.l_steps++; r.l_lanes += act; } else { r.n_steps++; r.n_lanes += act; }
        }
    }
    free(slots); free(mode); free(dead);
    return r;
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: sim_sched BVH8 RAYS tmin tmax\n"); return 1; }
    size_t bs, rs;
    const char* bvh = read_file(argv[1], &bs);
    const float* rf = read_file(argv[2], &rs);
    /* single-block .bvh written by this repo: magic u32, then [u64 size][u32 type][u32 nodes][u32 tris] */
    const unsigned* hdr = (const unsigned*)(bvh + 4 + 8);
    const unsigned nn = hdr[1], nt = hdr[2];
    const Node8* nodes = (const Node8*)(bvh + 4 + 8 + 12);
    const Tri4* tris = (const Tri4*)((const char*)nodes + (size_t)nn * sizeof(Node8));
    const int num_rays = (int)(rs / 24);
    fprintf(stderr, "type %u nodes %u tris %u rays %d\n", hdr[0], nn, nt, num_rays);
    pthread_once(&g_net_once, init_networks);
    size_t* start = malloc(((size_t)num_rays + 1) * sizeof(size_t));
    const float tmin = (float)atof(argv[3]), tmax = (float)atof(argv[4]);
    for (int i = 0; i < num_rays; i++) {
        Ray1 r = {{rf[6 * i], rf[6 * i + 1], rf[6 * i + 2]}, tmin, {rf[6 * i + 3], rf[6 * i + 4], rf[6 * i + 5]}, tmax};
        Hit1 h;
        start[i] = g_trace_len;
        traverse_single(8, 0, nodes, tris, &r, &h, NULL, NULL);
    }
    start[num_rays] = g_trace_len;
    size_t nN = 0; for (size_t i = 0; i < g_trace_len; i++) nN += g_trace[i] == 'N';
    printf("steps per ray: N %.2f L %.2f\n", (double)nN / num_rays, (double)(g_trace_len - nN) / num_rays);
    const double ideal = ((double)nN * COST_N + (double)(g_trace_len - nN) * COST_L) / 32.0;
    const int warps = 148 * 16;
    struct { int K, policy, TN, TL; } cfg[] = {
        {1, 0, 0, 0}, {1, 1, 0, 0}, {1, 2, 16, 24}, {2, 0, 0, 0}, {2, 1, 0, 0}, {2, 2, 24, 24}, {2, 2, 28, 28}, {2, 2, 31, 30},
        {3, 1, 0, 0}, {3, 2, 28, 28}, {3, 2, 31, 30}, {4, 1, 0, 0}, {4, 2, 28, 28}, {4, 2, 31, 30}, {4, 2, 32, 32}, {6, 2, 32, 32}, {8, 2, 32, 32},
    };
    for (size_t c = 0; c < sizeof cfg / sizeof cfg[0]; c++) {
        const int nw = warps / cfg[c].K > 148 ? warps / cfg[c].K : 148;
        SimResult r = simulate(start, num_rays, cfg[c].K, cfg[c].policy, cfg[c].TN, cfg[c].TL, nw);
        printf("K %d policy %d TN %2d TL %2d warps %4d: efficiency %.3f (ideal/issued %.3f)  N lanes %.1f  L lanes %.1f  warp-inst/ray %.0f\n",
               cfg[c].K, cfg[c].policy, cfg[c].TN, cfg[c].TL, nw, r.useful / r.issued, ideal / r.issued,
               r.n_lanes / r.n_steps, r.l_lanes / r.l_steps, r.issued / num_rays);
    }
    return 0;
}
