/* Developer tool: writes the number of traversal steps (inner-node visits + Tri4 packets) of every ray of a set
 * as uint32, for tail studies.  build: gcc -O2 -march=x86-64-v3 -ffp-contract=off -o /tmp/ray_steps scripts/ray_steps.c -lpthread -lm
 * usage: ray_steps BVH8_FILE RAYS_FILE tmin tmax OUT */
#include <stdio.h>
static unsigned g_steps;
#define ORACLE_TRACE(kind) do { if ((kind) == 'N' || (kind) == 'L') g_steps++; } while (0)
#include "../oracle/traversal_oracle.c"
static void* read_file(const char* path, size_t* size) { FILE* f = fopen(path, "rb"); if (!f) { perror(path); exit(1); } fseek(f, 0, SEEK_END); *size = (size_t)ftell(f); fseek(f, 0, SEEK_SET); void* p = malloc(*size); if (fread(p, 1, *size, f) != *size) exit(1); fclose(f); return p; }
int main(int argc, char** argv) {
    if (argc < 6) return 1;
    size_t bs, rs; const char* bvh = read_file(argv[1], &bs); const float* rf = read_file(argv[2], &rs);
    const unsigned* hdr = (const unsigned*)(bvh + 12); const Node8* nodes = (const Node8*)(bvh + 24);
    const Tri4* tris = (const Tri4*)((const char*)nodes + (size_t)hdr[1] * 256);
    const int n = (int)(rs / 24); pthread_once(&g_net_once, init_networks);
    unsigned* s = malloc((size_t)n * 4);
    for (int i = 0; i < n; i++) {
        Ray1 r = {{rf[6*i], rf[6*i+1], rf[6*i+2]}, (float)atof(argv[3]), {rf[6*i+3], rf[6*i+4], rf[6*i+5]}, (float)atof(argv[4])}; Hit1 h;
        g_steps = 0; traverse_single(8, 0, nodes, tris, &r, &h, NULL, NULL); s[i] = g_steps;
    }
    FILE* f = fopen(argv[5], "wb"); fwrite(s, 4, (size_t)n, f); fclose(f);
    return 0;
}
