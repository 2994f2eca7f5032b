/* ORACLE (test infrastructure, never shipped): single-threaded CPU marching cubes.
 *
 * Restates the contract of the reference's surface extraction call
 *   skimage.measure.marching_cubes(vol, level, method="lewiner")
 * (reference hy3dgen/shapegen/models/autoencoders/surface_extractors.py:69-73 and
 * project/image3d/__init__.py:61-65).  scikit-image is an un-vendored, un-pinned
 * third-party dependency that is absent from /root/reference and from this image
 * (requirements.txt:31 has it commented out), and the reference has no test or
 * golden vector at this boundary:  **PARITY UNPINNED**  (SURVEY §0.1, §8c, App. D).
 * What is restated, from the published algorithm (Lorensen & Cline 1987; Lewiner
 * et al. 2003 for the welded, indexed output):
 *   - cube corner order / inside test  "v - level > 0"  (NaN => outside);
 *   - one shared vertex per sign-change grid edge, placed at the inverse-distance
 *     weighted average of the edge end points, w = 1/(FLT_EPSILON + |v - level|);
 *   - vertex columns in array-axis order (axis0, axis1, axis2), int32 faces.
 * Deliberate, documented difference: ambiguous faces use one fixed sign-only
 * rule (inside corners separated) instead of Lewiner's value-based tests, so the
 * result is watertight but may triangulate ambiguous cubes differently.
 *
 * This file does NOT use the generated table of the product
 * (hunyuan3d-2_b200/csrc/mc_tables.inc): triangles are derived per cube at run
 * time by tracing oriented iso-contour loops over the cube faces.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const int CORNER[8][3] = { /* dx (axis2), dy (axis1), dz (axis0) */
    {0,0,0},{1,0,0},{1,1,0},{0,1,0},{0,0,1},{1,0,1},{1,1,1},{0,1,1}};
static const int EDGE[12][2] = {{0,1},{1,2},{2,3},{3,0},{4,5},{5,6},{6,7},{7,4},{0,4},{1,5},{2,6},{3,7}};
/* corner cycles of the six faces, counter-clockwise seen from outside */
static const int FACE[6][4] = {{0,4,7,3},{1,2,6,5},{0,1,5,4},{2,3,7,6},{0,3,2,1},{4,5,6,7}};

static int edge_between(int a, int b) {
    for (int e = 0; e < 12; ++e)
        if ((EDGE[e][0] == a && EDGE[e][1] == b) || (EDGE[e][0] == b && EDGE[e][1] == a)) return e;
    return -1;
}

/* Triangles of one cube as edge ids; returns the count (<= 5). */
static int cube_triangles(int cs, int tri[5][3]) {
    int nxt[12];
    for (int e = 0; e < 12; ++e) nxt[e] = -1;
    for (int f = 0; f < 6; ++f) {
        int in[4], ed[4];
        for (int m = 0; m < 4; ++m) {
            in[m] = (cs >> FACE[f][m]) & 1;
            ed[m] = edge_between(FACE[f][m], FACE[f][(m + 1) & 3]);
        }
        int nio = 0, io[2], oi[2], noi = 0;
        for (int m = 0; m < 4; ++m) {
            if (in[m] && !in[(m + 1) & 3]) io[nio++] = m;
            if (!in[m] && in[(m + 1) & 3]) oi[noi++] = m;
        }
        if (nio == 1) nxt[ed[io[0]]] = ed[oi[0]];
        else if (nio == 2)              /* ambiguous face: cut off each inside corner */
            for (int q = 0; q < 2; ++q) nxt[ed[io[q]]] = ed[(io[q] + 3) & 3];
    }
    int seen[12] = {0}, nt = 0;
    for (int s = 0; s < 12; ++s) {
        if (nxt[s] < 0 || seen[s]) continue;
        int loop[12], n = 0, e = s;
        while (!seen[e]) { seen[e] = 1; loop[n++] = e; e = nxt[e]; }
        for (int m = 1; m + 1 < n; ++m) {          /* fan, reversed winding */
            tri[nt][0] = loop[0]; tri[nt][1] = loop[m + 1]; tri[nt][2] = loop[m]; ++nt;
        }
    }
    return nt;
}

static inline int inside(float v, float level) { return (v - level) > 0.0f; }

/* vertex position along an edge, index units, per Appendix D */
static inline float interp(int base, float va, float vb, float level) {
    double a = (double)va - (double)level, b = (double)vb - (double)level;
    double wa = 1.0 / ((double)FLT_EPSILON + fabs(a));
    double wb = 1.0 / ((double)FLT_EPSILON + fabs(b));
    return (float)((double)base + wb / (wa + wb));
}

/* Returns 0 on success.  verts: [nV,3] float32 (axis0,axis1,axis2 index units),
 * faces: [nF,3] int32.  Buffers are malloc'ed; release with hy3d_oracle_free. */
int hy3d_oracle_mc(const float* vol, int n0, int n1, int n2, float level,
                   float** verts_out, int64_t* nV_out, int32_t** faces_out, int64_t* nF_out) {
    const int64_t s0 = (int64_t)n1 * n2, s1 = n2;
    const int64_t nvox = (int64_t)n0 * n1 * n2;
    int32_t* eid = (int32_t*)malloc(sizeof(int32_t) * 3 * (size_t)nvox);   /* edge -> vertex id */
    if (!eid) return -1;
    int64_t nV = 0;
    /* pass 1: vertices in lexicographic (voxel, axis) order */
    for (int i = 0; i < n0; ++i) for (int j = 0; j < n1; ++j) for (int k = 0; k < n2; ++k) {
        int64_t p = i * s0 + j * s1 + k;
        int a = inside(vol[p], level);
        eid[3*p+0] = (i + 1 < n0 && inside(vol[p + s0], level) != a) ? (int32_t)nV++ : -1;
        eid[3*p+1] = (j + 1 < n1 && inside(vol[p + s1], level) != a) ? (int32_t)nV++ : -1;
        eid[3*p+2] = (k + 1 < n2 && inside(vol[p + 1],  level) != a) ? (int32_t)nV++ : -1;
    }
    float* verts = (float*)malloc(sizeof(float) * 3 * (size_t)(nV > 0 ? nV : 1));
    for (int i = 0; i < n0; ++i) for (int j = 0; j < n1; ++j) for (int k = 0; k < n2; ++k) {
        int64_t p = i * s0 + j * s1 + k;
        int32_t v;
        if ((v = eid[3*p+0]) >= 0) { verts[3*v] = interp(i, vol[p], vol[p+s0], level); verts[3*v+1] = (float)j; verts[3*v+2] = (float)k; }
        if ((v = eid[3*p+1]) >= 0) { verts[3*v] = (float)i; verts[3*v+1] = interp(j, vol[p], vol[p+s1], level); verts[3*v+2] = (float)k; }
        if ((v = eid[3*p+2]) >= 0) { verts[3*v] = (float)i; verts[3*v+1] = (float)j; verts[3*v+2] = interp(k, vol[p], vol[p+1], level); }
    }
    /* pass 2: faces in lexicographic cube order */
    int64_t capF = 1024, nF = 0;
    int32_t* faces = (int32_t*)malloc(sizeof(int32_t) * 3 * (size_t)capF);
    for (int i = 0; i + 1 < n0; ++i) for (int j = 0; j + 1 < n1; ++j) for (int k = 0; k + 1 < n2; ++k) {
        int cs = 0;
        for (int m = 0; m < 8; ++m) {
            int64_t p = (i + CORNER[m][2]) * s0 + (j + CORNER[m][1]) * s1 + (k + CORNER[m][0]);
            cs |= inside(vol[p], level) << m;
        }
        if (cs == 0 || cs == 255) continue;
        int tri[5][3];
        int nt = cube_triangles(cs, tri);
        if (nF + nt > capF) { capF *= 2; faces = (int32_t*)realloc(faces, sizeof(int32_t) * 3 * (size_t)capF); }
        for (int t = 0; t < nt; ++t) {
            for (int c = 0; c < 3; ++c) {
                int e = tri[t][c];
                const int* ca = CORNER[EDGE[e][0]]; const int* cb = CORNER[EDGE[e][1]];
                int ox = ca[0] < cb[0] ? ca[0] : cb[0], oy = ca[1] < cb[1] ? ca[1] : cb[1], oz = ca[2] < cb[2] ? ca[2] : cb[2];
                int axis = (ca[2] != cb[2]) ? 0 : (ca[1] != cb[1]) ? 1 : 2;     /* array axis */
                int64_t p = (i + oz) * s0 + (j + oy) * s1 + (k + ox);
                faces[3*nF + c] = eid[3*p + axis];
            }
            ++nF;
        }
    }
    free(eid);
    *verts_out = verts; *nV_out = nV; *faces_out = faces; *nF_out = nF;
    return 0;
}

void hy3d_oracle_free(void* p) { free(p); }

/* 8-bit cube case index per cube (table-independent classification), [n0-1,n1-1,n2-1] uint8. */
void hy3d_oracle_mc_cases(const float* vol, int n0, int n1, int n2, float level, uint8_t* out) {
    const int64_t s0 = (int64_t)n1 * n2, s1 = n2;
    int64_t q = 0;
    for (int i = 0; i + 1 < n0; ++i) for (int j = 0; j + 1 < n1; ++j) for (int k = 0; k + 1 < n2; ++k) {
        int cs = 0;
        for (int m = 0; m < 8; ++m) {
            int64_t p = (i + CORNER[m][2]) * s0 + (j + CORNER[m][1]) * s1 + (k + CORNER[m][0]);
            cs |= inside(vol[p], level) << m;
        }
        out[q++] = (uint8_t)cs;
    }
}
