/* hitbuf.h -- growable hit container shared by the oracle and the reference
 * driver (TEST INFRASTRUCTURE ONLY, see jess_oracle.c).  Not reference code. */
#ifndef ORACLE_HITBUF_H
#define ORACLE_HITBUF_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/jess_b200.h"

/* ------------------------------------------------------------------------ */

typedef struct {
    int32_t mol, tpl;
    double rmsd, rot[9], c0[3], c1[3];
    int32_t atoms[1]; /* stride entries */
} hit_rec;

typedef struct oracle_result {
    int stride;
    size_t rec_size;
    int64_t count, cap;
    char *recs;
    jb_stats stats;
    /* SoA views handed out by oracle_result_hits */
    int32_t *mol, *tpl, *atoms;
    double *rmsd, *rot, *c0, *c1;
} oracle_result;

static int res_push(oracle_result *R, const hit_rec *h)
{
    if (R->count == R->cap) {
        int64_t nc = R->cap ? R->cap * 2 : 1024;
        char *p = (char *)realloc(R->recs, (size_t)nc * R->rec_size);
        if (!p) return -1;
        R->recs = p; R->cap = nc;
    }
    memcpy(R->recs + (size_t)R->count * R->rec_size, h, R->rec_size);
    R->count++;
    return 0;
}

static int g_stride;
static int rec_cmp(const void *pa, const void *pb)
{
    const hit_rec *a = (const hit_rec *)pa, *b = (const hit_rec *)pb;
    if (a->mol != b->mol) return a->mol < b->mol ? -1 : 1;
    if (a->tpl != b->tpl) return a->tpl < b->tpl ? -1 : 1;
    for (int k = 0; k < g_stride; k++)
        if (a->atoms[k] != b->atoms[k]) return a->atoms[k] < b->atoms[k] ? -1 : 1;
    return 0;
}


static void hb_sort(oracle_result *R)
{
    g_stride = R->stride;
    qsort(R->recs, (size_t)R->count, R->rec_size, rec_cmp);
}

static oracle_result *hb_create(int stride)
{
    oracle_result *R = (oracle_result *)calloc(1, sizeof *R);
    if (!R) return NULL;
    R->stride = stride;
    R->rec_size = (sizeof(hit_rec) + sizeof(int32_t) * (size_t)stride + 7) & ~(size_t)7;
    return R;
}

static int hb_hits(oracle_result *R, jb_hits *H)
{
    int64_t n = R->count;
    size_t m = (size_t)(n > 0 ? n : 1);
    if (!R->mol) {
        R->mol = (int32_t *)malloc(m * sizeof(int32_t));
        R->tpl = (int32_t *)malloc(m * sizeof(int32_t));
        R->atoms = (int32_t *)malloc(m * sizeof(int32_t) * (size_t)R->stride);
        R->rmsd = (double *)malloc(m * sizeof(double));
        R->rot = (double *)malloc(m * 9 * sizeof(double));
        R->c0 = (double *)malloc(m * 3 * sizeof(double));
        R->c1 = (double *)malloc(m * 3 * sizeof(double));
        if (!R->mol || !R->tpl || !R->atoms || !R->rmsd || !R->rot || !R->c0 || !R->c1) return JB_ERR_NOMEM;
        for (int64_t i = 0; i < n; i++) {
            const hit_rec *h = (const hit_rec *)(R->recs + (size_t)i * R->rec_size);
            R->mol[i] = h->mol; R->tpl[i] = h->tpl; R->rmsd[i] = h->rmsd;
            memcpy(R->atoms + (size_t)i * R->stride, h->atoms, sizeof(int32_t) * (size_t)R->stride);
            memcpy(R->rot + 9 * i, h->rot, sizeof h->rot);
            memcpy(R->c0 + 3 * i, h->c0, sizeof h->c0);
            memcpy(R->c1 + 3 * i, h->c1, sizeof h->c1);
        }
    }
    H->count = n; H->stride = R->stride; H->reserved = 0;
    H->molecule = R->mol; H->template_index = R->tpl; H->atoms = R->atoms;
    H->rmsd = R->rmsd; H->rotation = R->rot; H->centre0 = R->c0; H->centre1 = R->c1;
    return JB_OK;
}

static void hb_free(oracle_result *R)
{
    if (!R) return;
    free(R->recs); free(R->mol); free(R->tpl); free(R->atoms);
    free(R->rmsd); free(R->rot); free(R->c0); free(R->c1);
    free(R);
}
#endif
