/* ref_driver.c -- thin driver over the UNMODIFIED reference C library.
 *
 * TEST INFRASTRUCTURE ONLY (see jess_oracle.c).  Compiled by oracle/Makefile
 * together with the reference's own patched sources (taken from where they
 * lie under /root/reference, patched into a temporary directory with the
 * reference's src/scripts/apply_patch.py) into oracle/_ref/libjess_ref.so.
 * No reference source is stored in this repository.
 *
 * The driver turns the flat tables of include/jess_b200.h into the
 * reference's own structs exactly the way PyJess's Cython layer does
 * (Molecule.__init__ src/pyjess/_jess.pyx:632-684, TemplateAtom.__init__
 * :1399-1495, Template.__init__ :1825-1912, Jess.__init__ :2680-2712) and
 * then runs the reference query loop of Query.__next__ (:2197-2283):
 * JessQuery_next -> JessQuery_superposition -> rmsd filter -> NaN drop.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "Atom.h"
#include "Jess.h"
#include "Molecule.h"
#include "ResIndex.h"
#include "Super.h"
#include "Template.h"
#include "TessAtom.h"
#include "TessTemplate.h"

#include "hitbuf.h"

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static TessAtom *make_tess_atom(const jb_template_table *T, int ta)
{
    int ac = T->name_offset[ta + 1] - T->name_offset[ta];
    int rc = T->resname_offset[ta + 1] - T->resname_offset[ta];
    size_t rq = sizeof(TessAtom) + sizeof(char *) * (size_t)(ac + rc) + (size_t)(5 * ac + 4 * rc);
    TessAtom *A = (TessAtom *)calloc(1, rq);
    if (!A) return NULL;
    A->code = T->match_mode[ta];
    A->resSeq = T->residue_number[ta];
    A->nameCount = ac;
    A->resNameCount = rc;
    A->chainID1 = ' ';
    A->chainID2 = (char)T->chain_id2[ta];
    for (int j = 0; j < 3; j++) A->pos[j] = T->position[3 * ta + j];
    A->distWeight = T->distance_weight[ta];
    char *p = (char *)&A[1];
    A->name = (char **)p; p += sizeof(char *) * (size_t)ac;
    for (int m = 0; m < ac; m++) { A->name[m] = p; memcpy(p, &T->atom_names[T->name_offset[ta] + m], 4); p[4] = 0; p += 5; }
    A->resName = (char **)p; p += sizeof(char *) * (size_t)rc;
    for (int m = 0; m < rc; m++) { A->resName[m] = p; memcpy(p, &T->residue_names[T->resname_offset[ta] + m], 4); p += 4; }
    return A;
}

static Template *make_template(const jb_template_table *T, int t)
{
    int ta0 = T->atom_offset[t], n = T->atom_offset[t + 1] - ta0;
    size_t rq = sizeof(Template) + sizeof(TessTemplate) + (size_t)n * sizeof(TessAtom *)
              + (size_t)n * sizeof(double *) + (size_t)n * (size_t)n * sizeof(double);
    Template *tpl = (Template *)calloc(1, rq);
    if (!tpl) return NULL;
    TessTemplate *tess = (TessTemplate *)&tpl[1];
    tess->atom = (TessAtom **)&tess[1];
    tess->distance = (double **)&tess->atom[n];
    if (n > 0) tess->distance[0] = (double *)&tess->distance[n];
    for (int i = 1; i < n; i++) tess->distance[i] = tess->distance[i - 1] + n;
    tpl->free = TessTemplate_free;
    tpl->match = TessTemplate_match;
    tpl->position = TessTemplate_position;
    tpl->count = TessTemplate_count;
    tpl->range = TessTemplate_range;
    tpl->check = TessTemplate_check;
    tpl->candidates = TessTemplate_candidates;
    tpl->name = TessTemplate_name;
    tpl->logE = TessTemplate_logE;
    tpl->distWeight = TessTemplate_distWeight;
    tpl->copy = TessTemplate_copy;
    tess->count = n;
    tess->symbol = NULL;
    tess->dim = 0; /* only used by logE, which the Python layer evaluates */
    for (int i = 0; i < n; i++) {
        tess->atom[i] = make_tess_atom(T, ta0 + i);
        if (!tess->atom[i]) return NULL;
    }
    const double *d = T->distances + T->distance_offset[t];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) tess->distance[i][j] = d[i * n + j];
    return tpl;
}

typedef struct { Molecule *mol; Atom *atoms; } ref_mol;

static int make_molecule(const jb_molecule_batch *B, int m, ref_mol *out)
{
    int64_t a0 = B->atom_offset[m];
    int n = (int)(B->atom_offset[m + 1] - a0);
    Molecule *M = (Molecule *)calloc(1, sizeof(Molecule) + (size_t)n * sizeof(Atom *));
    Atom *A = (Atom *)calloc((size_t)(n > 0 ? n : 1), sizeof(Atom));
    if (!M || !A) return -1;
    for (int i = 0; i < n; i++) {
        Atom *a = &A[i];
        a->x[0] = B->x[a0 + i]; a->x[1] = B->y[a0 + i]; a->x[2] = B->z[a0 + i];
        a->serial = B->serial ? B->serial[a0 + i] : i + 1; /* ResIndex sorts by (resName, serial) */
        memcpy(a->name, &B->name[a0 + i], 4); a->name[4] = 0;
        memcpy(a->resName, &B->residue_name[a0 + i], 4);
        a->altLoc = ' '; a->iCode = ' ';
        a->chainID1 = ' ';
        a->chainID2 = (char)B->chain_id2[a0 + i];
        a->resSeq = B->residue_number[a0 + i];
        M->atom[i] = a;
    }
    M->count = n;
    M->index = ResIndex_create(M->atom, n);
    if (!M->index) return -1;
    out->mol = M; out->atoms = A;
    return 0;
}

static void free_molecule(ref_mol *r)
{
    if (r->mol) { if (r->mol->index) ResIndex_free(r->mol->index); free(r->mol); }
    free(r->atoms);
    r->mol = NULL; r->atoms = NULL;
}

/* one molecule against every template: the loop of Query.__next__ */
static int query_molecule(Jess *J, Template **tpls, int T, const ref_mol *rm, int m,
                          const jb_query_params *Q, oracle_result *R, jb_stats *st, char *rec_pool)
{
    hit_rec *h = (hit_rec *)rec_pool, *best = (hit_rec *)(rec_pool + R->rec_size);
    JessQuery *q = Jess_query(J, rm->mol, Q->distance_cutoff, Q->max_dynamic_distance, Q->reorder != 0);
    if (!q) return -1;
    Template *best_tpl = NULL; int have_best = 0, last = 0;
    IgnoreType ig = Q->ignore_chain == JB_IGNORE_ATOMS ? ignoreAtoms
                  : Q->ignore_chain == JB_IGNORE_RESIDUES ? ignoreResidues : ignoreNone;
    int rc = 0;
    while (JessQuery_next(q, ig)) {
        Template *tpl = JessQuery_template(q);
        if (Q->best_match && have_best && tpl != best_tpl) {
            st->hits++; if (res_push(R, best) != 0) rc = -1;
            have_best = 0;
        }
        int n = tpl->count(tpl), t;
        if (tpls[last] == tpl) t = last;
        else { for (t = 0; t < T && tpls[t] != tpl; t++); last = t; }
        Atom **A = JessQuery_atoms(q);
        Superposition *sup = JessQuery_superposition(q);
        double rmsd = Superposition_rmsd(sup);
        st->tuples++;
        double bound = (Q->best_match && have_best) ? best->rmsd : INFINITY;
        if (rmsd <= Q->rmsd_threshold && rmsd < bound) {
            const double *rot = Superposition_rotation(sup);
            int nan = 0;
            for (int i = 0; i < 9; i++) nan |= isnan(rot[i]);
            if (nan) st->nan_dropped++;
            else {
                h->mol = m; h->tpl = t; h->rmsd = rmsd;
                memcpy(h->rot, rot, sizeof h->rot);
                memcpy(h->c0, Superposition_centroid(sup, 0), sizeof h->c0);
                memcpy(h->c1, Superposition_centroid(sup, 1), sizeof h->c1);
                for (int k = 0; k < R->stride; k++) h->atoms[k] = k < n ? (int32_t)(A[k] - rm->atoms) : -1;
                if (Q->best_match) { memcpy(best, h, R->rec_size); best_tpl = tpl; have_best = 1; }
                else { st->hits++; if (res_push(R, h) != 0) rc = -1; }
            }
        }
        Superposition_free(sup);
    }
    if (Q->best_match && have_best) { st->hits++; if (res_push(R, best) != 0) rc = -1; }
    JessQuery_free(q);
    return rc;
}

int ref_query(const jb_template_table *T, const jb_molecule_batch *B, const jb_query_params *Q,
              int n_threads, oracle_result **out)
{
    int nt = T->n_templates, stride = 1;
    for (int t = 0; t < nt; t++) {
        int n = T->atom_offset[t + 1] - T->atom_offset[t];
        if (n > stride) stride = n;
    }
    oracle_result *R = hb_create(stride);
    if (!R) return JB_ERR_NOMEM;
    Jess *J = Jess_create();
    Template **tpls = (Template **)calloc((size_t)(nt > 0 ? nt : 1), sizeof(Template *));
    ref_mol *mols = (ref_mol *)calloc((size_t)(B->n_molecules > 0 ? B->n_molecules : 1), sizeof(ref_mol));
    int failed = (!J || !tpls || !mols);
    for (int t = 0; t < nt && !failed; t++) {
        tpls[t] = make_template(T, t);
        if (!tpls[t]) failed = 1; else Jess_addTemplate(J, tpls[t]);
    }
    for (int m = 0; m < B->n_molecules && !failed; m++)
        if (make_molecule(B, m, &mols[m]) != 0) failed = 1;

    double t0 = now_ms();
    /* hit_order != canonical: keep the reference's own iteration order (structures ascending,
     * templates in reverse insertion order, Scanner_next order inside) -- one thread, no sort */
    const int ordered = Q->hit_order != 0;
    if (ordered) n_threads = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
#endif
    {
        oracle_result L = *R;
        L.count = L.cap = 0; L.recs = NULL;
        jb_stats st; memset(&st, 0, sizeof st);
        char *rec_pool = (char *)malloc(2 * R->rec_size);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int m = 0; m < B->n_molecules; m++) {
            if (failed || !rec_pool) { failed = 1; continue; }
            if (query_molecule(J, tpls, nt, &mols[m], m, Q, &L, &st, rec_pool) != 0) failed = 1;
        }
        free(rec_pool);
#ifdef _OPENMP
#pragma omp critical
#endif
        {
            for (int64_t i = 0; i < L.count && !failed; i++)
                if (res_push(R, (hit_rec *)(L.recs + (size_t)i * R->rec_size)) != 0) failed = 1;
            R->stats.tuples += st.tuples; R->stats.hits += st.hits; R->stats.nan_dropped += st.nan_dropped;
        }
        free(L.recs);
    }
    R->stats.kernel_ms = now_ms() - t0; /* query loop only: no struct building, no sort */
    R->stats.pairs = (int64_t)nt * B->n_molecules;

    for (int m = 0; m < B->n_molecules; m++) free_molecule(&mols[m]);
    free(mols);
    if (J) Jess_free(J); /* frees the templates it owns */
    free(tpls);
    if (failed) { hb_free(R); return JB_ERR_NOMEM; }
    if (!ordered) hb_sort(R);
    *out = R;
    return JB_OK;
}

int ref_result_hits(oracle_result *R, jb_hits *H) { return hb_hits(R, H); }
int ref_result_stats(oracle_result *R, jb_stats *S) { *S = R->stats; return JB_OK; }
void ref_result_free(oracle_result *R) { hb_free(R); }
