/* amod_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into, imported by, or called from
 * the product path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker or as the timed CPU baseline).
 *
 * Plain-C restatement of the reference's OSP/SBA/NPO hot path (Leot6/AMoD, pure Python).
 * It follows the reference's loops literally — same iteration order, same early exits, same
 * left-to-right float64 accumulation — so that it can stand in for the reference on a box
 * where the reference cannot travel.  Pinned against the reference itself: tests/golden/
 * holds pools produced by running the unmodified reference in the build container
 * (tests/golden/make_golden.py) and tests/test_oracle_golden.py checks this file against
 * them.  The reference has no tests or known-answer vectors of its own.
 *
 * Reference lines followed:
 *   test_constraints()        scheduling.py:70-98
 *   compute_schedule()        scheduling.py:13-45 (+ insert_req_into_sche :49-65)
 *   sche_cost()/sche_delay()  scheduling.py:101-107 / :126-137
 *   basic_sches()             dispatch_osp.py:295-322
 *   build_trip_table()        dispatch_osp.py:132-228   (wall-clock cut-offs :219-224 disabled)
 *   oracle_osp_build()        dispatch_osp.py:94-129
 *   oracle_sba_build()        dispatch_sba.py:44-74
 *   oracle_npo_match()        rebalancing_npo.py:13-59
 */
#include "../include/amod_b200.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXS AMOD_MAX_STOPS

typedef struct {
    const amod_epoch* ep;
    const float* tt32;
    const double* tt64;
    int64_t n_nodes;
    int64_t n_screens, n_evals, n_lookups, n_tasks, n_sches;
} octx;

static inline double TT(octx* c, int o, int d) { /* route_functions.py:22-25, 1-based ids */
    c->n_lookups++;
    int64_t k = (int64_t)(o - 1) * c->n_nodes + (d - 1);
    return c->tt64 ? c->tt64[k] : (double)c->tt32[k];
}

static inline void stop_info(const amod_epoch* ep, int v, int code, int* node, double* ddl, int* pod) {
    if (code < 0) { *node = ep->veh_reb_node[v]; *ddl = ep->veh_reb_ddl[v]; *pod = 0; return; }
    int r = code >> 1;
    if (code & 1) { *node = ep->req_dnid[r]; *ddl = ep->req_cld[r]; *pod = -1; }
    else          { *node = ep->req_onid[r]; *ddl = ep->req_clp[r]; *pod = 1; }
}

/* scheduling.py:70-98.  Returns the violation code (-1 = feasible, cost in *cost). */
static int test_constraints(octx* c, int v, const int* sche, int len, int new_req, int idx_p, int idx_d,
                            double* cost) {
    const amod_epoch* ep = c->ep;
    c->n_evals++;
    int n = ep->veh_load[v];
    for (int i = 0; i < len; i++) {
        int node, pod; double ddl;
        stop_info(ep, v, sche[i], &node, &ddl, &pod);
        n += pod;
        if (n > ep->cap) return 1;
    }
    int nid = ep->veh_nid[v];
    double acc = ep->veh_t_to_nid[v];
    const double T = (double)ep->T;
    for (int i = 0; i < len; i++) {
        int node, pod; double ddl;
        stop_info(ep, v, sche[i], &node, &ddl, &pod);
        acc += TT(c, nid, node);
        if (i >= idx_p && T + acc > ddl) {
            if (sche[i] >= 0 && (sche[i] >> 1) == new_req) return pod == -1 ? 2 : 3;
            if (i < idx_d) return 4;
            return 0;
        }
        nid = node;
    }
    *cost = acc;
    return -1;
}

static double sche_cost(octx* c, int v, const int* sche, int len) { /* scheduling.py:101-107 */
    const amod_epoch* ep = c->ep;
    int nid = ep->veh_nid[v];
    double acc = ep->veh_t_to_nid[v];
    for (int i = 0; i < len; i++) {
        int node, pod; double ddl;
        stop_info(ep, v, sche[i], &node, &ddl, &pod);
        acc += TT(c, nid, node);
        nid = node;
    }
    return acc;
}

static double sche_delay(octx* c, int v, const int* sche, int len) { /* scheduling.py:126-137 */
    const amod_epoch* ep = c->ep;
    int nid = ep->veh_nid[v];
    double acc = ep->veh_t_to_nid[v];
    double delay = 0.0;
    const double T = (double)ep->T;
    for (int i = 0; i < len; i++) {
        int node, pod; double ddl;
        stop_info(ep, v, sche[i], &node, &ddl, &pod);
        acc += TT(c, nid, node);
        if (pod == -1) { int r = sche[i] >> 1; delay += (T + acc - ep->req_tr[r] - ep->req_ts[r]); }
        nid = node;
    }
    return delay;
}

/* growable array of fixed-length schedules */
typedef struct { int* d; int len; int n; int cap; } schedlist;
static void sl_init(schedlist* s, int len) { s->d = NULL; s->len = len; s->n = 0; s->cap = 0; }
static void sl_push(schedlist* s, const int* sche) {
    if (s->n == s->cap) { s->cap = s->cap ? s->cap * 2 : 8; s->d = (int*)realloc(s->d, sizeof(int) * (size_t)s->cap * (s->len ? s->len : 1)); }
    memcpy(s->d + (size_t)s->n * s->len, sche, sizeof(int) * s->len);
    s->n++;
}
static void sl_free(schedlist* s) { free(s->d); s->d = NULL; s->n = s->cap = 0; }

/* scheduling.py:13-45.  out = feasible_sches in the reference's list order:
 * running-minimum records newest-first (insert(0)), then the others in encounter order. */
static int compute_schedule(octx* c, int v, const schedlist* subs, int q, schedlist* out, double* min_cost) {
    const amod_epoch* ep = c->ep;
    int l = subs->len;
    schedlist rec, non;
    sl_init(&rec, l + 2); sl_init(&non, l + 2); sl_init(out, l + 2);
    double best = INFINITY;
    int viol = -1;
    int tmp[MAXS];
    c->n_tasks++;
    for (int s = 0; s < subs->n; s++) {
        const int* sub = subs->d + (size_t)s * l;
        for (int i = 0; i <= l; i++) {
            for (int j = i + 1; j <= l + 1; j++) {
                /* insert_req_into_sche, scheduling.py:56-57: pick-up at i, then drop-off at j */
                int k = 0;
                for (int m = 0; m < i; m++) tmp[k++] = sub[m];
                tmp[k++] = 2 * q;
                for (int m = i; m < l; m++) tmp[k++] = sub[m];
                memmove(tmp + j + 1, tmp + j, sizeof(int) * (size_t)(l + 1 - j));
                tmp[j] = 2 * q + 1;
                double cost;
                viol = test_constraints(c, v, tmp, l + 2, q, i, j, &cost);
                if (viol == -1) {
                    if (cost < best) { best = cost; sl_push(&rec, tmp); }
                    else sl_push(&non, tmp);
                }
                if (viol > 0) break;
            }
            if (viol == 3) break;
        }
    }
    (void)ep;
    for (int r = rec.n - 1; r >= 0; r--) sl_push(out, rec.d + (size_t)r * (l + 2));
    for (int r = 0; r < non.n; r++) sl_push(out, non.d + (size_t)r * (l + 2));
    sl_free(&rec); sl_free(&non);
    *min_cost = best;
    c->n_sches += out->n;
    return out->n;
}

/* dispatch_osp.py:295-322 */
static void basic_sches(octx* c, int v, schedlist* out) {
    const amod_epoch* ep = c->ep;
    int s0 = ep->veh_sche_off[v], s1 = ep->veh_sche_off[v + 1];
    int st = ep->veh_status[v];
    if (st == 3 || st == 1 || ep->mode != AMOD_MODE_OSP) {
        sl_init(out, s1 - s0);
        sl_push(out, ep->sche_code + s0);
        return;
    }
    int base[MAXS], L = 0;
    for (int k = s0; k < s1; k++) if (ep->sche_onboard[k]) base[L++] = ep->sche_code[k];
    sl_init(out, L);
    sl_push(out, base);
    /* itertools.permutations order = lexicographic in positions; identity skipped (:317) */
    int idx[MAXS];
    for (int i = 0; i < L; i++) idx[i] = i;
    for (;;) {
        int i = L - 2;
        while (i >= 0 && idx[i] > idx[i + 1]) i--;
        if (i < 0) break;
        int j = L - 1;
        while (idx[j] < idx[i]) j--;
        int t = idx[i]; idx[i] = idx[j]; idx[j] = t;
        for (int a = i + 1, b = L - 1; a < b; a++, b--) { t = idx[a]; idx[a] = idx[b]; idx[b] = t; }
        int perm[MAXS];
        for (int m = 0; m < L; m++) perm[m] = base[idx[m]];
        double cost;
        /* test_constraints_get_cost(veh_params, sche, 0, 0, 0, T): only the flag is used */
        if (test_constraints(c, v, perm, L, -2, 0, 0, &cost) == -1) sl_push(out, perm);
    }
}

typedef struct { int reqs[AMOD_MAX_TRIP]; schedlist sches; double cost; } trip;
typedef struct { trip* t; int n, cap; } triplist;
static void tl_push(triplist* l, const trip* t) {
    if (l->n == l->cap) { l->cap = l->cap ? l->cap * 2 : 16; l->t = (trip*)realloc(l->t, sizeof(trip) * (size_t)l->cap); }
    l->t[l->n++] = *t;
}

/* open-addressing set of k-int keys (the reference uses `list in list` scans: dispatch_osp.py:194,202) */
typedef struct { const int** keys; int* vals; int size; int k; } keyset;
static uint64_t key_hash(const int* a, int k) {
    uint64_t h = 1469598103934665603ULL;
    for (int i = 0; i < k; i++) { h ^= (uint64_t)(uint32_t)a[i]; h *= 1099511628211ULL; h ^= h >> 29; }
    return h;
}
static void ks_init(keyset* s, int n, int k) {
    int sz = 16; while (sz < 2 * n + 2) sz <<= 1;
    s->size = sz; s->k = k;
    s->keys = (const int**)calloc((size_t)sz, sizeof(int*));
    s->vals = (int*)malloc(sizeof(int) * (size_t)sz);
}
static int ks_find(const keyset* s, const int* a) {
    uint64_t h = key_hash(a, s->k) & (uint64_t)(s->size - 1);
    while (s->keys[h]) { if (!memcmp(s->keys[h], a, sizeof(int) * s->k)) return s->vals[h]; h = (h + 1) & (uint64_t)(s->size - 1); }
    return -1;
}
static void ks_add(keyset* s, const int* a, int val) {
    uint64_t h = key_hash(a, s->k) & (uint64_t)(s->size - 1);
    while (s->keys[h]) h = (h + 1) & (uint64_t)(s->size - 1);
    s->keys[h] = a; s->vals[h] = val;
}
static void ks_free(keyset* s) { free((void*)s->keys); free(s->vals); }

/* per-vehicle pool slice */
typedef struct {
    int n; int cap;
    int* kind; int* nfeas; double* cost; double* delay; int64_t* toff; int64_t* soff;
    int* treq; int64_t nt, ct; int* scode; int64_t ns, cs;
    int error;
} vpool;
static void vp_entry(vpool* p, int kind, const int* reqs, int k, const int* sche, int len, double cost, double delay, int nfeas) {
    if (p->n == p->cap) {
        p->cap = p->cap ? p->cap * 2 : 16;
        p->kind = (int*)realloc(p->kind, sizeof(int) * (size_t)p->cap);
        p->nfeas = (int*)realloc(p->nfeas, sizeof(int) * (size_t)p->cap);
        p->cost = (double*)realloc(p->cost, sizeof(double) * (size_t)p->cap);
        p->delay = (double*)realloc(p->delay, sizeof(double) * (size_t)p->cap);
        p->toff = (int64_t*)realloc(p->toff, sizeof(int64_t) * (size_t)(p->cap + 1));
        p->soff = (int64_t*)realloc(p->soff, sizeof(int64_t) * (size_t)(p->cap + 1));
    }
    if (p->nt + k > p->ct) { p->ct = (p->nt + k) * 2 + 16; p->treq = (int*)realloc(p->treq, sizeof(int) * (size_t)p->ct); }
    if (p->ns + len > p->cs) { p->cs = (p->ns + len) * 2 + 16; p->scode = (int*)realloc(p->scode, sizeof(int) * (size_t)p->cs); }
    if (p->n == 0) { p->toff[0] = 0; p->soff[0] = 0; }
    if (k) memcpy(p->treq + p->nt, reqs, sizeof(int) * (size_t)k);
    p->nt += k;
    if (len) memcpy(p->scode + p->ns, sche, sizeof(int) * (size_t)len);
    p->ns += len;
    p->kind[p->n] = kind; p->nfeas[p->n] = nfeas; p->cost[p->n] = cost; p->delay[p->n] = delay;
    p->n++;
    p->toff[p->n] = p->nt; p->soff[p->n] = p->ns;
}
static void vp_free(vpool* p) { free(p->kind); free(p->nfeas); free(p->cost); free(p->delay); free(p->toff); free(p->soff); free(p->treq); free(p->scode); }

static int cmp_int(const void* a, const void* b) { int x = *(const int*)a, y = *(const int*)b; return (x > y) - (x < y); }

/* dispatch_osp.py:107-125 for one vehicle (calls :132-228). */
static void build_trip_table(octx* c, int v, vpool* out) {
    const amod_epoch* ep = c->ep;
    const int nlev = ep->cap + 4;                 /* FEASIBLE_TRIP_TABLE levels, dispatch_osp.py:8,292 */
    schedlist basic;
    basic_sches(c, v, &basic);
    /* :148 basic pair [veh, [], basic_sches[0], compute_sche_cost, 0.0] */
    vp_entry(out, AMOD_KIND_BASIC, NULL, 0, basic.d, basic.len, sche_cost(c, v, basic.d, basic.len),
             sche_delay(c, v, basic.d, basic.len), basic.n);

    triplist* lev = (triplist*)calloc((size_t)nlev + 1, sizeof(triplist));
    const double T = (double)ep->T;
    /* :156-164 trips of size 1 */
    for (int s = 0; s < ep->n_search; s++) {
        int q = ep->search[s];
        c->n_screens++;
        if (TT(c, ep->veh_nid[v], ep->req_onid[q]) + ep->veh_t_to_nid[v] + T > ep->req_clp[q]) continue;
        trip t; memset(&t, 0, sizeof(t));
        t.reqs[0] = q;
        if (compute_schedule(c, v, &basic, q, &t.sches, &t.cost) > 0) tl_push(&lev[0], &t);
        else sl_free(&t.sches);
    }
    /* :167-226 trips of size k >= 2 */
    int k = 2;
    while (lev[k - 2].n != 0) {
        if (k - 1 >= nlev) { out->error = AMOD_E_LEVELS; break; }  /* FEASIBLE_TRIP_TABLE[veh.id][k-1] -> IndexError (:226) */
        triplist* prev = &lev[k - 2];
        triplist* cur = &lev[k - 1];
        int n = prev->n;
        keyset feas, searched;
        ks_init(&feas, n, k - 1);
        for (int i = 0; i < n; i++) ks_add(&feas, prev->t[i].reqs, i);
        int scap = 64, sn = 0;
        int* skeys = (int*)malloc(sizeof(int) * (size_t)scap * k);
        ks_init(&searched, scap, k);
        for (int i = 0; i < n - 1; i++) {
            const int* a = prev->t[i].reqs;
            for (int j = i + 1; j < n; j++) {
                const int* b = prev->t[j].reqs;
                /* :187 sorted(set(trip1) | set(trip2)) */
                int u[2 * AMOD_MAX_TRIP], nu = 0, x = 0, y = 0;
                while (x < k - 1 || y < k - 1) {
                    if (y >= k - 1 || (x < k - 1 && a[x] < b[y])) u[nu++] = a[x++];
                    else if (x >= k - 1 || b[y] < a[x]) u[nu++] = b[y++];
                    else { u[nu++] = a[x]; x++; y++; }
                }
                if (k > 2) {
                    if (nu != k) continue;                                   /* :191 */
                    if (ks_find(&searched, u) >= 0) continue;                /* :194 */
                    int ok = 1;
                    for (int d = 0; d < k && ok; d++) {                      /* :198-204 */
                        int sub[AMOD_MAX_TRIP], m = 0;
                        for (int e = 0; e < k; e++) if (e != d) sub[m++] = u[e];
                        if (ks_find(&feas, sub) < 0) ok = 0;
                    }
                    if (!ok) continue;
                    if (sn == scap) {   /* keys are held by pointer: rebuild the set when the store moves */
                        scap *= 2;
                        skeys = (int*)realloc(skeys, sizeof(int) * (size_t)scap * k);
                        ks_free(&searched); ks_init(&searched, scap, k);
                        for (int e = 0; e < sn; e++) ks_add(&searched, skeys + (size_t)e * k, e);
                    }
                    memcpy(skeys + (size_t)sn * k, u, sizeof(int) * (size_t)k);
                    ks_add(&searched, skeys + (size_t)sn * k, sn);
                    sn++;
                }
                /* :210-216 insert the one request of the union missing from trip1 */
                int q = -1;
                for (int e = 0, xx = 0; e < k; e++) { if (xx < k - 1 && a[xx] == u[e]) xx++; else { q = u[e]; } }
                trip t; memset(&t, 0, sizeof(t));
                memcpy(t.reqs, u, sizeof(int) * (size_t)k);
                if (compute_schedule(c, v, &prev->t[i].sches, q, &t.sches, &t.cost) > 0) tl_push(cur, &t);
                else sl_free(&t.sches);
            }
        }
        ks_free(&feas); ks_free(&searched); free(skeys);
        k++;
    }
    /* :115-117 all searched trips, level by level */
    for (int L = 0; L < nlev + 1; L++)
        for (int i = 0; i < lev[L].n; i++) {
            trip* t = &lev[L].t[i];
            vp_entry(out, AMOD_KIND_TRIP, t->reqs, L + 1, t->sches.d, t->sches.len, t->cost,
                     sche_delay(c, v, t->sches.d, t->sches.len), t->sches.n);
        }
    /* :120-125 current working schedule */
    if (ep->mode == AMOD_MODE_OSP) {
        int s0 = ep->veh_sche_off[v], s1 = ep->veh_sche_off[v + 1];
        int reqs[MAXS], nr = 0;
        for (int m = s0; m < s1; m++) if (ep->sche_code[m] >= 0 && !(ep->sche_code[m] & 1)) reqs[nr++] = ep->sche_code[m] >> 1;
        qsort(reqs, (size_t)nr, sizeof(int), cmp_int);
        vp_entry(out, AMOD_KIND_CURRENT, reqs, nr, ep->sche_code + s0, s1 - s0, sche_cost(c, v, ep->sche_code + s0, s1 - s0),
                 sche_delay(c, v, ep->sche_code + s0, s1 - s0), 1);
    }
    for (int L = 0; L < nlev + 1; L++) { for (int i = 0; i < lev[L].n; i++) sl_free(&lev[L].t[i].sches); free(lev[L].t); }
    free(lev);
    sl_free(&basic);
}

/* ---- result object handed to Python (ctypes) ---- */
typedef struct oracle_result {
    amod_pool pool;
    amod_stats stats;
    int error;
} oracle_result;

static oracle_result* assemble(vpool* vp, const int* owner, int n, int order_is_owner) {
    oracle_result* r = (oracle_result*)calloc(1, sizeof(oracle_result));
    int64_t P = 0, NT = 0, NS = 0;
    for (int i = 0; i < n; i++) { P += vp[i].n; NT += vp[i].nt; NS += vp[i].ns; if (vp[i].error && !r->error) r->error = vp[i].error; }
    amod_pool* p = &r->pool;
    p->n_entries = P; p->n_trip_reqs = NT; p->n_stops = NS;
    p->ent_veh = (int32_t*)malloc(sizeof(int32_t) * (size_t)(P + 1));
    p->ent_kind = (int32_t*)malloc(sizeof(int32_t) * (size_t)(P + 1));
    p->ent_nfeas = (int32_t*)malloc(sizeof(int32_t) * (size_t)(P + 1));
    p->ent_cost = (double*)malloc(sizeof(double) * (size_t)(P + 1));
    p->ent_delay = (double*)malloc(sizeof(double) * (size_t)(P + 1));
    p->ent_trip_off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(P + 1));
    p->ent_sche_off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(P + 1));
    p->trip_req = (int32_t*)malloc(sizeof(int32_t) * (size_t)(NT + 1));
    p->sche_code = (int32_t*)malloc(sizeof(int32_t) * (size_t)(NS + 1));
    int64_t e = 0, t = 0, s = 0;
    p->ent_trip_off[0] = 0; p->ent_sche_off[0] = 0;
    for (int i = 0; i < n; i++) {
        vpool* q = &vp[i];
        for (int k = 0; k < q->n; k++) {
            p->ent_veh[e] = order_is_owner ? i : owner[i];
            p->ent_kind[e] = q->kind[k]; p->ent_nfeas[e] = q->nfeas[k];
            p->ent_cost[e] = q->cost[k]; p->ent_delay[e] = q->delay[k];
            e++;
            p->ent_trip_off[e] = t + q->toff[k + 1];
            p->ent_sche_off[e] = s + q->soff[k + 1];
        }
        if (q->nt) memcpy(p->trip_req + t, q->treq, sizeof(int) * (size_t)q->nt);
        if (q->ns) memcpy(p->sche_code + s, q->scode, sizeof(int) * (size_t)q->ns);
        t += q->nt; s += q->ns;
    }
    return r;
}

void oracle_free(oracle_result* r) {
    if (!r) return;
    amod_pool* p = &r->pool;
    free(p->ent_veh); free(p->ent_kind); free(p->ent_nfeas); free(p->ent_cost); free(p->ent_delay);
    free(p->ent_trip_off); free(p->ent_sche_off); free(p->trip_req); free(p->sche_code);
    free(r);
}

static void add_stats(amod_stats* s, const octx* c) {
    s->n_screens += c->n_screens; s->n_evals += c->n_evals; s->n_lookups += c->n_lookups;
    s->n_tasks += c->n_tasks; s->n_sches_stored += c->n_sches;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* B1: dispatch_osp.py:94-129.  veh_lo/veh_hi/veh_step select a vehicle subset (sampling, sharding). */
oracle_result* oracle_osp_build(const amod_epoch* ep, const void* tt, int tt_dtype, int n_nodes, int n_threads,
                                int veh_lo, int veh_hi, int veh_step) {
    if (veh_hi > ep->n_veh) veh_hi = ep->n_veh;
    if (veh_step < 1) veh_step = 1;
    int n = veh_hi > veh_lo ? (veh_hi - veh_lo + veh_step - 1) / veh_step : 0;
    vpool* vp = (vpool*)calloc((size_t)n + 1, sizeof(vpool));
    int* owner = (int*)malloc(sizeof(int) * (size_t)(n + 1));
    amod_stats st; memset(&st, 0, sizeof(st));
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel num_threads(n_threads)
    {
        octx c; memset(&c, 0, sizeof(c));
        c.ep = ep; c.n_nodes = n_nodes;
        if (tt_dtype == AMOD_TT_F64) c.tt64 = (const double*)tt; else c.tt32 = (const float*)tt;
#pragma omp for schedule(dynamic, 1)
        for (int i = 0; i < n; i++) {
            owner[i] = veh_lo + i * veh_step;
            build_trip_table(&c, owner[i], &vp[i]);
        }
#pragma omp critical
        add_stats(&st, &c);
    }
    oracle_result* r = assemble(vp, owner, n, 0);
    for (int i = 0; i < n; i++) vp_free(&vp[i]);
    free(vp); free(owner);
    r->stats = st; r->stats.n_entries = r->pool.n_entries;
    return r;
}

/* B2: dispatch_sba.py:44-74.  Request-major, then one empty pair per vehicle. */
oracle_result* oracle_sba_build(const amod_epoch* ep, const void* tt, int tt_dtype, int n_nodes, int n_threads) {
    int nq = ep->n_search, nv = ep->n_veh;
    vpool* vp = (vpool*)calloc((size_t)nq + 2, sizeof(vpool));
    int* owner = (int*)calloc((size_t)nq + 2, sizeof(int));
    amod_stats st; memset(&st, 0, sizeof(st));
    const double T = (double)ep->T;
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel num_threads(n_threads)
    {
        octx c; memset(&c, 0, sizeof(c));
        c.ep = ep; c.n_nodes = n_nodes;
        if (tt_dtype == AMOD_TT_F64) c.tt64 = (const double*)tt; else c.tt32 = (const float*)tt;
#pragma omp for schedule(dynamic, 1)
        for (int s = 0; s <= nq; s++) {
            if (s == nq) {   /* :68-69 */
                for (int v = 0; v < nv; v++) {
                    int s0 = ep->veh_sche_off[v], s1 = ep->veh_sche_off[v + 1];
                    vp_entry(&vp[s], AMOD_KIND_SBA_EMPTY, NULL, 0, ep->sche_code + s0, s1 - s0,
                             sche_cost(&c, v, ep->sche_code + s0, s1 - s0), sche_delay(&c, v, ep->sche_code + s0, s1 - s0), 1);
                    vp[s].kind[vp[s].n - 1] = AMOD_KIND_SBA_EMPTY;
                }
                continue;
            }
            int q = ep->search[s];
            for (int v = 0; v < nv; v++) {   /* :58-65 */
                c.n_screens++;
                if (TT(&c, ep->veh_nid[v], ep->req_onid[q]) + ep->veh_t_to_nid[v] + T > ep->req_clp[q]) continue;
                int s0 = ep->veh_sche_off[v], s1 = ep->veh_sche_off[v + 1];
                schedlist sub, out; double cost;
                sl_init(&sub, s1 - s0); sl_push(&sub, ep->sche_code + s0);
                if (compute_schedule(&c, v, &sub, q, &out, &cost) > 0) {
                    /* the owning vehicle is stored in nfeas' sibling array below */
                    vp_entry(&vp[s], v, &q, 1, out.d, out.len, cost, sche_delay(&c, v, out.d, out.len), out.n);
                }
                sl_free(&sub); sl_free(&out);
            }
        }
#pragma omp critical
        add_stats(&st, &c);
    }
    /* vp[s].kind temporarily carries the vehicle index for pair entries */
    oracle_result* r = assemble(vp, owner, nq + 1, 1);
    int64_t e = 0;
    for (int s = 0; s <= nq; s++)
        for (int k = 0; k < vp[s].n; k++, e++) {
            if (s < nq) { r->pool.ent_veh[e] = vp[s].kind[k]; r->pool.ent_kind[e] = AMOD_KIND_SBA_PAIR; }
            else { r->pool.ent_veh[e] = k; r->pool.ent_kind[e] = AMOD_KIND_SBA_EMPTY; }
        }
    for (int s = 0; s <= nq; s++) vp_free(&vp[s]);
    free(vp); free(owner);
    r->stats = st; r->stats.n_entries = r->pool.n_entries;
    return r;
}

/* B3: rebalancing_npo.py:26-59.  Candidates request-major; stable sort by rebl_dt; greedy disjoint. */
typedef struct { double dt; int64_t seq; } npo_edge;
static int cmp_edge(const void* a, const void* b) {
    const npo_edge* x = (const npo_edge*)a; const npo_edge* y = (const npo_edge*)b;
    if (x->dt < y->dt) return -1;
    if (x->dt > y->dt) return 1;
    return (x->seq > y->seq) - (x->seq < y->seq);
}
int oracle_npo_match(const void* tt, int tt_dtype, int n_nodes, const int32_t* idle_nid, int n_idle,
                     const int32_t* pend_onid, int n_pend, int32_t* out_idle, int32_t* out_pend, double* out_dt) {
    int64_t n = (int64_t)n_idle * n_pend;
    if (n == 0) return 0;
    npo_edge* e = (npo_edge*)malloc(sizeof(npo_edge) * (size_t)n);
    for (int r = 0; r < n_pend; r++)
        for (int v = 0; v < n_idle; v++) {
            int64_t k = (int64_t)(idle_nid[v] - 1) * n_nodes + (pend_onid[r] - 1);
            e[(int64_t)r * n_idle + v].dt = tt_dtype == AMOD_TT_F64 ? ((const double*)tt)[k] : (double)((const float*)tt)[k];
            e[(int64_t)r * n_idle + v].seq = (int64_t)r * n_idle + v;
        }
    qsort(e, (size_t)n, sizeof(npo_edge), cmp_edge);   /* list.sort(key=-score) is stable (:48) */
    char* vu = (char*)calloc((size_t)n_idle, 1);
    char* ru = (char*)calloc((size_t)n_pend, 1);
    int m = 0;
    for (int64_t i = 0; i < n; i++) {
        int r = (int)(e[i].seq / n_idle), v = (int)(e[i].seq % n_idle);
        if (vu[v] || ru[r]) continue;
        vu[v] = 1; ru[r] = 1;
        out_idle[m] = v; out_pend[m] = r; out_dt[m] = e[i].dt; m++;
    }
    free(e); free(vu); free(ru);
    return m;
}

/* scheduling.py:70-98 exposed for the unit golden vectors. */
int oracle_test_constraints(const amod_epoch* ep, const void* tt, int tt_dtype, int n_nodes, int v, const int* sche,
                            int len, int new_req, int idx_p, int idx_d, double* cost) {
    octx c; memset(&c, 0, sizeof(c));
    c.ep = ep; c.n_nodes = n_nodes;
    if (tt_dtype == AMOD_TT_F64) c.tt64 = (const double*)tt; else c.tt32 = (const float*)tt;
    *cost = INFINITY;
    return test_constraints(&c, v, sche, len, new_req, idx_p, idx_d, cost);
}

/* ---- next row: greedy assignment over the pool, ilp_assign.py:101-130 --------------
 * list.sort(key=-score) is stable; then one pass keeps a pair iff its vehicle and all its requests are
 * still free.  order_out = pool indices in sorted order; sel_out = positions IN THE SORTED LIST (that is
 * what the reference returns, because it sorts the caller's list in place, ilp_assign.py:109). */
typedef struct { double neg_score; int64_t idx; } ga_key;
static int cmp_ga(const void* a, const void* b) {
    const ga_key* x = (const ga_key*)a; const ga_key* y = (const ga_key*)b;
    if (x->neg_score < y->neg_score) return -1;
    if (x->neg_score > y->neg_score) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}
int oracle_greedy_assign(const amod_pool* p, const double* score, int n_veh, int n_req, int32_t* order_out, int32_t* sel_out) {
    int64_t P = p->n_entries;
    if (P == 0) return 0;
    ga_key* k = (ga_key*)malloc(sizeof(ga_key) * (size_t)P);
    for (int64_t e = 0; e < P; e++) { k[e].neg_score = -score[e]; k[e].idx = e; }
    qsort(k, (size_t)P, sizeof(ga_key), cmp_ga);
    char* vu = (char*)calloc((size_t)n_veh + 1, 1);
    char* ru = (char*)calloc((size_t)n_req + 1, 1);
    int m = 0;
    for (int64_t i = 0; i < P; i++) {
        int64_t e = k[i].idx;
        order_out[i] = (int32_t)e;
        if (vu[p->ent_veh[e]]) continue;                                   /* :116-117 */
        int clash = 0;
        for (int64_t t = p->ent_trip_off[e]; t < p->ent_trip_off[e + 1]; t++) clash |= ru[p->trip_req[t]];   /* :119-120 */
        if (clash) continue;
        vu[p->ent_veh[e]] = 1;
        for (int64_t t = p->ent_trip_off[e]; t < p->ent_trip_off[e + 1]; t++) ru[p->trip_req[t]] = 1;
        sel_out[m++] = (int32_t)i;
    }
    free(k); free(vu); free(ru);
    return m;
}

/* ---- route_functions.build_route_from_origin_to_dest (route_functions.py:36-70), restated -----------
 * Steps of the best path o -> d: (t, d, [u, v]) per consecutive node pair of the path recovered by walking the
 * predecessor table back from d (:38-43), then the closing (0, 0, [tnid, tnid]) step (:59-61); *leg_t / *leg_d are the
 * left-to-right sums (:56-57).  Returns the number of steps, or -(steps needed) if cap is too small, 0 on a broken table. */
int64_t oracle_route_build(const int32_t* pred, const void* tt, const void* dist, int dtype, int n, int o, int d,
                           int32_t* su, int32_t* sv, double* st, double* sd, int64_t cap, double* leg_t, double* leg_d) {
    const int32_t* row = pred + (size_t)(o - 1) * (size_t)n;
    int64_t len = 1;
    for (int x = row[d - 1]; x > 0; x = row[x - 1]) if (++len > n) return 0;
    if (len > cap) return -len;
    int x = d;
    for (int64_t p = len - 1; p >= 0; p--) { sv[p] = x; if (p > 0) x = row[x - 1]; }      /* path.reverse() */
    for (int64_t k = 0; k < len; k++) su[k] = sv[k];
    double duration = 0.0, distance = 0.0;
    for (int64_t k = 0; k + 1 < len; k++) {
        int u = su[k], v = su[k + 1];
        size_t at = (size_t)(u - 1) * (size_t)n + (size_t)(v - 1);
        st[k] = dtype == AMOD_TT_F64 ? ((const double*)tt)[at] : (double)((const float*)tt)[at];
        sd[k] = dtype == AMOD_TT_F64 ? ((const double*)dist)[at] : (double)((const float*)dist)[at];
        sv[k] = v;
        duration += st[k]; distance += sd[k];
    }
    su[len - 1] = d; sv[len - 1] = d; st[len - 1] = 0.0; sd[len - 1] = 0.0;
    *leg_t = duration; *leg_d = distance;
    return len;
}
