/* TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's fleet state advance and apply step:
 *   Veh.move_to_time                              src/simulator/vehicle.py:73-162
 *   Veh.update_service_status_after_veh_finish_a_leg, pop_leg, pop_step, cut_step, jump_to_location   vehicle.py:164-226
 *   Veh.build_route, add_leg, clear_route          vehicle.py:231-310
 *   Platform.upd_vehs_and_reqs_stat_to_time        src/simulator/platform.py:213-233  (pick/drop info request.py:43-57)
 * It follows the reference's statements literally (same order of float64 operations, the deque of legs and steps kept as
 * arrays with a head index).  Parity status: PINNED by tests/golden/trace_fleet_{sba,osp}.npz, recorded from the unmodified
 * reference simulator by tests/golden/make_trace_fleet.py (every vehicle / request field after every epoch's advance).
 * Only tests/ may load this; the product (amod_b200/) never does. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { ST_IDLE = 1, ST_WORKING = 2, ST_REBALANCING = 3 };                      /* types.py:18-21 */
enum { RQ_PENDING = 1, RQ_PICKING = 2, RQ_ONBOARD = 3, RQ_COMPLETE = 4, RQ_WALKAWAY = 5 };   /* types.py:11-16 */

typedef struct { double t, d; int u, v; double g0[2], g1[2]; } ostep;         /* types.py:44-59 Step */
typedef struct { int rid, pod, tnid; double ddl, t, d; ostep* steps; int head, n, cap; } oleg;    /* types.py:62-84 Leg */
typedef struct { int rid, pod, tnid; double ddl; } ostop;

typedef struct {
    int id, status, nid, tnid, load;
    double state_time, lng, lat, t_to_nid, d_to_nid, t, d;
    double Ds, Ts, Ds_empty, Ts_empty, Dr, Tr, Lt, Ld;
    int has_step; ostep step_to_nid;
    ostop* sche; int sche_n, sche_cap;
    oleg* route; int route_head, route_n, route_cap;
    int* onboard; int n_onboard, onboard_cap;
    int* picking; int n_picking, picking_cap;
} oveh;

typedef struct { int32_t veh, rid, pod; double time; } oevent;

typedef struct ofleet {
    int n_veh, n_nodes, dtype;
    const void* tt; const void* dist; const int32_t* pred; const double* lng; const double* lat;
    oveh* vehs;
    /* requests (request.py:25-41) */
    int n_req, req_cap;
    int32_t* r_status; double *r_tr, *r_clp, *r_ts, *r_tp, *r_td;
    oevent* ev; int n_ev, ev_cap;
    int error;
} ofleet;

static double tab(const ofleet* f, const void* p, int o, int d) {
    size_t at = (size_t)(o - 1) * (size_t)f->n_nodes + (size_t)(d - 1);
    return f->dtype == 1 ? ((const double*)p)[at] : (double)((const float*)p)[at];
}

ofleet* ofleet_create(int n_veh, int n_nodes, int dtype, const void* tt, const void* dist, const int32_t* pred, const double* lng,
                      const double* lat, const int32_t* start_nid) {
    ofleet* f = (ofleet*)calloc(1, sizeof *f);
    f->n_veh = n_veh; f->n_nodes = n_nodes; f->dtype = dtype; f->tt = tt; f->dist = dist; f->pred = pred; f->lng = lng; f->lat = lat;
    f->vehs = (oveh*)calloc((size_t)(n_veh > 0 ? n_veh : 1), sizeof(oveh));
    for (int i = 0; i < n_veh; i++) {                /* vehicle.py:41-70 */
        oveh* v = &f->vehs[i];
        v->id = i; v->status = ST_IDLE; v->state_time = 0.0; v->nid = start_nid[i]; v->tnid = v->nid;
        v->lng = lng[v->nid - 1]; v->lat = lat[v->nid - 1];
    }
    return f;
}

static void leg_free(oleg* l) { free(l->steps); l->steps = NULL; }

void ofleet_destroy(ofleet* f) {
    if (!f) return;
    for (int i = 0; i < f->n_veh; i++) {
        oveh* v = &f->vehs[i];
        for (int k = 0; k < v->route_n; k++) leg_free(&v->route[k]);
        free(v->route); free(v->sche); free(v->onboard); free(v->picking);
    }
    free(f->vehs); free(f->r_status); free(f->r_tr); free(f->r_clp); free(f->r_ts); free(f->r_tp); free(f->r_td); free(f->ev);
    free(f);
}

void ofleet_add_requests(ofleet* f, int n, const double* tr, const double* clp, const double* ts) {
    if (f->n_req + n > f->req_cap) {
        int c = (f->n_req + n) * 2 + 64;
        f->r_status = (int32_t*)realloc(f->r_status, 4 * (size_t)c);
        f->r_tr = (double*)realloc(f->r_tr, 8 * (size_t)c); f->r_clp = (double*)realloc(f->r_clp, 8 * (size_t)c); f->r_ts = (double*)realloc(f->r_ts, 8 * (size_t)c);
        f->r_tp = (double*)realloc(f->r_tp, 8 * (size_t)c); f->r_td = (double*)realloc(f->r_td, 8 * (size_t)c);
        f->req_cap = c;
    }
    for (int i = 0; i < n; i++) {                    /* request.py:25-41 */
        int r = f->n_req + i;
        f->r_status[r] = RQ_PENDING; f->r_tr[r] = tr[i]; f->r_clp[r] = clp[i]; f->r_ts[r] = ts[i]; f->r_tp[r] = -1.0; f->r_td[r] = -1.0;
    }
    f->n_req += n;
}

void ofleet_set_status(ofleet* f, const int32_t* rids, int n, int status) { for (int i = 0; i < n; i++) f->r_status[rids[i]] = status; }

/* ---- vehicle.py:219-224 */
static void jump_to_node(ofleet* f, oveh* v, int nid) { v->nid = nid; v->lng = f->lng[nid - 1]; v->lat = f->lat[nid - 1]; }
static void jump_to(oveh* v, int nid, double lng, double lat) { v->nid = nid; v->lng = lng; v->lat = lat; }

static void list_remove(int* a, int* n, int x) {
    for (int i = 0; i < *n; i++) if (a[i] == x) { memmove(a + i, a + i + 1, sizeof(int) * (size_t)(*n - i - 1)); (*n)--; return; }
}
static void list_append(int** a, int* n, int* cap, int x) {
    if (*n == *cap) { *cap = *cap * 2 + 8; *a = (int*)realloc(*a, sizeof(int) * (size_t)*cap); }
    (*a)[(*n)++] = x;
}

static void event(ofleet* f, int veh, int rid, int pod, double time) {
    if (f->n_ev == f->ev_cap) { f->ev_cap = f->ev_cap * 2 + 64; f->ev = (oevent*)realloc(f->ev, sizeof(oevent) * (size_t)f->ev_cap); }
    oevent e = {veh, rid, pod, time};
    f->ev[f->n_ev++] = e;
}

/* vehicle.py:188-191 pop_leg */
static void pop_leg(oveh* v) {
    oleg* leg = &v->route[v->route_head];
    v->d -= leg->d; v->t -= leg->t;
    leg_free(leg);
    v->route_head++;
}

/* vehicle.py:164-186 update_service_status_after_veh_finish_a_leg */
static void finish_leg(ofleet* f, oveh* v, oleg* leg) {
    jump_to_node(f, v, leg->tnid);
    v->load += leg->pod;
    if (leg->rid != -2) {
        event(f, v->id, leg->rid, leg->pod, v->state_time);             /* done.append((leg.rid, leg.pod, self.state_time)) */
        if (v->sche_n <= 0 || v->sche[0].rid != leg->rid) f->error |= 1;    /* assert finished_sche_step[0] == leg.rid */
        memmove(v->sche, v->sche + 1, sizeof(ostop) * (size_t)(v->sche_n - 1)); v->sche_n--;      /* self.sche.pop(0) */
        if (leg->pod == 1) { list_remove(v->picking, &v->n_picking, leg->rid); list_append(&v->onboard, &v->n_onboard, &v->onboard_cap, leg->rid); }
        else if (leg->pod == -1) list_remove(v->onboard, &v->n_onboard, leg->rid);
    }
    if (v->load != v->n_onboard) f->error |= 2;
    pop_leg(v);
}

static void stat_add(oveh* v, int upd, double t, double d) {               /* vehicle.py:89-98, 107-116, 131-140 */
    if (!upd) return;
    v->Ts += t; v->Ds += d; v->Lt += t * v->load; v->Ld += d * v->load;
    if (v->status == ST_WORKING && v->load == 0) { v->Ts_empty += t; v->Ds_empty += d; }
    if (v->status == ST_REBALANCING) { v->Tr += t; v->Dr += d; }
}

/* vehicle.py:73-162 move_to_time */
static void move_to_time(ofleet* f, oveh* v, double T, int upd) {
    double dT = T - v->state_time;
    if (dT <= 0) return;
    v->has_step = 0; v->t_to_nid = 0.0; v->d_to_nid = 0.0;
    while (dT > 0 && v->route_head < v->route_n) {
        oleg* leg = &v->route[v->route_head];
        if (leg->t < dT) {
            dT -= leg->t; v->state_time += leg->t;
            stat_add(v, upd, leg->t, leg->d);
            finish_leg(f, v, leg);
        } else {
            while (dT > 0 && leg->head < leg->n) {
                ostep* step = &leg->steps[leg->head];
                if (step->t < dT) {
                    dT -= step->t; v->state_time += step->t;
                    stat_add(v, upd, step->t, step->d);
                    jump_to(v, step->v, step->g1[0], step->g1[1]);
                    /* pop_step (vehicle.py:194-199) */
                    v->t -= step->t; v->d -= step->d; leg->t -= step->t; leg->d -= step->d; leg->head++;
                    if (leg->head == leg->n) { finish_leg(f, v, leg); break; }
                } else {
                    double pct = dT / step->t;
                    if (upd) {      /* (the partial-step statistics use dT and step.d * pct, vehicle.py:131-140) */
                        v->Ts += dT; v->Ds += step->d * pct; v->Lt += dT * v->load; v->Ld += step->d * pct * v->load;
                        if (v->status == ST_WORKING && v->load == 0) { v->Ts_empty += dT; v->Ds_empty += step->d * pct; }
                        if (v->status == ST_REBALANCING) { v->Tr += dT; v->Dr += step->d * pct; }
                    }
                    /* cut_step (vehicle.py:202-217) */
                    step->u = step->v;
                    step->g0[0] += pct * (step->g1[0] - step->g0[0]);
                    step->g0[1] += pct * (step->g1[1] - step->g0[1]);
                    v->t_to_nid = step->t * (1 - pct); v->d_to_nid = step->d * (1 - pct);
                    v->t -= step->t * pct; v->d -= step->d * pct;
                    leg->t -= step->t * pct; leg->d -= step->d * pct;
                    { double st = step->t, sd = step->d; step->t -= st * pct; step->d -= sd * pct; }
                    v->step_to_nid = *step; v->has_step = 1;
                    jump_to(v, step->u, step->g0[0], step->g0[1]);
                    v->state_time = T;
                    return;
                }
            }
        }
    }
    /* finished the whole schedule (vehicle.py:152-162) */
    if (v->route_head != v->route_n || v->sche_n != 0 || v->load != 0 || v->n_onboard != 0 || v->n_picking != 0) f->error |= 4;
    v->route_head = v->route_n = 0;
    v->state_time = T; v->d = 0.0; v->t = 0.0; v->status = ST_IDLE;
}

/* platform.py:213-233; returns the number of done events of this epoch (ofleet_events copies them out) */
int ofleet_advance(ofleet* f, double T, int upd) {
    f->n_ev = 0;
    for (int i = 0; i < f->n_veh; i++) {
        int first = f->n_ev;
        move_to_time(f, &f->vehs[i], T, upd);
        for (int k = first; k < f->n_ev; k++) {
            oevent* e = &f->ev[k];
            if (e->pod == 1) {                       /* request.py:43-50 update_pick_info */
                f->r_tp[e->rid] = e->time;
                if (f->r_status[e->rid] != RQ_PICKING) f->error |= 8;
                f->r_status[e->rid] = RQ_ONBOARD;
            } else if (e->pod == -1) {               /* request.py:52-57 update_drop_info */
                f->r_td[e->rid] = e->time;
                if (f->r_status[e->rid] != RQ_ONBOARD) f->error |= 16;
                f->r_status[e->rid] = RQ_COMPLETE;
            }
        }
    }
    for (int r = 0; r < f->n_req; r++) {             /* platform.py:227-232 */
        if (f->r_status[r] != RQ_PENDING) continue;
        if (T >= f->r_tr[r] + 150 || T >= f->r_clp[r]) f->r_status[r] = RQ_WALKAWAY;
    }
    return f->n_ev;
}

void ofleet_events(const ofleet* f, int32_t* veh, int32_t* rid, int32_t* pod, double* time) {
    for (int k = 0; k < f->n_ev; k++) { veh[k] = f->ev[k].veh; rid[k] = f->ev[k].rid; pod[k] = f->ev[k].pod; time[k] = f->ev[k].time; }
}

static oleg* route_push(oveh* v) {
    if (v->route_n == v->route_cap) { v->route_cap = v->route_cap * 2 + 8; v->route = (oleg*)realloc(v->route, sizeof(oleg) * (size_t)v->route_cap); }
    oleg* l = &v->route[v->route_n++];
    memset(l, 0, sizeof *l);
    return l;
}
static ostep* leg_push(oleg* l) {
    if (l->n == l->cap) { l->cap = l->cap * 2 + 16; l->steps = (ostep*)realloc(l->steps, sizeof(ostep) * (size_t)l->cap); }
    return &l->steps[l->n++];
}

/* vehicle.py:290-302 add_leg + route_functions.py:36-70 build_route_from_origin_to_dest */
static void add_leg(ofleet* f, oveh* v, int rid, int pod, int tnid, double ddl) {
    const int o = v->tnid, n = f->n_nodes;
    const int32_t* row = f->pred + (size_t)(o - 1) * (size_t)n;
    int len = 1;
    for (int x = row[tnid - 1]; x > 0; x = row[x - 1]) if (++len > n) { f->error |= 32; return; }
    int* path = (int*)malloc(sizeof(int) * (size_t)len);
    { int x = tnid; for (int p = len - 1; p >= 0; p--) { path[p] = x; if (p > 0) x = row[x - 1]; } }
    oleg* leg = route_push(v);
    leg->rid = rid; leg->pod = pod; leg->tnid = tnid; leg->ddl = ddl;
    double duration = 0.0, distance = 0.0;
    for (int i = 0; i + 1 < len; i++) {
        ostep* s = leg_push(leg);
        s->u = path[i]; s->v = path[i + 1];
        s->t = tab(f, f->tt, s->u, s->v); s->d = tab(f, f->dist, s->u, s->v);
        s->g0[0] = f->lng[s->u - 1]; s->g0[1] = f->lat[s->u - 1]; s->g1[0] = f->lng[s->v - 1]; s->g1[1] = f->lat[s->v - 1];
        duration += s->t; distance += s->d;
    }
    { ostep* s = leg_push(leg); int e = path[len - 1];
      s->t = 0.0; s->d = 0.0; s->u = e; s->v = e; s->g0[0] = s->g1[0] = f->lng[e - 1]; s->g0[1] = s->g1[1] = f->lat[e - 1]; }
    leg->t = duration; leg->d = distance;
    v->tnid = path[len - 1];
    v->d += leg->d; v->t += leg->t;
    free(path);
}

/* vehicle.py:231-288 build_route; `sche` may lose its rebalancing stop exactly as the reference mutates the caller's list */
static void build_route(ofleet* f, oveh* v, ostop* sche, int n) {
    if (v->status == ST_REBALANCING && n > 1) {
        if (v->sche_n != 1) f->error |= 64;
        for (int i = 0; i < n; i++) if (sche[i].rid == -1) { memmove(sche + i, sche + i + 1, sizeof(ostop) * (size_t)(n - i - 1)); n--; break; }
    }
    /* clear_route (vehicle.py:304-310) */
    v->n_picking = 0;
    for (int k = v->route_head; k < v->route_n; k++) leg_free(&v->route[k]);
    v->route_head = v->route_n = 0;
    v->d = 0.0; v->t = 0.0; v->tnid = v->nid;
    if (n > v->sche_cap) { v->sche_cap = n + 8; v->sche = (ostop*)realloc(v->sche, sizeof(ostop) * (size_t)v->sche_cap); }
    memcpy(v->sche, sche, sizeof(ostop) * (size_t)n); v->sche_n = n;
    if (v->has_step) {                               /* the unfinished step from the last move (vehicle.py:246-266) */
        oleg* leg = route_push(v);
        const ostep* s = &v->step_to_nid;
        leg->rid = -2; leg->pod = 0; leg->tnid = s->v; leg->ddl = v->state_time + s->t; leg->t = s->t; leg->d = s->d;
        *leg_push(leg) = *s;
        ostep* e = leg_push(leg);
        e->t = 0.0; e->d = 0.0; e->u = s->v; e->v = s->v; e->g0[0] = e->g1[0] = s->g1[0]; e->g0[1] = e->g1[1] = s->g1[1];
        v->tnid = s->v;
        v->d += leg->d; v->t += leg->t;
    }
    for (int i = 0; i < n; i++) {
        add_leg(f, v, v->sche[i].rid, v->sche[i].pod, v->sche[i].tnid, v->sche[i].ddl);
        if (v->sche[i].pod == 1) list_append(&v->picking, &v->n_picking, &v->picking_cap, v->sche[i].rid);
    }
    if (n > 0) {
        if (v->sche[0].pod == 1 || v->sche[0].pod == -1) {
            v->status = ST_WORKING;
            int m = v->load;
            for (int k = v->route_head; k < v->route_n; k++) m += v->route[k].pod;
            if (m != 0) f->error |= 128;
        } else {
            v->status = ST_REBALANCING;
            if (!(v->sche[0].rid == -1 && n == 1)) f->error |= 256;
        }
    } else v->status = ST_IDLE;
}

/* build_route for a batch of vehicles: call i gives vehicle veh[i] the stops [off[i], off[i+1]) */
void ofleet_apply(ofleet* f, int n, const int32_t* veh, const int32_t* off, const int32_t* rid, const int32_t* pod, const int32_t* tnid, const double* ddl) {
    for (int i = 0; i < n; i++) {
        int m = off[i + 1] - off[i];
        ostop* s = (ostop*)malloc(sizeof(ostop) * (size_t)(m > 0 ? m : 1));
        for (int k = 0; k < m; k++) { s[k].rid = rid[off[i] + k]; s[k].pod = pod[off[i] + k]; s[k].tnid = tnid[off[i] + k]; s[k].ddl = ddl[off[i] + k]; }
        build_route(f, &f->vehs[veh[i]], s, m);
        free(s);
    }
}

int ofleet_error(const ofleet* f) { return f->error; }
int ofleet_n_req(const ofleet* f) { return f->n_req; }

/* per-vehicle scalars, 24 doubles per vehicle: nid lng lat t_to_nid d_to_nid tnid load status state_time t d
 * Ds Ts Ds_empty Ts_empty Dr Tr Lt Ld has_step step_t step_d step_u step_v */
void ofleet_export(const ofleet* f, double* out, int32_t* sche_len, int32_t* n_legs) {
    for (int i = 0; i < f->n_veh; i++) {
        const oveh* v = &f->vehs[i];
        double* o = out + 24 * (size_t)i;
        o[0] = v->nid; o[1] = v->lng; o[2] = v->lat; o[3] = v->t_to_nid; o[4] = v->d_to_nid; o[5] = v->tnid; o[6] = v->load; o[7] = v->status;
        o[8] = v->state_time; o[9] = v->t; o[10] = v->d; o[11] = v->Ds; o[12] = v->Ts; o[13] = v->Ds_empty; o[14] = v->Ts_empty;
        o[15] = v->Dr; o[16] = v->Tr; o[17] = v->Lt; o[18] = v->Ld; o[19] = v->has_step;
        o[20] = v->has_step ? v->step_to_nid.t : 0.0; o[21] = v->has_step ? v->step_to_nid.d : 0.0;
        o[22] = v->has_step ? v->step_to_nid.u : 0; o[23] = v->has_step ? v->step_to_nid.v : 0;
        sche_len[i] = v->sche_n; n_legs[i] = v->route_n - v->route_head;
    }
}

/* schedules (CSR over vehicles, offsets from ofleet_export's sche_len) */
void ofleet_export_sche(const ofleet* f, int32_t* rid, int32_t* pod, int32_t* tnid, double* ddl) {
    size_t k = 0;
    for (int i = 0; i < f->n_veh; i++)
        for (int s = 0; s < f->vehs[i].sche_n; s++, k++) { rid[k] = f->vehs[i].sche[s].rid; pod[k] = f->vehs[i].sche[s].pod; tnid[k] = f->vehs[i].sche[s].tnid; ddl[k] = f->vehs[i].sche[s].ddl; }
}

/* legs (CSR over vehicles, offsets from ofleet_export's n_legs): rid pod tnid ddl t d remaining steps */
void ofleet_export_legs(const ofleet* f, int32_t* rid, int32_t* pod, int32_t* tnid, double* ddl, double* t, double* d, int32_t* n_steps) {
    size_t k = 0;
    for (int i = 0; i < f->n_veh; i++)
        for (int s = f->vehs[i].route_head; s < f->vehs[i].route_n; s++, k++) {
            const oleg* l = &f->vehs[i].route[s];
            rid[k] = l->rid; pod[k] = l->pod; tnid[k] = l->tnid; ddl[k] = l->ddl; t[k] = l->t; d[k] = l->d; n_steps[k] = l->n - l->head;
        }
}

void ofleet_export_requests(const ofleet* f, int32_t* status, double* tp, double* td) {
    for (int r = 0; r < f->n_req; r++) { status[r] = f->r_status[r]; tp[r] = f->r_tp[r]; td[r] = f->r_td[r]; }
}
