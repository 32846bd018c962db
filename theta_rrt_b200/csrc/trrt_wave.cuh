// trrt_wave.cuh -- schedule 2 of the fused RRT loop: the speculative window of trrt_rrt.cuh cut into three kernels
// per window over ALL queries of the batch (one warp per query):
//
//   wave_scan    sample, freespace(qrand), `qrand in G` probe, nearest scan            (tiny code, high occupancy)
//   wave_expand  steer, clearance rays, 1/3 re-drive, edge raster, `qnew in G` probe   (the ~40 KB of expansion code)
//   wave_commit  commit of the 32 iterations in order, as in rrt_kernel_spec's phase B
//
// Why: in the persistent kernel a third of the warp time is spent at the CTA barrier that keeps the expansion code
// in the instruction cache, and its occupancy is set by the fattest phase.  Here every warp of the GPU is in the same
// phase at the same time by construction (no barrier, ideal instruction reuse), and each phase gets the registers
// and occupancy it needs.  The window state travels through a record per lane in the workspace (SoA, lane fastest).
// Same arithmetic, same commit order: results are bit-identical to the other two schedules.
//
// STATUS (round 1): correct and parity-tested, but EXPERIMENTAL and not the default -- 105 ms per cfg-3 step against
// 69 ms for the persistent kernel.  Per window (ncu, mid-run): scan 141 us, expand 117 us, re-expand 56 us, commit
// 292 us.  2.9% of the iterations (0.92 per window and query) find a node of their own window nearer than their
// snapshot winner; wave_reexpand predicts them from the tentative nodes and expands them GPU-wide, which took the
// one-lane expansions out of the commit (39 M -> 26 M instructions).  What is left in the commit is a serial chain
// of dependent memory operations that walks ALL queries once per window: the hash-table probe of tree_insert alone
// is 17% of its warp time at ~3000 cycles per load -- every access lands on a page and an L2 line the SM has not
// seen since the previous window (2 GB of per-query state against the persistent kernel's few pages per warp).
// To make this schedule win, the window-hot state of a query (index slots, tail of the tree) has to be packed.
#pragma once
#include "trrt_rrt.cuh"

namespace trrt {

struct WaveDev {
    // per lane [nq][32]
    int *pre, *near, *exist, *code, *flags, *aux; // aux packs drive | lospx<<1 .. (counters only)
    double *bd, *qx, *qy, *qth, *wx, *wy, *wth, *usteer, *iccx, *iccy, *rad, *udist;
    int *lospx, *arcpx, *arcang;
    // per lane, second expansion of the lanes predicted to be re-expanded (wave_reexpand): from = lane whose tentative
    // node it starts from (-1 = none)
    int *a_from, *a_code, *a_flags, *a_exist, *a_aux, *a_lospx, *a_arcpx, *a_arcang;
    double *a_wx, *a_wy, *a_wth, *a_usteer, *a_iccx, *a_iccy, *a_rad, *a_udist;
    // per query [nq]
    int *n, *nlos, *sol, *status, *iters, *active;
    unsigned long long *cnt; // [nq][8]
};

#define TRRT_WAVE_LANE_BYTES (17 * 4 + 20 * 8)
#define TRRT_WAVE_QUERY_BYTES (6 * 4 + 8 * 8)

__host__ __device__ inline size_t wave_bytes(int64_t nq) {
    return (size_t)nq * 32 * TRRT_WAVE_LANE_BYTES + (size_t)nq * TRRT_WAVE_QUERY_BYTES + 256;
}
inline WaveDev wave_carve(void *base, int64_t nq) {
    WaveDev w;
    char *p = (char *)base;
    const size_t L = (size_t)nq * 32;
    double **dbl[] = {&w.bd, &w.qx, &w.qy, &w.qth, &w.wx, &w.wy, &w.wth, &w.usteer, &w.iccx, &w.iccy, &w.rad, &w.udist,
                      &w.a_wx, &w.a_wy, &w.a_wth, &w.a_usteer, &w.a_iccx, &w.a_iccy, &w.a_rad, &w.a_udist};
    for (double **d : dbl) { *d = (double *)p; p += L * 8; }
    w.cnt = (unsigned long long *)p; p += (size_t)nq * 64;
    int **il[] = {&w.pre, &w.near, &w.exist, &w.code, &w.flags, &w.aux, &w.lospx, &w.arcpx, &w.arcang,
                  &w.a_from, &w.a_code, &w.a_flags, &w.a_exist, &w.a_aux, &w.a_lospx, &w.a_arcpx, &w.a_arcang};
    for (int **d : il) { *d = (int *)p; p += L * 4; }
    int **iq[] = {&w.n, &w.nlos, &w.sol, &w.status, &w.iters, &w.active};
    for (int **d : iq) { *d = (int *)p; p += (size_t)nq * 4; }
    return w;
}

#define TRRT_WAVE_THREADS 256 /* scan, init */
#ifndef TRRT_WAVE_SMALL
#define TRRT_WAVE_SMALL 64   /* expand, commit: warps differ a lot in duration, small CTAs keep the SMs filled */
#endif

// one warp per query: empty index, start node, per-query state
__global__ void __launch_bounds__(TRRT_WAVE_THREADS) wave_init(const RrtDev a, const WaveDev w) {
    const Group<32> g;
    const int64_t q = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (q >= a.nq) return;
    RrtQuery Q;
    rrt_setup<32>(a, q, g, Q);
    if (g.gl == 0) {
        w.n[q] = 1; w.nlos[q] = 0; w.sol[q] = -1; w.status[q] = TRRT_OK_NOT_FOUND; w.iters[q] = 0; w.active[q] = (a.K > 1) ? 1 : 0;
        for (int j = 0; j < 8; j++) w.cnt[q * 8 + j] = 0ull;
        if (a.K <= 1) { // no iteration at all: the result is the start node
            RrtCounters c = {0, 0, 0, 0, 0, 0, 0, 0};
            rrt_finish<32>(a, q, g, Q, a.K, 0, 1, -1, TRRT_OK_NOT_FOUND, 0, c);
        }
    }
}

__global__ void __launch_bounds__(TRRT_WAVE_THREADS) wave_scan(const RrtDev a, const WaveDev w, const int k0) {
    __shared__ __align__(16) double2 scan_tiles[(TRRT_WAVE_THREADS / 32) * 4 * TRRT_TILE_PAIRS];
    const int lane = threadIdx.x & 31;
    const int64_t q = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (q >= a.nq || !w.active[q]) return; // whole warps leave together
    double2 *scan_tile = scan_tiles + (threadIdx.x >> 5) * 4 * TRRT_TILE_PAIRS;
    RrtQuery Q;
    rrt_ptrs(a, q, Q);
    const int K = a.K, n0 = w.n[q];
    const int my_it = k0 + lane;
    int pre = TRRT_IT_NOT_RUN;
    double qx = 0, qy = 0, qth = 0, bd = INFINITY;
    int near = 0x7fffffff;
    unsigned long long probes = 0;
    if (my_it < K - 1) {
        const int sx = __ldg(Q.sxy + 2 * my_it), sy = __ldg(Q.sxy + 2 * my_it + 1);
        qx = (double)sx; qy = (double)sy;
        qth = standardangle(__ldg(Q.sth + my_it));
        if (!Q.m.freespace(sx, sy)) pre = TRRT_IT_QRAND_BLOCKED; // rrt.py:148
        else pre = (tree_find(Q.tab, Q.tmask, Q.nx, Q.ny, Q.nth, qx, qy, qth, probes) >= 0) ? TRRT_IT_QRAND_IN_TREE : -1; // rrt.py:151
    }
    nearest_staged(scan_tile, Q.nx, Q.ny, n0, qx, qy, bd, near);
    const int64_t s = q * 32 + lane;
    w.pre[s] = pre; w.near[s] = near; w.bd[s] = bd; w.qx[s] = qx; w.qy[s] = qy; w.qth[s] = qth;
    if (a.counters && probes) atomicAdd(&w.cnt[q * 8 + 7], probes);
}

// one thread per (query, lane): everything between "nearest node chosen" and "edge tested" (rrt.py:161-176)
__global__ void __launch_bounds__(TRRT_WAVE_SMALL) wave_expand(const RrtDev a, const WaveDev w) {
    const Group<1> solo;
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t q = s >> 5;
    if (q >= a.nq || !w.active[q]) return;
    Expand e;
    e.code = TRRT_IT_NOT_RUN; e.flags = 0; e.lospx = e.arcpx = e.arcang = e.drive = 0;
    e.wx = e.wy = e.wth = NAN; e.usteer = e.iccx = e.iccy = e.rad = e.udist = 0;
    int exist = -1;
    if (w.pre[s] == -1) {
        RrtQuery Q;
        rrt_ptrs(a, q, Q);
        const int near = w.near[s];
        unsigned long long probes = 0;
        expand_from<1>(solo, Q.m, a.P, Q.nx[near], Q.ny[near], Q.nth[near], w.qx[s], w.qy[s], w.qth[s], Q.gx, Q.gy, Q.gth, e);
        if (e.code == EX_ACCEPT) exist = tree_find(Q.tab, Q.tmask, Q.nx, Q.ny, Q.nth, e.wx, e.wy, e.wth, probes);
        if (a.counters && probes) atomicAdd(&w.cnt[q * 8 + 7], probes);
    }
    w.code[s] = e.code; w.flags[s] = e.flags; w.exist[s] = exist; w.aux[s] = e.drive;
    w.lospx[s] = e.lospx; w.arcpx[s] = e.arcpx; w.arcang[s] = e.arcang;
    w.wx[s] = e.wx; w.wy[s] = e.wy; w.wth[s] = e.wth;
    w.usteer[s] = e.usteer; w.iccx[s] = e.iccx; w.iccy[s] = e.iccy; w.rad[s] = e.rad; w.udist[s] = e.udist;
}

// one warp per query: predict the lanes that the commit will have to re-expand, and re-expand them now, all queries
// at once.  Lane j is re-expanded when a node inserted earlier in its window is strictly nearer than its snapshot
// winner.  The nodes of the window are not known yet, but the TENTATIVE ones are: lane i will insert (wx_i, wy_i) if
// its edge was accepted and its node is new.  So lane j takes the nearest tentative node of the lanes i < j (first
// minimum, like the commit) and, if it beats the snapshot winner, expands from it.  The commit checks the prediction
// (same lane, and that lane really inserted its tentative node) and falls back to its own one-lane expansion otherwise.
__global__ void __launch_bounds__(TRRT_WAVE_SMALL) wave_reexpand(const RrtDev a, const WaveDev w) {
    const Group<32> g;
    const Group<1> solo;
    const int64_t q = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (q >= a.nq || !w.active[q]) return;
    const int64_t s = q * 32 + g.gl;
    const int pre = w.pre[s];
    const double qx = w.qx[s], qy = w.qy[s];
    const double bd = (pre == -1) ? w.bd[s] : INFINITY;
    const bool tentative = (pre == -1) && w.code[s] == EX_ACCEPT && w.exist[s] < 0;
    const double mx = w.wx[s], my = w.wy[s];
    double best = INFINITY;
    int from = -1;
    const unsigned acc = g.ballot(tentative);
    for (unsigned m = acc; m; m &= m - 1) { // lanes with a tentative node, in iteration order
        const int i = __ffs(m) - 1;
        const double vx = g.bcast(mx, i), vy = g.bcast(my, i);
        if (i < g.gl) {
            const double dx = qx - vx, dy = qy - vy;
            const double d = dx * dx + dy * dy;
            if (d < best) { best = d; from = i; }
        }
    }
    if (pre != -1 || !(best < bd)) from = -1;
    w.a_from[s] = from;
    if (from >= 0) {
        RrtQuery Q;
        rrt_ptrs(a, q, Q);
        const int64_t sf = q * 32 + from;
        Expand e;
        unsigned long long probes = 0;
        expand_from<1>(solo, Q.m, a.P, w.wx[sf], w.wy[sf], w.wth[sf], qx, qy, w.qth[s], Q.gx, Q.gy, Q.gth, e);
        int exist = -1;
        if (e.code == EX_ACCEPT) exist = tree_find(Q.tab, Q.tmask, Q.nx, Q.ny, Q.nth, e.wx, e.wy, e.wth, probes);
        w.a_code[s] = e.code; w.a_flags[s] = e.flags; w.a_exist[s] = exist; w.a_aux[s] = e.drive;
        w.a_lospx[s] = e.lospx; w.a_arcpx[s] = e.arcpx; w.a_arcang[s] = e.arcang;
        w.a_wx[s] = e.wx; w.a_wy[s] = e.wy; w.a_wth[s] = e.wth;
        w.a_usteer[s] = e.usteer; w.a_iccx[s] = e.iccx; w.a_iccy[s] = e.iccy; w.a_rad[s] = e.rad; w.a_udist[s] = e.udist;
    }
}

// one warp per query: commit the window in iteration order (phase B of rrt_kernel_spec, same code path)
__global__ void __launch_bounds__(TRRT_WAVE_SMALL) wave_commit(const RrtDev a, const WaveDev w, const int k0) {
    const Group<32> g;
    const Group<1> solo;
    const int64_t q = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (q >= a.nq || !w.active[q]) return;
    const int K = a.K;
    RrtQuery Q;
    rrt_ptrs(a, q, Q);
    const int64_t s = q * 32 + g.gl;
    const int pre = w.pre[s];
    bool q_in_tree = (pre == TRRT_IT_QRAND_IN_TREE);
    int near = w.near[s], exist = w.exist[s];
    const double bd = (pre == -1) ? w.bd[s] : INFINITY, qx = w.qx[s], qy = w.qy[s], qth = w.qth[s];
    Expand e;
    e.code = w.code[s]; e.flags = w.flags[s]; e.drive = w.aux[s];
    e.lospx = w.lospx[s]; e.arcpx = w.arcpx[s]; e.arcang = w.arcang[s];
    e.wx = w.wx[s]; e.wy = w.wy[s]; e.wth = w.wth[s];
    e.usteer = w.usteer[s]; e.iccx = w.iccx[s]; e.iccy = w.iccy[s]; e.rad = w.rad[s]; e.udist = w.udist[s];
    double wbest = INFINITY;
    int widx = -1, wlane = -1;     // nearest node inserted earlier in this window, and the lane that inserted it
    const int a_from = w.a_from[s];
    unsigned changed = 0;          // lanes whose committed expansion is not their first one (uniform)
    int n = w.n[q], nlos = w.nlos[q], sol = w.sol[q], status = w.status[q], iters = w.iters[q];
    RrtCounters c = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned long long probes = 0;
    bool running = true;
    for (int j = 0; j < 32; j++) {
        const int it = k0 + j;
        int pre_j = g.bcast(pre, j);
        if (pre_j == TRRT_IT_NOT_RUN) break;
        const bool in_tree_j = g.bcast((int)q_in_tree, j);
        if (pre_j == TRRT_IT_QRAND_IN_TREE) pre_j = -1; // the flag travels in q_in_tree from here on
        int code = pre_j, near_j = -1, newi = -1;
        bool go = true;
        if (pre_j != TRRT_IT_QRAND_BLOCKED) {
            if (in_tree_j) code = TRRT_IT_QRAND_IN_TREE; // rrt.py:151
            else {
                if (g.bcast((int)(wbest < bd), j)) {
                    // a node of this window is strictly nearer: lane j's iteration starts from it.  wave_reexpand has
                    // usually done that expansion already; it is valid if it started from the tentative node of the
                    // same lane and that lane committed its first expansion.
                    const int wl = g.bcast(wlane, j);
                    const bool predicted = g.bcast((int)(a_from == wlane), j) && !((changed >> wl) & 1u);
                    changed |= 1u << j;
                    g.sync(); // nodes written by earlier steps are visible to lane j
                    if (g.gl == j) {
                        near = widx;
                        if (predicted) {
                            e.code = w.a_code[s]; e.flags = w.a_flags[s]; e.drive = w.a_aux[s];
                            e.lospx = w.a_lospx[s]; e.arcpx = w.a_arcpx[s]; e.arcang = w.a_arcang[s];
                            e.wx = w.a_wx[s]; e.wy = w.a_wy[s]; e.wth = w.a_wth[s];
                            e.usteer = w.a_usteer[s]; e.iccx = w.a_iccx[s]; e.iccy = w.a_iccy[s]; e.rad = w.a_rad[s]; e.udist = w.a_udist[s];
                        } else {
                            expand_from<1>(solo, Q.m, a.P, Q.nx[near], Q.ny[near], Q.nth[near], qx, qy, qth, Q.gx, Q.gy, Q.gth, e);
                        }
                        // the index now also holds the nodes of this window: one probe answers `qnew in G`
                        exist = -1;
                        if (e.code == EX_ACCEPT) exist = tree_find(Q.tab, Q.tmask, Q.nx, Q.ny, Q.nth, e.wx, e.wy, e.wth, probes);
                    }
                }
                near_j = g.bcast(near, j);
                const int ecode = g.bcast(e.code, j), eflags = g.bcast(e.flags, j);
                const bool mine = g.gl == j;
                if (a.counters) { c.scan += (unsigned long long)n; if (mine) c.steer++; }
                if (ecode == TRRT_IT_STEER_CONSTRAINT) code = TRRT_IT_STEER_CONSTRAINT;
                else {
                    const int nl = (eflags >> 4) & 3;
                    if (mine) {
                        if (Q.los_log) {
                            if (nl >= 1) Q.los_log[nlos] = (eflags >> 6) & 1;
                            if (nl >= 2) Q.los_log[nlos + 1] = (eflags >> 7) & 1;
                        }
                        if (a.counters) { c.los += nl; c.lospx += e.lospx; c.arcpx += e.arcpx; c.arcang += e.arcang; c.drive += e.drive; }
                    }
                    nlos += nl;
                    if (eflags & 2) { status = TRRT_ERR_REF_RAISES_DRIVE_NONE; code = TRRT_IT_NOT_RUN; go = false; }
                    else if (ecode == TRRT_IT_ARC_BLOCKED) code = TRRT_IT_ARC_BLOCKED;
                    else { // rrt.py:179-201
                        int idx = g.bcast(exist, j);
                        if (idx < 0) {
                            if (n >= K) { status = TRRT_ERR_CAPACITY; code = TRRT_IT_NOT_RUN; go = false; }
                            else {
                                idx = n++;
                                code = TRRT_IT_NEW_NODE;
                                if (mine) {
                                    Q.nx[idx] = e.wx; Q.ny[idx] = e.wy; Q.nth[idx] = e.wth;
                                    tree_insert(Q.tab, Q.tmask, e.wx, e.wy, e.wth, idx);
                                }
                                const double vx = g.bcast(e.wx, j), vy = g.bcast(e.wy, j), vth = g.bcast(e.wth, j);
                                if (g.gl > j) {
                                    const double dx = qx - vx, dy = qy - vy;
                                    const double d = dx * dx + dy * dy;
                                    if (d < wbest) { wbest = d; widx = idx; wlane = j; }
                                    if (qx == vx && qy == vy && qth == vth) q_in_tree = true;
                                    if (exist < 0 && e.wx == vx && e.wy == vy && e.wth == vth) exist = idx;
                                }
                            }
                        } else code = TRRT_IT_EXISTING_NODE;
                        if (go) {
                            newi = idx;
                            if (idx != near_j && mine) { // rrt.py:187-188
                                Q.parent[idx] = near_j;
                                if (Q.uo) {
                                    const bool st = eflags & 1;
                                    Q.uo[5 * idx] = e.usteer; Q.uo[5 * idx + 1] = st ? NAN : e.iccx; Q.uo[5 * idx + 2] = st ? NAN : e.iccy;
                                    Q.uo[5 * idx + 3] = st ? NAN : e.rad; Q.uo[5 * idx + 4] = e.udist;
                                }
                            }
                            if (eflags & 4) { sol = idx; status = TRRT_OK_FOUND; go = false; }
                        }
                    }
                }
            }
        }
        if (g.gl == j) {
            if (Q.it_near) Q.it_near[it] = near_j;
            if (Q.it_new) Q.it_new[it] = newi;
            if (Q.it_code) Q.it_code[it] = (uint8_t)code;
        }
        iters = it + 1;
        if (!go) {
            if (status != TRRT_OK_FOUND) iters = it; // the iteration that raises is not counted
            running = false;
            break;
        }
    }
    g.sync();
    if (a.counters) { // accumulate this window's counters (lane-private sums folded here)
        c.probe = probes;
        c.los = g.sum(c.los); c.lospx = g.sum(c.lospx); c.arcpx = g.sum(c.arcpx); c.arcang = g.sum(c.arcang);
        c.steer = g.sum(c.steer); c.drive = g.sum(c.drive); c.probe = g.sum(c.probe);
        if (g.gl == 0) {
            unsigned long long *o = w.cnt + q * 8;
            o[0] += c.scan; o[1] += c.los; o[2] += c.lospx; o[3] += c.arcpx; o[4] += c.arcang; o[5] += c.steer; o[6] += c.drive; o[7] += c.probe;
        }
        g.sync();
    }
    const bool done = !running || k0 + 32 >= K - 1;
    if (done) {
        RrtCounters t = {0, 0, 0, 0, 0, 0, 0, 0};
        if (a.counters) { const unsigned long long *o = w.cnt + q * 8; t = RrtCounters{o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]}; }
        rrt_finish<32>(a, q, g, Q, K, iters, n, sol, status, nlos, t);
    }
    if (g.gl == 0) { w.n[q] = n; w.nlos[q] = nlos; w.sol[q] = sol; w.status[q] = status; w.iters[q] = iters; if (done) w.active[q] = 0; }
}

} // namespace trrt
