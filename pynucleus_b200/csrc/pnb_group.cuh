// Dense assembly, 2D: cell-group path.
//
// Cells are ordered along a Hilbert curve and cut into groups of GC cells.  A unit is a pair of groups
// (I <= J); every cell pair belongs to exactly one unit, so every pair is evaluated exactly once (the DoF-tile
// path evaluates pairs whose cells straddle tile borders once per tile: ~2x).  The unit accumulates the cross
// blocks of its pairs in a shared-memory block over (local dofs of I) x (local dofs of J) and adds the block to
//     U[dofs(I), dofs(J)]                (one orientation only; the operator is F = U + U^T, see symmetrize_kernel)
// Groups that share a vertex get different colours; units are launched in phases (colour of I, colour of J):
// two units of one phase never touch the same entry of U, and the phases are ordered by kernel launches, so
// the read-modify-write of U needs no atomics and the summation order is fixed (bitwise reproducible).
//
// Three kernels (unit kinds are decided on the host from bounding boxes with a rigorous bound on getQuadOrder):
//   gf2_kernel   units whose pairs all have order 2 (3-node rule): no classification, unrolled 3x3 evaluation
//   gmix_kernel  every other unit: classification, binning by order, thread-per-pair evaluation of orders
//                2..5; pairs it does not take (touching pairs, higher orders) are fetched from the results of
//   gnear_eval_kernel (runs first): singular pairs and regular pairs of order > 5, one warp per slice of at
//                most PNB_NEAR_ITEM quadrature nodes of a pair, from a pair list built once per problem
//                (gnear_list_kernel); equal-sized items keep all SMs busy
// Cell-diagonal blocks (xx / yy of nonlocalOperator_{SCALAR}.pxi:769-789) are staged per (partner group, cell):
// slot Dp[g][c] has exactly one writer, the unit (group(c), g).
#pragma once

// Several GPUs (row sets per part, see pnb_dist_plan in pnb200.cu): a unit does not add its block to the matrix but
// stores it, row by row, into the staging buffer of the part that owns the row -- once as rows of I (columns dofs(J))
// and once transposed as rows of J (columns dofs(I)).  Every (unit, row) fragment has its own place in the staging
// buffer of its destination (plain stores through peer memory over NVLink: no read-modify-write, no atomics, no
// collective); the owner sums the fragments of its rows in a fixed order afterwards (dist_apply_kernel).
#define PNB_MAX_PARTS 16
struct DistSched {
    int nparts, part;
    double *stage[PNB_MAX_PARTS];     // staging buffer of every part (peer-mapped device pointers)
    const unsigned char *gown;        // per group-local dof (gdptr[g] + l): part that owns the row
    const unsigned char *gpos;        // ... its position among the dofs of the group owned by that part
    const int *gcnt;                  // ngroups x nparts: dofs of group g owned by part o
    const long long *uoff_f2;         // per unit of the f2 list x nparts: start of its fragments in stage[o]
    const long long *uoff_mix;        // the same for the mix list
};

struct GroupSched {
    DistSched dist;
    int ngroups, cap, maxld, ldS, ncolors;
    const int *gptr;     // ngroups+1: first cell slot of a group (multiples of PNB_SB)
    const int *gcells;   // cell id per slot, -1 = padding; batches of PNB_SB slots share no vertex
    const int *gloc;     // packed group-local dof index of the 3 vertices (8 bits each, 0xFF = no dof)
    const int *gdptr;    // ngroups+1
    const int *gdofs;    // group-local dof -> global dof
    double *Dp;          // ngroups x nc x ND
    int nbmax;           // largest number of batches of a group
    const int4 *npairs;  // near pair list: (row cell, column cell, panel, first item)
    const int *nearbase; // [near slot][row batch][column batch]: position of the first pair of the sub-batch
    const unsigned char *nearrow;   // ... x PNB_SB: pairs of the sub-batch in earlier rows
    const int *gincptr;  // per group-local dof (gdptr[g] + l, one more at the end): its (cell slot, local vertex) incidences
    const unsigned short *ginc;     // slot * 4 + vertex, ascending
    const double *R;     // results of the near items, NL doubles each
    const double *F;     // per near pair: finished cross block (9) and cell-diagonal blocks (6 + 6)
    // ordered updates of U: every unit list is processed by persistent CTAs in list order (tickets); a unit adds
    // its block to U only after all units of smaller ticket that touch the same entries have done so
    const int *adjptr;   // ngroups+1: groups sharing a vertex (sorted, includes the group itself)
    const int *adj;
    const int *ticket;   // ngroups x ngroups: list position | (list << 30) of unit (I <= J), -1 = not evaluated here
    int *done;           // completion flags: list 0 (f2) at [0, nf2), list 1 (mix) at [nf2, nf2 + nmix)
    int *counters_i;     // ticket counters, one per launch: [2 * panel + list]
    int nlist0;
    int *err;
    unsigned long long *counters;
};

struct GUnit { int I, J, kind, slot; };   // kind 0: uniform order 2, 1: orders <= 5, 2: near (slot = staging slot)

// constants of the 3-node rule, passed as kernel argument (constant bank operands)
struct F2Rule {
    double bary[3][3];   // [vertex][node]
    double wphi[3][3];   // [node][vertex] = w[node] * bary[vertex][node]
    double w[3];
    double qq[6][3];     // w[node] * bary[a][node] * bary[b][node], a <= b
    double c[8];         // binomial series of the power function
    int eoff, pad;       // PowTab::eoff minus the exponent bias
};

// power function of the order-2 units: replicated shared-memory table (lane-private bank group, see PowTabS),
// series coefficients as constant-bank operands; degree 6
__device__ __forceinline__ double f2_pow(const double2 *__restrict__ itl, const double *__restrict__ T1, const F2Rule &R, double d2)
{
    const int hi = __double2hiint(d2), lo = __double2loint(d2);
    const int E = min(max(((hi >> 20) & 0x7ff) + R.eoff, 0), 255);
    const int idx = (hi >> 10) & (0x7f * PNB_POW_REP);
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double2 it = itl[idx];
    const double r = fma(m, it.x, -1.0);
    double p = fma(R.c[6], r, R.c[5]);
    p = fma(p, r, R.c[4]);
    p = fma(p, r, R.c[3]);
    p = fma(p, r, R.c[2]);
    p = fma(p, r, R.c[1]);
    p = fma(p, r, R.c[0]);
    return T1[E] * (it.y * p);
}

__device__ __forceinline__ unsigned char *carve(unsigned char *&p, size_t bytes)
{
    unsigned char *r = p;
    p += (bytes + 15) & ~(size_t)15;
    return r;
}


// ---- ordered update of U -------------------------------------------------------------------------
// Unit (I,J) writes U[dofs(I), dofs(J)].  It shares entries exactly with the units (a,b), a <= b, whose row group a
// touches I and whose column group b touches J.  Tickets are handed out in list order, so a waiting CTA only
// waits for CTAs that are already running or done: no deadlock, and the order of the additions to every entry of
// U is the list order (bitwise reproducible).
__device__ __forceinline__ void g_wait_predecessors(const GroupSched &G, int listid, int myticket, int I, int J, int tid, int nthreads)
{
    const int a0 = G.adjptr[I], na = G.adjptr[I + 1] - a0, b0 = G.adjptr[J], nb = G.adjptr[J + 1] - b0;
    const volatile int *done = G.done + (listid ? G.nlist0 : 0);
    for (int idx = tid; idx < na * nb; idx += nthreads) {
        const int a = G.adj[a0 + idx / nb], b = G.adj[b0 + idx % nb];
        if (a > b) continue;
        const int tk = G.ticket[(size_t)a * G.ngroups + b];
        if (tk < 0 || (tk >> 30) != listid) continue;
        const int t = tk & 0x3FFFFFFF;
        if (t >= myticket) continue;
        while (done[t] == 0) __nanosleep(64);
    }
    __threadfence();
    __syncthreads();
}

__device__ __forceinline__ void g_signal_done(const GroupSched &G, int listid, int myticket, int tid)
{
    __threadfence();
    __syncthreads();
    if (tid == 0) *((volatile int *)(G.done + (listid ? G.nlist0 : 0) + myticket)) = 1;
}

// adds the unit block to U (L2 operations: other SMs update neighbouring entries of the same lines)
__device__ __forceinline__ void g_flush_block(const GroupSched &G, const double *S, int ldS, int dI, int nldI, int dJ, int nldJ, double *A,
                                              int64_t ld, int tid, int nthreads)
{
    // four independent read-modify-writes in flight per thread (L2 latency)
    const int n = nldI * nldJ;
    for (int e0 = tid; e0 < n; e0 += 4 * nthreads) {
        double *dst[4];
        double o[4], v[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int e = e0 + t * nthreads;
            dst[t] = nullptr;
            o[t] = v[t] = 0.;
            if (e < n) {
                const int a = e / nldJ, b = e - a * nldJ;
                dst[t] = &A[(size_t)G.gdofs[dI + a] * ld + G.gdofs[dJ + b]];
                v[t] = S[a * ldS + b];
            }
        }
#pragma unroll
        for (int t = 0; t < 4; t++) if (dst[t]) o[t] = __ldcg(dst[t]);
#pragma unroll
        for (int t = 0; t < 4; t++) if (dst[t]) __stcg(dst[t], o[t] + v[t]);
    }
}

// several GPUs: the unit block goes to the staging buffers of the row owners (see DistSched).  Both passes write runs of
// consecutive addresses (the transposed pass reads the block column by column: ldS is odd, no bank conflicts).
__device__ __forceinline__ void g_flush_staged(const GroupSched &G, const long long *__restrict__ uoff, const double *S, int ldS, int I,
                                               int dI, int nldI, int dJ, int nldJ, int tid, int nthreads)
{
    const DistSched &D = G.dist;
    for (int e = tid; e < nldI * nldJ; e += nthreads) {
        const int a = e / nldJ, b = e - a * nldJ;
        const int o = D.gown[dI + a];
        D.stage[o][uoff[o] + (long long)D.gpos[dI + a] * nldJ + b] = S[a * ldS + b];
    }
    for (int e = tid; e < nldI * nldJ; e += nthreads) {
        const int b = e / nldI, a = e - b * nldI;
        const int o = D.gown[dJ + b];
        D.stage[o][uoff[o] + (long long)D.gcnt[I * D.nparts + o] * nldJ + (long long)D.gpos[dJ + b] * nldI + a] = S[a * ldS + b];
    }
}

// -------------------------------------------------------------------------------------------------
// uniform order 2.  PNB_F2T = 256 threads, two CTAs per SM, one 16 x 16 sub-batch per step (thread = cell pair).
// Warp w holds the row cells 2w, 2w+1 of the row batch during the whole sweep over the column batches: the row cells of
// a batch share no vertex, so no other warp touches its rows of the unit block, and the column cells of a step share
// no vertex either -- the block updates need no barrier and no atomics.
// Cell-diagonal blocks: xx[e] = sum_i qq[e][i] r_i and yy[e] = sum_j qq[e][j] c_j are linear in the row sums
// r_i / column sums c_j of the kernel matrix, so only those (3 + 3 values per pair) are reduced over the
// partner cells; qq is applied once per cell at the end.  Row sums stay in registers over the sweep; the column sums
// of a step are combined over the 8 warps through a double-buffered staging array: ONE barrier per step.
// -------------------------------------------------------------------------------------------------
#define PNB_GT 512
#define PNB_F2T 256
inline size_t gf2_smem_bytes(int cap, int maxld, int ldS)
{
    size_t b = 0;
    auto add = [&](size_t x) { b += (x + 15) & ~(size_t)15; };
    add(sizeof(PowTabS));
    add((size_t)maxld * ldS * 8);
    add((size_t)6 * cap * 8);         // nodes of the column side
    add((size_t)cap * 8);             // vol of the column side
    add((size_t)2 * cap * 4);         // cell, loc of the column side
    add((size_t)2 * 8 * 16 * 3 * 8);  // Yw: column sums of a step per warp, two buffers
    add((size_t)cap * 3 * 8);         // column sums
    return b;
}

__global__ void __launch_bounds__(PNB_F2T, 2)
gf2_kernel(DProblem P, GroupSched G, const GUnit *__restrict__ units, int first, int nunits, int *__restrict__ next,
           double *__restrict__ A, int64_t ld, F2Rule R)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *sp = smem_raw;
    const int cap = G.cap, ldS = G.ldS;
    PowTabS *pw = reinterpret_cast<PowTabS *>(carve(sp, sizeof(PowTabS)));
    double *S = reinterpret_cast<double *>(carve(sp, (size_t)G.maxld * ldS * 8));
    double *yj = reinterpret_cast<double *>(carve(sp, (size_t)6 * cap * 8));
    double *volj = reinterpret_cast<double *>(carve(sp, (size_t)cap * 8));
    int *cellj = reinterpret_cast<int *>(carve(sp, (size_t)2 * cap * 4));
    int *locj = cellj + cap;
    double *Yw = reinterpret_cast<double *>(carve(sp, (size_t)2 * 8 * 16 * 3 * 8));
    double *CYs = reinterpret_cast<double *>(carve(sp, (size_t)cap * 3 * 8));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ int s_ticket;
    powtab_stage(pw, P.pow_int, tid, PNB_F2T);
    const double2 *itl = pw->IT + (lane & (PNB_POW_REP - 1));
    const double *T1s = pw->T1;
    unsigned long long my_pairs = 0;
    const int k1 = tid >> 4, k2 = tid & 15;
    for (;;) {
    __syncthreads();
    if (tid == 0) s_ticket = first + atomicAdd(next, 1);     // list positions [first, nunits) of this launch
    __syncthreads();
    const int ticket = s_ticket;
    if (ticket >= nunits) break;
    const GUnit u = units[ticket];
    const int I = u.I, J = u.J;
    const int ibeg = G.gptr[I], nI = G.gptr[I + 1] - ibeg, jbeg = G.gptr[J], nJ = G.gptr[J + 1] - jbeg;
    const int dI = G.gdptr[I], nldI = G.gdptr[I + 1] - dI, dJ = G.gdptr[J], nldJ = G.gdptr[J + 1] - dJ;
    {
        for (int e = tid; e < nldI * ldS; e += PNB_F2T) S[e] = 0.;
        for (int e = tid; e < cap * 3; e += PNB_F2T) CYs[e] = 0.;
        for (int s = tid; s < nJ; s += PNB_F2T) {
            const int c = G.gcells[jbeg + s];
            cellj[s] = c;
            locj[s] = G.gloc[jbeg + s];
            if (c >= 0) {
                const double *v = P.simplices + (size_t)c * 6;
                volj[s] = P.vol[c];
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    yj[(2 * q) * cap + s] = R.bary[0][q] * v[0] + R.bary[1][q] * v[2] + R.bary[2][q] * v[4];
                    yj[(2 * q + 1) * cap + s] = R.bary[0][q] * v[1] + R.bary[1][q] * v[3] + R.bary[2][q] * v[5];
                }
            }
        }
    }
    __syncthreads();
    int step = 0;
    for (int rb = 0; rb < nI; rb += PNB_SB) {
        const int s1 = rb + k1;
        const int c1 = G.gcells[ibeg + s1];
        const int l1 = G.gloc[ibeg + s1];
        double x[3][2];
        double v1 = 0.;
        if (c1 >= 0) {
            const double *v = P.simplices + (size_t)c1 * 6;
            const double v0 = v[0], v1_ = v[1], v2 = v[2], v3 = v[3], v4 = v[4], v5 = v[5];
#pragma unroll
            for (int q = 0; q < 3; q++) {
                x[q][0] = R.bary[0][q] * v0 + R.bary[1][q] * v2 + R.bary[2][q] * v4;
                x[q][1] = R.bary[0][q] * v1_ + R.bary[1][q] * v3 + R.bary[2][q] * v5;
            }
            v1 = 2.0 * P.vol[c1];
        } else {
#pragma unroll
            for (int q = 0; q < 3; q++) x[q][0] = x[q][1] = 0.;
        }
        // rows of the unit block that belong to the three dofs of the row cell (-1: no dof)
        int ro[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const int ra = (l1 >> (8 * a)) & 0xFF;
            ro[a] = ra == 0xFF ? -1 : ra * ldS;
        }
        double rx[3] = {0., 0., 0.};
        for (int cb = 0; cb < nJ; cb += PNB_SB, step++) {
            const int s2 = cb + k2;
            const int c2 = cellj[s2];
            const int l2 = locj[s2];
            double cy[3] = {0., 0., 0.};
            // a pair is skipped only when neither cell carries a dof (as the reference does)
            const bool live = c1 >= 0 && c2 >= 0 && !((l1 & 0x00FFFFFF) == 0x00FFFFFF && (l2 & 0x00FFFFFF) == 0x00FFFFFF);
            if (live) {
                my_pairs++;
                double g[3][3];
                {
                    // the nine powers stage by stage (see PowCtxT::batch): nine independent FMA chains in flight
                    double rr[9], yv[9], tv[9], pp[9];
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        const double y0 = yj[(2 * j) * cap + s2], y1 = yj[(2 * j + 1) * cap + s2];
#pragma unroll
                        for (int i = 0; i < 3; i++) {
                            const double a = x[i][0] - y0, b = x[i][1] - y1;
                            const double d2 = a * a + b * b;
                            const int hi = __double2hiint(d2), lo = __double2loint(d2);
                            const int E = min(max(((hi >> 20) & 0x7ff) + R.eoff, 0), 255);
                            const int idx = (hi >> 10) & (0x7f * PNB_POW_REP);
                            const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
                            const double2 it = itl[idx];
                            tv[i * 3 + j] = T1s[E];
                            yv[i * 3 + j] = it.y;
                            rr[i * 3 + j] = fma(m, it.x, -1.0);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 9; k++) pp[k] = fma(R.c[6], rr[k], R.c[5]);
#pragma unroll
                    for (int q = 4; q >= 0; q--) {
#pragma unroll
                        for (int k = 0; k < 9; k++) pp[k] = fma(pp[k], rr[k], R.c[q]);
                    }
#pragma unroll
                    for (int k = 0; k < 9; k++) g[k / 3][k % 3] = tv[k] * (yv[k] * pp[k]);
                }
                const double sc = v1 * volj[s2];
                double X[9];
#pragma unroll
                for (int k = 0; k < 9; k++) X[k] = 0.;
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    double t0 = 0., t1 = 0., t2 = 0., r = 0.;
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        t0 = fma(g[i][j], R.wphi[j][0], t0);
                        t1 = fma(g[i][j], R.wphi[j][1], t1);
                        t2 = fma(g[i][j], R.wphi[j][2], t2);
                        r = fma(g[i][j], R.w[j], r);
                    }
                    t0 *= sc; t1 *= sc; t2 *= sc;
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        const double q = R.wphi[i][a];
                        X[a * 3 + 0] = fma(-q, t0, X[a * 3 + 0]);
                        X[a * 3 + 1] = fma(-q, t1, X[a * 3 + 1]);
                        X[a * 3 + 2] = fma(-q, t2, X[a * 3 + 2]);
                    }
                    rx[i] = fma(r, sc, rx[i]);
                }
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    double c = 0.;
#pragma unroll
                    for (int i = 0; i < 3; i++) c = fma(g[i][j], R.w[i], c);
                    cy[j] = c * sc;
                }
                // conflict free: this warp owns its rows for the sweep, the 16 column cells of a step share no vertex
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    const int cbk = (l2 >> (8 * b)) & 0xFF;
                    if (cbk == 0xFF) continue;
#pragma unroll
                    for (int a = 0; a < 3; a++)
                        if (ro[a] >= 0) S[ro[a] + cbk] += X[a * 3 + b];
                }
            }
            // column sums: the two row cells of the warp, then the 8 warps through shared memory
#pragma unroll
            for (int e = 0; e < 3; e++) cy[e] += __shfl_xor_sync(0xffffffffu, cy[e], 16);
            double *yw = Yw + (size_t)(step & 1) * (8 * 16 * 3);
            if (lane < 16) {
#pragma unroll
                for (int e = 0; e < 3; e++) yw[(warp * 16 + k2) * 3 + e] = cy[e];
            }
            __syncthreads();     // the one barrier of the step: publishes yw (the other buffer is written in the next step)
            if (lane < 6) {
                const int t = warp * 6 + lane;        // 48 values: (column cell, node)
                double s = 0.;
#pragma unroll
                for (int w = 0; w < 8; w++) s += yw[w * 48 + t];
                CYs[cb * 3 + t] += s;
            }
        }
        // row sums over the 16 lanes of the row cell (fixed tree)
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) {
#pragma unroll
            for (int e = 0; e < 3; e++) rx[e] += __shfl_xor_sync(0xffffffffu, rx[e], off);
        }
        if (k2 == 0 && c1 >= 0) {
#pragma unroll
            for (int e = 0; e < 6; e++)
                G.Dp[((size_t)J * P.nc + c1) * 6 + e] = R.qq[e][0] * rx[0] + R.qq[e][1] * rx[1] + R.qq[e][2] * rx[2];
        }
    }
    __syncthreads();
    for (int e = tid; e < nJ * 6; e += PNB_F2T) {
        const int s2 = e / 6, k = e - s2 * 6;
        const int c2 = cellj[s2];
        if (c2 >= 0)
            G.Dp[((size_t)I * P.nc + c2) * 6 + k] = R.qq[k][0] * CYs[s2 * 3] + R.qq[k][1] * CYs[s2 * 3 + 1] + R.qq[k][2] * CYs[s2 * 3 + 2];
    }
    if (G.dist.nparts > 0) g_flush_staged(G, G.dist.uoff_f2 + (size_t)ticket * G.dist.nparts, S, ldS, I, dI, nldI, dJ, nldJ, tid, PNB_F2T);
    else {
        g_wait_predecessors(G, 0, ticket, I, J, tid, PNB_F2T);
        g_flush_block(G, S, ldS, dI, nldI, dJ, nldJ, A, ld, tid, PNB_F2T);
        g_signal_done(G, 0, ticket, tid);
    }
    }
    for (int off = 16; off > 0; off >>= 1) my_pairs += __shfl_xor_sync(0xffffffffu, my_pairs, off);
    if (lane == 0 && my_pairs) atomicAdd(G.counters + 2, my_pairs);
}

// -------------------------------------------------------------------------------------------------
// classification of one slot pair of a sub-batch (shared by the list builder and the unit kernel).
// Returns the panel (order >= 1, or -(shared vertices) for touching pairs), 0 when the slot holds no pair.
// -------------------------------------------------------------------------------------------------
struct GCls {
    const double *cxI, *cxJ;      // [2][capI], [2][cap]
    const float *lhI, *ahI, *lhJ, *ahJ;
    const int *cellI, *locI, *cellJ, *locJ;
    int cap;
    int capI, ioff;               // row side: arrays of capI slots that start at slot ioff of the group (unit kernel: one row batch)
    float cf, sf;
};

__device__ __forceinline__ int g_classify(const DProblem &P, const GCls &c, bool diag, bool maybe_touching, int rb, int cb, int k1, int k2)
{
    const int s1 = rb + k1 - c.ioff, s2 = cb + k2;
    const int K1 = c.cellI[s1], K2 = c.cellJ[s2];
    // diagonal units: every unordered pair once (batches rb <= cb; inside a batch k1 <= k2)
    if (!(K1 >= 0 && K2 >= 0 && K1 != K2 ? (!diag || rb < cb || k1 < k2) : (K1 >= 0 && K1 == K2 && diag))) return 0;
    // the reference skips pairs of cells without any dof
    if ((c.locI[s1] & 0x00FFFFFF) == 0x00FFFFFF && (c.locJ[s2] & 0x00FFFFFF) == 0x00FFFFFF) return 0;
    // piecewise variable kernels: pairs of other classes belong to another problem instance
    if (P.labels && !pnb_class_active(P, P.labels[K1], P.labels[K2])) return 0;
    if (K1 == K2) return -3;
    int panel = 0;
    if (maybe_touching) {
        int v1[3], v2[3];
#pragma unroll
        for (int m = 0; m < 3; m++) { v1[m] = P.cells[(size_t)K1 * 3 + m]; v2[m] = P.cells[(size_t)K2 * 3 + m]; }
        panel = -shared_vertices(v1, 3, v2, 3);
    }
    if (panel == 0) {
        const double a = c.cxI[s1] - c.cxJ[s2], b = c.cxI[c.capI + s1] - c.cxJ[c.cap + s2];
        panel = fast_order_2d(a * a + b * b, c.lhI[s1], c.lhJ[s2], c.ahI[s1], c.ahJ[s2], c.cf, c.sf);
        if (panel < 0) {
            // getPanelType evaluates (c1 <= c2): keep the operand order of the reference
            const int c1 = min(K1, K2), c2 = max(K1, K2);
            const double d = center_distance(P.centers + (size_t)c1 * 2, P.centers + (size_t)c2 * 2, 2);
            panel = quad_order_interior(P, P.h[c1], P.h[c2], d);
        }
    }
    return panel;
}

// work items of a near pair: slices of at most PNB_NEAR_ITEM quadrature nodes, one warp each
#define PNB_NEAR_ITEM 2048
__device__ __forceinline__ int near_slices(const DProblem &P, int panel)
{
    int nodes;
    if (panel >= 1) { const int n = P.reg_cell[panel].n; nodes = n * n; }
    else nodes = panel == -3 ? P.q_id.n : (panel == -2 ? P.q_edge.n : P.q_vertex.n);
    return max(1, (nodes + PNB_NEAR_ITEM - 1) / PNB_NEAR_ITEM);
}

__device__ __forceinline__ void g_load_cls_side(const DProblem &P, const GroupSched &G, int g, int *cell, int *loc, double *cx, float *lh, float *ah,
                                                int cap, int tid)
{
    const int beg = G.gptr[g], n = G.gptr[g + 1] - beg;
    for (int s = tid; s < n; s += PNB_THREADS) {
        const int c = G.gcells[beg + s];
        cell[s] = c;
        loc[s] = G.gloc[beg + s];
        if (c >= 0) {
            cx[s] = P.centers[(size_t)c * 2];
            cx[cap + s] = P.centers[(size_t)c * 2 + 1];
            lh[s] = P.lhf[c];
            ah[s] = P.ahf[c];
        }
    }
}

// -------------------------------------------------------------------------------------------------
// near pair list (built once per problem and table set): one CTA per near unit.  Pairs that the
// thread-per-pair evaluator does not take (touching pairs, orders outside far_mask) are appended in sub-batch /
// slot order to a segment of the global list (segments are reserved with an integer atomic: their order does
// not matter, results are addressed through nearbase).  fill == 0 only counts.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PNB_THREADS)
gnear_list_kernel(DProblem P, GroupSched G, const GUnit *__restrict__ units, int far_mask, int fill, int *cursor, int4 *pairs, int2 *items,
                  int *nearbase, unsigned char *nearrow, int *bins, const int *binbase, int *perm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *sp = smem_raw;
    const int cap = G.cap;
    double *cxs = reinterpret_cast<double *>(carve(sp, (size_t)4 * cap * 8));
    float *lhs = reinterpret_cast<float *>(carve(sp, (size_t)4 * cap * 4));
    int *ints = reinterpret_cast<int *>(carve(sp, (size_t)4 * cap * 4));
    __shared__ int wcnt[PNB_THREADS / 32], wits[PNB_THREADS / 32], tot[2], base[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const GUnit u = units[blockIdx.x];
    const int I = u.I, J = u.J;
    const bool diag = I == J;
    const int nI = G.gptr[I + 1] - G.gptr[I], nJ = G.gptr[J + 1] - G.gptr[J];
    GCls c;
    c.cxI = cxs; c.cxJ = cxs + 2 * cap;
    c.lhI = lhs; c.ahI = lhs + cap; c.lhJ = lhs + 2 * cap; c.ahJ = lhs + 3 * cap;
    c.cellI = ints; c.locI = ints + cap; c.cellJ = ints + 2 * cap; c.locJ = ints + 3 * cap;
    c.cap = cap; c.capI = cap; c.ioff = 0;
    c.cf = (float)P.c_int; c.sf = (float)fmax(-0.5 * (P.sing + 2), 0.);
    g_load_cls_side(P, G, I, ints, ints + cap, cxs, lhs, lhs + cap, cap, tid);
    g_load_cls_side(P, G, J, ints + 2 * cap, ints + 3 * cap, cxs + 2 * cap, lhs + 2 * cap, lhs + 3 * cap, cap, tid);
    if (tid < 2) tot[tid] = 0;
    __syncthreads();
    const int k1 = tid / PNB_SB, k2 = tid % PNB_SB;
    int run_pairs = 0, run_items = 0;
    for (int pass = 0; pass < (fill ? 2 : 1); pass++) {
        for (int rb = 0; rb < nI; rb += PNB_SB)
            for (int cb = diag ? rb : 0; cb < nJ; cb += PNB_SB) {
                const int panel = g_classify(P, c, diag, true, rb, cb, k1, k2);
                if (panel > P.max_order) atomicMax(G.err, panel);
                const bool is_far = panel >= 2 && panel <= PNB_FAR_MAX_ORDER && ((far_mask >> panel) & 1);
                const bool near = panel != 0 && !is_far && panel <= P.max_order;
                const int sl = near ? near_slices(P, panel) : 0;
                const unsigned bal = __ballot_sync(0xffffffffu, near);
                // inclusive warp scan of the slice counts
                int inc = sl;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, inc, off);
                    if (lane >= off) inc += t;
                }
                if (pass == 0) {
                    if (lane == 31) { atomicAdd(&tot[0], __popc(bal)); atomicAdd(&tot[1], inc); }
                    // items per evaluation key (regular: the order, singular: 0), counted once (fill == 0)
                    if (near && !fill) atomicAdd(bins + (panel >= 1 ? min(panel, 63) : 0), sl);
                    continue;
                }
                if (lane == 31) { wcnt[warp] = __popc(bal); wits[warp] = inc; }
                __syncthreads();
                int pp = 0, pi = 0, tp = 0, ti = 0;
                for (int w = 0; w < PNB_THREADS / 32; w++) {
                    if (w < warp) { pp += wcnt[w]; pi += wits[w]; }
                    tp += wcnt[w]; ti += wits[w];
                }
                const size_t sbi = ((size_t)u.slot * G.nbmax + rb / PNB_SB) * G.nbmax + cb / PNB_SB;
                if (tid == 0) nearbase[sbi] = base[0] + run_pairs;
                // pairs of the sub-batch in the rows before k1 (the unit kernel works row by row)
                if (k2 == 0) nearrow[sbi * PNB_SB + k1] = (unsigned char)(pp + __popc(bal & ((1u << lane) - 1)));
                if (near) {
                    const int pos = base[0] + run_pairs + pp + __popc(bal & ((1u << lane) - 1));
                    const int it0 = base[1] + run_items + pi + inc - sl;
                    pairs[pos] = make_int4(c.cellI[rb + k1], c.cellJ[cb + k2], panel, it0);
                    const int key = panel >= 1 ? min(panel, 63) : 0;
                    for (int q = 0; q < sl; q++) {
                        items[it0 + q] = make_int2(pos, q);
                        // processing order: grouped by key; the order inside a key does not matter
                        perm[binbase[key] + atomicAdd(bins + 64 + key, 1)] = it0 + q;
                    }
                }
                run_pairs += tp;
                run_items += ti;
                __syncthreads();
            }
        if (pass == 0) {
            __syncthreads();
            if (tid == 0) {
                base[0] = atomicAdd(cursor, tot[0]);
                base[1] = atomicAdd(cursor + 1, tot[1]);
            }
            __syncthreads();
        }
    }
}

// Regular near pair: one column slice of its n x n node pairs per lane group of W lanes (K = 32 / W items per warp).
// The lanes of a group split the ROWS in tiles of PNB_NEAR_R rows; every lane sweeps all columns of the slice with its
// rows in registers, so that the column data (node, weights, products) is read once per PNB_NEAR_R node pairs and
// by all lanes of the group at the same address (broadcast) -- round 1 read it once per node pair and ran at 80 % of
// the shared-memory pipe against 33 % of the FP64 pipe (ncu).  The row sums of nonlocalOperator_{SCALAR}.pxi:769-789
// are factored out: per node pair 5 (distance) + 11 (power) + 4 (row sums) + 1 (column sum) FP64 operations, per
// column 6 more for the second cell's diagonal block, per row 15 for the cross block and the first cell's block.
// The node coordinates are computed once per item, un-fused and in the reference's order (see
// lanes_regular_interior), and kept in shared memory.
template <class KV>
__device__ __forceinline__ void near_regular_group(const DProblem &P, const KV &kv, const double2 *__restrict__ der, int n, int Ka, int Kb,
                                                   int slice, int nsl, double2 *xs, int gl, int W, bool valid, double *acc)
{
    constexpr int R = PNB_NEAR_R;
    const int per = (n + nsl - 1) / nsl;
    const int j0 = slice * per, j1 = min(n, j0 + per), ncol = max(j1 - j0, 0);
    // shared memory of the group: the vertices of the first cell (3 points), then the nodes of the column slice
    double2 *V1 = xs, *Y = xs + 3;
    if (valid) {
        for (int k = gl; k < 3; k += W) V1[k] = make_double2(P.simplices[(size_t)Ka * 6 + 2 * k], P.simplices[(size_t)Ka * 6 + 2 * k + 1]);
        double s2[3][2];
        load_simplex<2>(P.simplices, Kb, 3, s2);
        for (int k = j0 + gl; k < j1; k += W) {
            const double2 ba = der[k * PNB_DER2 + 5], bb = der[k * PNB_DER2 + 6];
            const double b0 = ba.x, b1 = ba.y, b2 = bb.x;
            double y0 = PNB_MUL(b0, s2[0][0]), y1 = PNB_MUL(b0, s2[0][1]);
            y0 = PNB_ADD(y0, PNB_MUL(b1, s2[1][0])); y1 = PNB_ADD(y1, PNB_MUL(b1, s2[1][1]));
            y0 = PNB_ADD(y0, PNB_MUL(b2, s2[2][0])); y1 = PNB_ADD(y1, PNB_MUL(b2, s2[2][1]));
            Y[k - j0] = make_double2(y0, y1);
        }
    }
    __syncwarp();
    double xy[9], xx[6], yy[6];
#pragma unroll
    for (int k = 0; k < 9; k++) xy[k] = 0.;
#pragma unroll
    for (int k = 0; k < 6; k++) xx[k] = yy[k] = 0.;
    if (valid) {
        const int ntiles = (n + R - 1) / R;
        const double2 *dcol = der + (size_t)j0 * PNB_DER2;
        for (int tile = gl; tile < ntiles; tile += W) {
            const int i0 = tile * R;
            double X0[R], X1[R], wq[R], rs[R], t0[R], t1[R], t2[R];
            {
                // nodes of the row tile, un-fused and in the reference's order (nodesInGlobalCoords, quadrature.pyx:76-87)
                const double2 va = V1[0], vb = V1[1], vc = V1[2];
#pragma unroll
                for (int q = 0; q < R; q++) {
                    const int i = min(i0 + q, n - 1);
                    const double2 ba = der[i * PNB_DER2 + 5], bb = der[i * PNB_DER2 + 6];
                    double x0 = PNB_MUL(ba.x, va.x), x1 = PNB_MUL(ba.x, va.y);
                    x0 = PNB_ADD(x0, PNB_MUL(ba.y, vb.x)); x1 = PNB_ADD(x1, PNB_MUL(ba.y, vb.y));
                    x0 = PNB_ADD(x0, PNB_MUL(bb.x, vc.x)); x1 = PNB_ADD(x1, PNB_MUL(bb.x, vc.y));
                    X0[q] = x0; X1[q] = x1;
                    wq[q] = i0 + q < n ? der[i * PNB_DER2].x : 0.;     // rows beyond the rule: weight 0
                    rs[q] = t0[q] = t1[q] = t2[q] = 0.;
                }
            }
#pragma unroll 2
            for (int j = 0; j < ncol; j++) {
                const double2 y = Y[j];
                double g[R], d2[R];
#pragma unroll
                for (int q = 0; q < R; q++) {
                    const double a = X0[q] - y.x, b = X1[q] - y.y;
                    d2[q] = PNB_ADD(PNB_MUL(a, a), PNB_MUL(b, b));
                }
                kv.template batch<R>(d2, g);
                const double2 *dj = dcol + j * PNB_DER2;
                const double2 c0 = dj[0], c1 = dj[1];
                double cw = 0.;
#pragma unroll
                for (int q = 0; q < R; q++) {
                    rs[q] = fma(g[q], c0.x, rs[q]);
                    t0[q] = fma(g[q], c0.y, t0[q]);
                    t1[q] = fma(g[q], c1.x, t1[q]);
                    t2[q] = fma(g[q], c1.y, t2[q]);
                    cw = fma(g[q], wq[q], cw);
                }
                const double2 c2 = dj[2], c3 = dj[3], c4 = dj[4];
                yy[0] = fma(cw, c2.x, yy[0]); yy[1] = fma(cw, c2.y, yy[1]);
                yy[2] = fma(cw, c3.x, yy[2]); yy[3] = fma(cw, c3.y, yy[3]);
                yy[4] = fma(cw, c4.x, yy[4]); yy[5] = fma(cw, c4.y, yy[5]);
            }
#pragma unroll
            for (int q = 0; q < R; q++) {
                if (i0 + q < n) {
                    const double2 *di = der + (size_t)(i0 + q) * PNB_DER2;
                    const double2 d0 = di[0], d1 = di[1], d2 = di[2], d3 = di[3], d4 = di[4];
                    const double wp[3] = {d0.y, d1.x, d1.y};
#pragma unroll
                    for (int aa = 0; aa < 3; aa++) {
                        xy[aa * 3 + 0] = fma(-wp[aa], t0[q], xy[aa * 3 + 0]);
                        xy[aa * 3 + 1] = fma(-wp[aa], t1[q], xy[aa * 3 + 1]);
                        xy[aa * 3 + 2] = fma(-wp[aa], t2[q], xy[aa * 3 + 2]);
                    }
                    xx[0] = fma(d2.x, rs[q], xx[0]); xx[1] = fma(d2.y, rs[q], xx[1]);
                    xx[2] = fma(d3.x, rs[q], xx[2]); xx[3] = fma(d3.y, rs[q], xx[3]);
                    xx[4] = fma(d4.x, rs[q], xx[4]); xx[5] = fma(d4.y, rs[q], xx[5]);
                }
            }
        }
    }
    // the reference's flattened upper triangle of the 6 x 6 local matrix over (dofs of cell 1, dofs of cell 2)
    int k = 0;
#pragma unroll
    for (int II = 0; II < 6; II++)
#pragma unroll
        for (int JJ = II; JJ < 6; JJ++) {
            acc[k++] = (II < 3 && JJ >= 3) ? xy[II * 3 + (JJ - 3)] : (JJ < 3 ? xx[tri_idx(3, II, JJ)] : yy[tri_idx(3, II - 3, JJ - 3)]);
        }
}

// the same with the rule table in global memory (not inlined: one extra copy of the loop, outside the hot path's registers)
template <class KV>
__device__ __noinline__ void near_regular_group_global(const DProblem &P, const KV &kv, const double2 *__restrict__ der, int n, int Ka, int Kb,
                                                       int slice, int nsl, double2 *xs, int gl, int W, bool valid, double *acc)
{
    near_regular_group(P, kv, der, n, Ka, Kb, slice, nsl, xs, gl, W, valid, acc);
}

// Persistent CTAs over chunks of items of one key (regular: the order; 0: singular pairs).  A chunk holds up to
// 8 x K items, one lane group each; the derived rule table of the key is staged in shared memory.
__global__ void __launch_bounds__(PNB_NEAR_THREADS, 2)
gnear_eval_kernel(DProblem P, const int4 *__restrict__ pairs, const int2 *__restrict__ items, const int *__restrict__ perm,
                  const int4 *__restrict__ chunks, int nchunks, double *__restrict__ R, int der_nodes, int warp_points)
{
    constexpr int NV = 3, NL = PairDims<2>::NL, NRr = 2 * NV - 1, NA = NRr * (NRr + 1) / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PowTabS *pw = reinterpret_cast<PowTabS *>(smem_raw);
    double2 *der = reinterpret_cast<double2 *>(smem_raw + sizeof(PowTabS));
    // per warp: per item the vertices of the first cell and the nodes of the column slice (warp_points points)
    double2 *xsw = der + (size_t)PNB_DER2 * der_nodes + (size_t)(threadIdx.x >> 5) * warp_points;
    powtab_stage(pw, P.pow_int, threadIdx.x, PNB_NEAR_THREADS);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const PowCtxS kv(pw, lane);
    int staged = -1;
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const int4 c = chunks[ch];
        const int key = c.x;
        // rules with more nodes than the shared-memory table holds (the few highest orders) are read from global memory
        // (L1 / L2): sizing the table for them cost the second resident CTA of every SM (ncu, round 2: 125 KB per CTA,
        // 8 warps per SM)
        const bool in_smem = key >= 1 && P.reg_cell[key].n <= der_nodes;
        if (in_smem && key != staged) {
            __syncthreads();      // everybody is done with the previous table
            const int n = P.reg_cell[key].n;
            const double2 *src = reinterpret_cast<const double2 *>(P.reg_derived) + (size_t)P.reg_doff[key] * PNB_DER2;
            for (int e = threadIdx.x; e < n * PNB_DER2; e += PNB_NEAR_THREADS) der[e] = src[e];
            staged = key;
            __syncthreads();
        }
        if (key >= 1) {
            const int4 grid = P.reg_grid[key];
            const int W = grid.x, K = grid.y, g = lane / W, gl = lane - g * W;
            const int idx = warp * K + g;
            const bool valid = g < K && idx < c.z;
            const int item = valid ? perm[c.y + idx] : 0;
            const int2 it = items[item];
            const int4 pr = pairs[it.x];
            const int n = P.reg_cell[key].n;
            const int nsl = near_slices(P, key);
            const int per = (n + nsl - 1) / nsl;
            double acc[NL];
            __syncwarp();
            if (in_smem) near_regular_group(P, kv, der, n, pr.x, pr.y, it.y, nsl, xsw + (size_t)min(g, K - 1) * (3 + per), gl, W, valid, acc);
            else
                near_regular_group_global(P, kv, reinterpret_cast<const double2 *>(P.reg_derived) + (size_t)P.reg_doff[key] * PNB_DER2, n, pr.x,
                                          pr.y, it.y, nsl, xsw + (size_t)min(g, K - 1) * (3 + per), gl, W, valid, acc);
            // fixed tree over the W lanes of the group (W need not be a power of two)
            for (int off = 16; off > 0; off >>= 1) {
                if (off >= W) continue;
#pragma unroll
                for (int k = 0; k < NL; k++) {
                    const double v = __shfl_down_sync(0xffffffffu, acc[k], off);
                    if (gl + off < W) acc[k] += v;
                }
            }
            if (valid && gl == 0) {
#pragma unroll
                for (int k = 0; k < NL; k++) R[(size_t)item * NL + k] = acc[k];
            }
        } else {
            const int idx = warp;
            if (idx < c.z) {
                const int item = perm[c.y + idx];
                const int2 it = items[item];
                const int4 pr = pairs[it.x];
                const int Ka = pr.x, Kb = pr.y;
                const int Sl = near_slices(P, pr.z);
                // reference orientation of singular pairs: smaller cell index first
                const int c1 = min(Ka, Kb), c2 = max(Ka, Kb);
                int p1[3] = {0, 1, 2}, p2[3] = {0, 1, 2};
                const int pan = proto_panel(P.cells + (size_t)c1 * NV, NV, P.cells + (size_t)c2 * NV, NV, c1 == c2, p1, p2);
                double acc[NL];
#pragma unroll
                for (int k = 0; k < NL; k++) acc[k] = 0.;
                lanes_singular_interior<2>(P, c1, c2, pan, p1, p2, it.y * 32 + lane, 32 * Sl, acc);
                warp_allreduce<NA>(acc);
#pragma unroll
                for (int k = 0; k < NL; k++)
                    if (k == lane) R[(size_t)item * NL + k] = acc[k];
            }
        }
    }
}

// sums the slices of a near pair and maps the 21 values to the cross block and the two cell-diagonal blocks
// of (row cell Ka, column cell Kb).  Deliberately not inlined (keeps the registers of the unit kernel low).
__device__ __forceinline__ void near_fetch(const DProblem &P, const double *__restrict__ R, int4 pr, double *xy, double *dxy)
{
    constexpr int NV = 3, ND = 6, NL = PairDims<2>::NL, NRr = 2 * NV - 1;
    const int Ka = pr.x, Kb = pr.y, panel = pr.z;
    const int Sl = near_slices(P, panel);
    double acc[NL];
#pragma unroll
    for (int k = 0; k < NL; k++) acc[k] = 0.;
    for (int q = 0; q < Sl; q++) {
        const double *r = R + (size_t)(pr.w + q) * NL;
#pragma unroll
        for (int k = 0; k < NL; k++) acc[k] += r[k];
    }
#pragma unroll
    for (int k = 0; k < 9; k++) xy[k] = 0.;
#pragma unroll
    for (int k = 0; k < 12; k++) dxy[k] = 0.;
    if (panel >= 1) {
        const double sc = 2.0 * P.vol[Ka] * P.vol[Kb];
        int k = 0;
#pragma unroll
        for (int II = 0; II < 2 * NV; II++)
#pragma unroll
            for (int JJ = II; JJ < 2 * NV; JJ++) {
                const double v = acc[k] * sc;
                if (II < NV && JJ >= NV) xy[II * NV + (JJ - NV)] = v;
                else if (JJ < NV) dxy[tri_idx(NV, II, JJ)] = v;
                else dxy[ND + tri_idx(NV, II - NV, JJ - NV)] = v;
                k++;
            }
    } else {
        const bool swapped = Ka > Kb;
        const int c1 = swapped ? Kb : Ka, c2 = swapped ? Ka : Kb;
        int p1[3] = {0, 1, 2}, p2[3] = {0, 1, 2};
        const int pan = proto_panel(P.cells + (size_t)c1 * NV, NV, P.cells + (size_t)c2 * NV, NV, c1 == c2, p1, p2);
        const double sc = (c1 == c2 ? 1.0 : 2.0) * 4.0 * P.vol[c1] * P.vol[c2];
        const int common = -pan, rows = 2 * NV - common;
        int k = 0;
        for (int II = 0; II < NRr; II++)
            for (int JJ = II; JJ < NRr; JJ++) {
                if (JJ < rows) {
                    const double v = acc[k] * sc;
                    int i = II < NV ? p1[II] : NV + p2[II - NV + common];
                    int j = JJ < NV ? p1[JJ] : NV + p2[JJ - NV + common];
                    if (j < i) { const int t = i; i = j; j = t; }
                    // (i,j) in the reference's 2NV x 2NV local numbering of (c1,c2)
                    if (i < NV && j >= NV) xy[!swapped ? i * NV + (j - NV) : (j - NV) * NV + i] = v;
                    else {
                        const bool first = j < NV;   // block of c1
                        const int a = first ? i : i - NV, b = first ? j : j - NV;
                        dxy[((first != swapped) ? 0 : ND) + tri_idx(NV, a, b)] = v;
                    }
                }
                k++;
            }
    }
}

// one thread per near pair: slices summed and mapped once, so that the unit kernel only reads 21 finished values
// (cross block 9, row-cell block 6, column-cell block 6) per near pair
__global__ void __launch_bounds__(256) gnear_finalize_kernel(DProblem P, const int4 *__restrict__ pairs, int npairs,
                                                             const double *__restrict__ R, double *__restrict__ F)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= npairs) return;
    double xy[9], d12[12];
    near_fetch(P, R, pairs[q], xy, d12);
#pragma unroll
    for (int k = 0; k < 9; k++) F[(size_t)q * 21 + k] = xy[k];
#pragma unroll
    for (int k = 0; k < 12; k++) F[(size_t)q * 21 + 9 + k] = d12[k];
}

// -------------------------------------------------------------------------------------------------
// unit kernel: regular pairs of order 2..5 thread-per-pair (binned by order); every other pair was evaluated by
// gnear_eval_kernel and is fetched here, so that all contributions of a unit are added in one fixed order
// -------------------------------------------------------------------------------------------------
#define PNB_MW_MAX 12          // warps of the unit kernel (one CTA per SM)
struct GMixFixed {
    PowTabS pw;
    FarRule far[PNB_FAR_MAX_ORDER - 1];      // orders 2..PNB_FAR_MAX_ORDER
};

// per warp: staged cross blocks of the pairs of the current row cell (9 x cap), column-cell sums (6 x cap),
// classification result per column slot (cap ints) and the slots sorted by evaluation class (cap bytes)
__host__ __device__ inline size_t gmix_warp_bytes_dev(int cap)
{
    return (size_t)cap * (9 * 8 + 6 * 8 + 4) + (((size_t)cap + 15) & ~(size_t)15);
}
inline size_t gmix_warp_bytes(int cap)
{
    return (size_t)cap * (9 * 8 + 6 * 8 + 4) + (((size_t)cap + 15) & ~(size_t)15);
}

inline size_t gmix_shared_bytes(int cap, int maxld)
{
    size_t b = 0;
    auto add = [&](size_t x) { b += (x + 15) & ~(size_t)15; };
    add(sizeof(GMixFixed));
    add((size_t)6 * cap * 8);         // sx of the column side
    add((size_t)2 * cap * 8);         // cx
    add((size_t)cap * 8);             // vol
    add((size_t)2 * cap * 4);         // lh, ah
    add((size_t)2 * cap * 4);         // cell, loc
    add((size_t)(maxld + 2) * 2);     // incidence lists of the column dofs: pointers
    add((size_t)3 * cap * 2);         // ... entries
    return b;
}

// warps of the unit kernel that fit into the shared memory of an SM (0: none)
inline int gmix_warps(int cap, int maxld, size_t budget)
{
    const size_t sh = gmix_shared_bytes(cap, maxld) + 64;
    if (sh >= budget) return 0;
    return (int)std::min<size_t>(PNB_MW_MAX, (budget - sh) / gmix_warp_bytes(cap));
}

inline size_t gmix_smem_bytes(int cap, int maxld, int nw) { return gmix_shared_bytes(cap, maxld) + (size_t)nw * gmix_warp_bytes(cap); }

// doubles of global scratch per CTA: the unit block with one row per (row cell slot, local vertex), then dof-indexed
inline size_t gmix_scratch_doubles(int cap, int maxld, int ldS) { return (size_t)(3 * cap + maxld) * ldS; }

inline size_t gnear_list_smem_bytes(int cap)
{
    size_t b = 0;
    auto add = [&](size_t x) { b += (x + 15) & ~(size_t)15; };
    add((size_t)4 * cap * 8);
    add((size_t)4 * cap * 4);
    add((size_t)4 * cap * 4);
    return b;
}

// -------------------------------------------------------------------------------------------------
// unit kernel of all units that are not uniformly of order 2.  One CTA per SM, ONE WARP PER ROW CELL: the warp classifies
// the partners of its row cell in the column group (shared vertices only for adjacent groups; getQuadOrder in FP32 with
// exact FP64 re-evaluation near an integer), sorts them by evaluation class with ballots (order 5, 2, 3, 4, then the
// pairs that gnear_eval_kernel evaluated: touching pairs, higher orders), and evaluates them thread-per-pair in sorted
// order, so that the lanes of a warp run the same evaluator except in the one or two chunks where the class changes.
// There is no barrier inside a unit: rounds 1 and 2 binned 256 pairs per step over the whole CTA and waited at 4-5
// barriers per step for the warps that held the high-order pairs (ncu: barrier 3.0-3.7 of ~10 stall cycles per issue).
// Accumulation without atomics and in a fixed order:
//  * cross blocks: every pair stores its 3 x 3 block into the warp's staging array (a private place per column slot);
//    after the row the warp gathers them per column dof through the incidence lists of the column group (fixed order)
//    and writes the three rows (row cell, local vertex) x column dofs of the unit block, which lives in CTA-private
//    global scratch (L2 resident).  The rows of the dofs of the row group are gathered the same way before the block
//    goes to the matrix (or, several GPUs, to the staging buffers) in ticket order;
//  * row-cell diagonal blocks: per-lane registers over the row, fixed butterfly over the lanes;
//  * column-cell diagonal blocks: per-warp sums over the rows of the warp (rows are dealt out statically), summed over
//    the warps in warp order at the end of the unit.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PNB_MW_MAX * 32, 1)
gmix_kernel(DProblem P, GroupSched G, const GUnit *__restrict__ units, int first, int nunits, int *__restrict__ next,
            double *__restrict__ A, int64_t ld, int far_mask, double *__restrict__ scratch)
{
    constexpr int ND = 6, SB = PNB_SB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *sp = smem_raw;
    const int cap = G.cap, ldS = G.ldS;
    const int NT = blockDim.x, NW = NT >> 5;
    GMixFixed &sm = *reinterpret_cast<GMixFixed *>(carve(sp, sizeof(GMixFixed)));
    double *sxJ = reinterpret_cast<double *>(carve(sp, (size_t)6 * cap * 8));
    double *cxJ = reinterpret_cast<double *>(carve(sp, (size_t)2 * cap * 8));
    double *volJ = reinterpret_cast<double *>(carve(sp, (size_t)cap * 8));
    float *lhJ = reinterpret_cast<float *>(carve(sp, (size_t)2 * cap * 4));
    int *cellJ = reinterpret_cast<int *>(carve(sp, (size_t)2 * cap * 4));
    unsigned short *incptr = reinterpret_cast<unsigned short *>(carve(sp, (size_t)(G.maxld + 2) * 2));
    unsigned short *inc = reinterpret_cast<unsigned short *>(carve(sp, (size_t)3 * cap * 2));
    float *ahJ = lhJ + cap;
    int *locJ = cellJ + cap;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char *wbase = sp + (size_t)warp * gmix_warp_bytes_dev(cap);
    double *Sw = reinterpret_cast<double *>(wbase);                 // [9][cap]
    double *DYw = Sw + (size_t)9 * cap;                             // [6][cap]
    int *info = reinterpret_cast<int *>(DYw + (size_t)6 * cap);     // [cap]
    unsigned char *list = reinterpret_cast<unsigned char *>(info + cap);
    // the unit block of this CTA in global memory
    double *Sg = scratch + (size_t)blockIdx.x * ((size_t)(3 * cap + G.maxld) * ldS);
    double *Sd = Sg + (size_t)3 * cap * ldS;
    const float cf = (float)P.c_int, sf = (float)fmax(-0.5 * (P.sing + 2), 0.);

    __shared__ int s_ticket;
    {
        powtab_stage(&sm.pw, P.pow_int, tid, NT);
        const double *fs = reinterpret_cast<const double *>(P.far_rules + 2);
        double *fd = reinterpret_cast<double *>(&sm.far[0]);
        for (int e = tid; e < (int)((PNB_FAR_MAX_ORDER - 1) * sizeof(FarRule) / sizeof(double)); e += NT) fd[e] = fs[e];
    }
    unsigned long long my_pairs = 0, my_near = 0;
    __syncthreads();        // the power table is complete before its coefficients go to registers
    const PowCtxS kv(&sm.pw, lane);
    const unsigned lt = (1u << lane) - 1, half = lane < 16 ? 0x0000FFFFu : 0xFFFF0000u;
    for (;;) {
    __syncthreads();
    if (tid == 0) s_ticket = first + atomicAdd(next, 1);     // list positions [first, nunits) of this launch
    __syncthreads();
    const int ticket = s_ticket;
    if (ticket >= nunits) break;
    const GUnit u = units[ticket];
    const int I = u.I, J = u.J;
    const bool diag = I == J, nearunit = u.kind == 2;
    const int ibeg = G.gptr[I], nI = G.gptr[I + 1] - ibeg, jbeg = G.gptr[J], nJ = G.gptr[J + 1] - jbeg;
    const int dI = G.gdptr[I], nldI = G.gdptr[I + 1] - dI, dJ = G.gdptr[J], nldJ = G.gdptr[J + 1] - dJ;
    {
        for (int s = tid; s < nJ; s += NT) {
            const int cc = G.gcells[jbeg + s];
            cellJ[s] = cc;
            locJ[s] = G.gloc[jbeg + s];
            if (cc >= 0) {
                cxJ[s] = P.centers[(size_t)cc * 2];
                cxJ[cap + s] = P.centers[(size_t)cc * 2 + 1];
                lhJ[s] = P.lhf[cc];
                ahJ[s] = P.ahf[cc];
#pragma unroll
                for (int k = 0; k < 6; k++) sxJ[k * cap + s] = P.simplices[(size_t)cc * 6 + k];
                volJ[s] = P.vol[cc];
            }
        }
        const int i0 = G.gincptr[dJ];
        for (int b = tid; b <= nldJ; b += NT) incptr[b] = (unsigned short)(G.gincptr[dJ + b] - i0);
        const int ninc = G.gincptr[dJ + nldJ] - i0;
        for (int e = tid; e < ninc; e += NT) inc[e] = G.ginc[i0 + e];
        for (int e = lane; e < 6 * cap; e += 32) DYw[e] = 0.;
    }
    __syncthreads();
    for (int r1 = warp; r1 < nI; r1 += NW) {
        const int K1 = G.gcells[ibeg + r1];
        if (K1 < 0) continue;      // padding slot (warp uniform)
        const int l1 = G.gloc[ibeg + r1];
        const int rb = r1 & ~(SB - 1), k1 = r1 & (SB - 1);
        const bool nodof1 = (l1 & 0x00FFFFFF) == 0x00FFFFFF;
        double s1v[3][2];
#pragma unroll
        for (int m = 0; m < 3; m++) {
            s1v[m][0] = P.simplices[(size_t)K1 * 6 + 2 * m];
            s1v[m][1] = P.simplices[(size_t)K1 * 6 + 2 * m + 1];
        }
        const double c10 = P.centers[(size_t)K1 * 2], c11 = P.centers[(size_t)K1 * 2 + 1];
        const float lh1 = P.lhf[K1], ah1 = P.ahf[K1];
        const double vol1 = 2.0 * P.vol[K1];
        int v1[3] = {0, 0, 0};
        if (nearunit) {
#pragma unroll
            for (int m = 0; m < 3; m++) v1[m] = P.cells[(size_t)K1 * 3 + m];
        }
        const int lab1 = P.labels ? P.labels[K1] : 0;
        // ---- classify the partners: info[slot] = 0 no pair, 2..5 order of the thread-per-pair evaluator, 6 | pos << 3 near ----
        int c5 = 0, c2 = 0, c3 = 0, c4 = 0, cn = 0;
        for (int s0 = 0; s0 < nJ; s0 += 32) {
            const int s2 = s0 + lane;
            int code = 0, todo = 0;
            if (s2 < nJ) {
                const int K2 = cellJ[s2];
                const int cb = s2 & ~(SB - 1), k2 = s2 & (SB - 1);
                // diagonal units: every unordered pair once (batches rb <= cb; inside a batch k1 <= k2)
                bool live = K2 >= 0 && (K1 != K2 ? (!diag || rb < cb || (rb == cb && k1 < k2)) : diag);
                // the reference skips pairs of cells without any dof
                if (live && nodof1 && (locJ[s2] & 0x00FFFFFF) == 0x00FFFFFF) live = false;
                // piecewise variable kernels: pairs of other classes belong to another problem instance
                if (live && P.labels && !pnb_class_active(P, lab1, P.labels[K2])) live = false;
                if (live) {
                    int panel = 0;
                    if (K1 == K2) panel = -3;
                    else if (nearunit) {
                        int v2[3];
#pragma unroll
                        for (int m = 0; m < 3; m++) v2[m] = P.cells[(size_t)K2 * 3 + m];
                        panel = -shared_vertices(v1, 3, v2, 3);
                    }
                    if (panel == 0) {
                        const double a = c10 - cxJ[s2], b = c11 - cxJ[cap + s2];
                        panel = fast_order_2d(a * a + b * b, lh1, lhJ[s2], ah1, ahJ[s2], cf, sf);
                        if (panel < 0) {
                            // getPanelType evaluates (c1 <= c2): keep the operand order of the reference
                            const int ca = min(K1, K2), cbb = max(K1, K2);
                            const double d = center_distance(P.centers + (size_t)ca * 2, P.centers + (size_t)cbb * 2, 2);
                            panel = quad_order_interior(P, P.h[ca], P.h[cbb], d);
                        }
                    }
                    const bool is_far = panel >= 2 && panel <= PNB_FAR_MAX_ORDER && ((far_mask >> panel) & 1);
                    if (panel > P.max_order) atomicMax(G.err, panel);
                    else if (is_far) code = panel;
                    else if (nearunit) todo = panel;
                    else atomicMax(G.err + 1, 1);   // host bound violated (never expected)
                }
            }
            if (nearunit) {
                // position in the near pair list: sub-batch base + pairs of earlier rows of the sub-batch + rank in the row
                const unsigned nbal = __ballot_sync(0xffffffffu, todo != 0);
                if (todo != 0) {
                    const size_t sbi = ((size_t)u.slot * G.nbmax + rb / SB) * G.nbmax + (s2 / SB);
                    const int pos = G.nearbase[sbi] + G.nearrow[sbi * SB + k1] + __popc(nbal & half & lt);
                    const int4 pr = G.npairs[pos];
                    if (pr.x != K1 || pr.y != cellJ[s2] || pr.z != todo) atomicMax(G.err + 1, 2);
                    else code = 6 | (pos << 3);
                }
            }
            if (s2 < nJ) {
                info[s2] = code;
                if (code == 0) {
#pragma unroll
                    for (int k = 0; k < 9; k++) Sw[k * cap + s2] = 0.;
                }
            }
            const int kind = code & 7;
            c5 += __popc(__ballot_sync(0xffffffffu, kind == 5));
            c2 += __popc(__ballot_sync(0xffffffffu, kind == 2));
            c3 += __popc(__ballot_sync(0xffffffffu, kind == 3));
            c4 += __popc(__ballot_sync(0xffffffffu, kind == 4));
            cn += __popc(__ballot_sync(0xffffffffu, kind == 6));
        }
        // ---- slots sorted by class (counting sort, slot order inside a class) ----
        const int ntot = c5 + c2 + c3 + c4 + cn;
        {
            int b5 = 0, b2 = c5, b3 = b2 + c2, b4 = b3 + c3, bn = b4 + c4;
            for (int s0 = 0; s0 < nJ; s0 += 32) {
                const int s2 = s0 + lane;
                const int kind = s2 < nJ ? (info[s2] & 7) : 0;
                const unsigned m5 = __ballot_sync(0xffffffffu, kind == 5), m2 = __ballot_sync(0xffffffffu, kind == 2);
                const unsigned m3 = __ballot_sync(0xffffffffu, kind == 3), m4 = __ballot_sync(0xffffffffu, kind == 4);
                const unsigned mn = __ballot_sync(0xffffffffu, kind == 6);
                int pos = -1;
                if (kind == 5) pos = b5 + __popc(m5 & lt);
                else if (kind == 2) pos = b2 + __popc(m2 & lt);
                else if (kind == 3) pos = b3 + __popc(m3 & lt);
                else if (kind == 4) pos = b4 + __popc(m4 & lt);
                else if (kind == 6) pos = bn + __popc(mn & lt);
                if (pos >= 0) list[pos] = (unsigned char)s2;
                b5 += __popc(m5); b2 += __popc(m2); b3 += __popc(m3); b4 += __popc(m4); bn += __popc(mn);
            }
        }
        __syncwarp();
        // ---- evaluate in sorted order ----
        double xa[ND];
#pragma unroll
        for (int k = 0; k < ND; k++) xa[k] = 0.;
        for (int p = lane; p < ntot; p += 32) {
            const int slot = list[p];
            const int inf = info[slot], kind = inf & 7;
            double xy[9], xx[6], yy[6];
            if (kind == 6) {
                const double *f = G.F + (size_t)(inf >> 3) * 21;
#pragma unroll
                for (int k = 0; k < 9; k++) xy[k] = __ldg(f + k);
#pragma unroll
                for (int k = 0; k < 6; k++) { xx[k] = __ldg(f + 9 + k); yy[k] = __ldg(f + 15 + k); }
                my_near++;
            } else {
                double s2v[3][2];
#pragma unroll
                for (int m = 0; m < 3; m++) {
                    s2v[m][0] = sxJ[(2 * m) * cap + slot];
                    s2v[m][1] = sxJ[(2 * m + 1) * cap + slot];
                }
                const double sc = vol1 * volJ[slot];
                // node counts of the adopted rule family (far_expected_nodes): 3, 6, 6, 7
                if (kind == 2) far_eval_n<3>(sm.far[0], s1v, s2v, kv, xy, xx, yy);
                else if (kind == 5) far_eval_n<7>(sm.far[3], s1v, s2v, kv, xy, xx, yy);
                else far_eval_n<6>(sm.far[kind - 2], s1v, s2v, kv, xy, xx, yy);
#pragma unroll
                for (int k = 0; k < 9; k++) xy[k] *= sc;
#pragma unroll
                for (int k = 0; k < 6; k++) { xx[k] *= sc; yy[k] *= sc; }
                my_pairs++;
            }
#pragma unroll
            for (int k = 0; k < 9; k++) Sw[k * cap + slot] = xy[k];
#pragma unroll
            for (int k = 0; k < ND; k++) {
                xa[k] += xx[k];
                DYw[k * cap + slot] += yy[k];
            }
        }
        // ---- row-cell block: fixed butterfly over the lanes ----
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
            for (int k = 0; k < ND; k++) xa[k] += __shfl_xor_sync(0xffffffffu, xa[k], off);
        }
        if (lane < ND) {
            double v = xa[0];
#pragma unroll
            for (int k = 1; k < ND; k++) if (lane == k) v = xa[k];
            G.Dp[((size_t)J * P.nc + K1) * ND + lane] = v;
        }
        __syncwarp();
        // ---- cross blocks per column dof: rows (r1, local vertex) of the unit block ----
        for (int b = lane; b < nldJ; b += 32) {
            double o0 = 0., o1 = 0., o2 = 0.;
            const int e1 = incptr[b + 1];
            for (int e = incptr[b]; e < e1; e++) {
                const int v = inc[e], sl = v >> 2, j = v & 3;
                o0 += Sw[j * cap + sl];
                o1 += Sw[(3 + j) * cap + sl];
                o2 += Sw[(6 + j) * cap + sl];
            }
            double *dst = Sg + (size_t)(r1 * 3) * ldS + b;
            dst[0] = o0;
            dst[ldS] = o1;
            dst[2 * ldS] = o2;
        }
        __syncwarp();
    }
    __syncthreads();
    // column-cell sums over the warps (warp order); same group on both sides: slot Dp[I][c] takes the row sums
    // (written above) and these
    for (int e = tid; e < nJ * ND; e += NT) {
        const int s2 = e / ND, k = e - s2 * ND;
        const int cc = cellJ[s2];
        if (cc < 0) continue;
        double s = 0.;
        for (int w = 0; w < NW; w++)
            s += reinterpret_cast<const double *>(sp + (size_t)w * gmix_warp_bytes_dev(cap))[(size_t)(9 + k) * cap + s2];
        double *dp = &G.Dp[((size_t)I * P.nc + cc) * ND + k];
        *dp = diag ? *dp + s : s;
    }
    // rows of the dofs of the row group, gathered over the (row cell, local vertex) rows in a fixed order: one warp per
    // row a, lanes over the columns; the loads of up to eight incidences are issued together (L2 latency)
    const bool staged = G.dist.nparts > 0;
    if (!staged) g_wait_predecessors(G, 1, ticket, I, J, tid, NT);
    for (int a = warp; a < nldI; a += NW) {
        const int q0 = G.gincptr[dI + a], nq = G.gincptr[dI + a + 1] - q0;
        double *arow = staged ? nullptr : A + (size_t)G.gdofs[dI + a] * ld;
        for (int b0 = 0; b0 < nldJ; b0 += 64) {
            const int b = b0 + lane, b2 = b + 32;
            double s = 0., s2 = 0.;
            for (int qq = 0; qq < nq; qq += 8) {
                double v[8], w[8];
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const int inc = qq + t < nq ? G.ginc[q0 + qq + t] : -1;
                    const double *src = Sg + (size_t)((inc >> 2) * 3 + (inc & 3)) * ldS;
                    v[t] = (inc >= 0 && b < nldJ) ? __ldcg(src + b) : 0.;
                    w[t] = (inc >= 0 && b2 < nldJ) ? __ldcg(src + b2) : 0.;
                }
#pragma unroll
                for (int t = 0; t < 8; t++) { s += v[t]; s2 += w[t]; }
            }
            if (staged) {
                if (b < nldJ) Sd[a * ldS + b] = s;
                if (b2 < nldJ) Sd[a * ldS + b2] = s2;
            } else {
                // the matrix (L2 operations: other SMs update neighbouring entries of the same lines)
                double *d1 = b < nldJ ? arow + G.gdofs[dJ + b] : nullptr, *d2 = b2 < nldJ ? arow + G.gdofs[dJ + b2] : nullptr;
                const double o1 = d1 ? __ldcg(d1) : 0., o2 = d2 ? __ldcg(d2) : 0.;
                if (d1) __stcg(d1, o1 + s);
                if (d2) __stcg(d2, o2 + s2);
            }
        }
    }
    if (staged) {
        __syncthreads();
        g_flush_staged(G, G.dist.uoff_mix + (size_t)ticket * G.dist.nparts, Sd, ldS, I, dI, nldI, dJ, nldJ, tid, NT);
    } else g_signal_done(G, 1, ticket, tid);
    }
    for (int off = 16; off > 0; off >>= 1) {
        my_pairs += __shfl_xor_sync(0xffffffffu, my_pairs, off);
        my_near += __shfl_xor_sync(0xffffffffu, my_near, off);
    }
    if (lane == 0 && my_pairs) atomicAdd(G.counters, my_pairs);
    if (lane == 0 && my_near) atomicAdd(G.counters + 1, my_near);
}

// F = U + U^T in place, 32 x 32 tiles; bitwise symmetric by construction
__global__ void __launch_bounds__(256) symmetrize_kernel(double *A, int64_t ld, int N, int row_tile0)
{
    __shared__ double T1[32][33], T2[32][33];
    const int r = row_tile0 + blockIdx.y, c = blockIdx.x;
    if (r > c) return;
    const int r0 = r * 32, c0 = c * 32;
    for (int e = threadIdx.x; e < 32 * 32; e += 256) {
        const int a = e >> 5, b = e & 31;
        T1[a][b] = (r0 + a < N && c0 + b < N) ? A[(size_t)(r0 + a) * ld + c0 + b] : 0.;
        T2[a][b] = (c0 + a < N && r0 + b < N) ? A[(size_t)(c0 + a) * ld + r0 + b] : 0.;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * 32; e += 256) {
        const int a = e >> 5, b = e & 31;
        if (r0 + a < N && c0 + b < N) A[(size_t)(r0 + a) * ld + c0 + b] = T1[a][b] + T2[b][a];
        if (r != c && c0 + a < N && r0 + b < N) A[(size_t)(c0 + a) * ld + r0 + b] = T1[b][a] + T2[a][b];
    }
}

// D[c] = sum_g Dp[g][c] + Dbnd[c], fixed order
__global__ void greduce_D_kernel(GroupSched G, double *D, const double *Dbnd, int nc, int use_bnd)
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= (int64_t)nc * 6) return;
    double s = 0.;
    for (int g = 0; g < G.ngroups; g++) s += G.Dp[(size_t)g * nc * 6 + e];
    if (use_bnd) s += Dbnd[e];
    D[e] = s;
}

// -------------------------------------------------------------------------------------------------
// several GPUs: rows of the operator from the staged fragments.  One CTA per owned row a: for every group g that holds
// a and every group h, the row of unit (min(g,h), max(g,h)) that belongs to a (as a row of I and / or of J) is added
// to the columns dofs(h).  Groups of one colour share no dof, so the warps take different h of a colour without
// synchronisation; colours, and the groups g of a, follow each other in a fixed order: every entry is summed in the
// same order on every run (bitwise reproducible, no atomics).  The cell-diagonal blocks of the cells around a
// (D, summed over the parts by the caller) are added last, like scatter_D_kernel does.
// -------------------------------------------------------------------------------------------------
struct DistApply {
    int nrows;
    const int *rows;            // owned rows (global dofs), ascending
    const int *d2g_ptr;         // N+1: dof -> (group, local index in the group)
    const int2 *d2g;
    const int *colptr;          // ncolors+1: groups by colour
    const int *collist;
    const long long *uoff;      // ngroups x ngroups (I <= J): start of the unit's fragments in this part's staging, -1 = none
    const double *stage;        // this part's staging buffer
};

__global__ void __launch_bounds__(256) dist_apply_kernel(DProblem P, GroupSched G, DistApply X, const int *__restrict__ dof_ptr,
                                                         const int *__restrict__ dof_cells, const double *__restrict__ D, int use_D,
                                                         double *__restrict__ A, int64_t ld)
{
    const int r = blockIdx.x;
    if (r >= X.nrows) return;
    const int a = X.rows[r];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int me = G.dist.part, np = G.dist.nparts, ng = G.ngroups;
    double *row = A + (size_t)r * ld;
    for (int c = tid; c < P.N; c += blockDim.x) row[c] = 0.;
    __syncthreads();
    for (int col = 0; col < G.ncolors; col++) {
        const int h0 = X.colptr[col], nh = X.colptr[col + 1] - h0;
        for (int k = X.d2g_ptr[a]; k < X.d2g_ptr[a + 1]; k++) {
            const int g = X.d2g[k].x, la = X.d2g[k].y;
            const int nldg = G.gdptr[g + 1] - G.gdptr[g];
            const long long pos = G.dist.gpos[G.gdptr[g] + la];
            for (int hi = warp; hi < nh; hi += nwarps) {
                const int h = X.collist[h0 + hi];
                const int nldh = G.gdptr[h + 1] - G.gdptr[h];
                const int *cols = G.gdofs + G.gdptr[h];
                if (g <= h) {       // a is a row of I = g, columns dofs(J = h)
                    const long long off = X.uoff[(size_t)g * ng + h];
                    if (off >= 0) {
                        const double *f = X.stage + off + pos * nldh;
                        for (int c = lane; c < nldh; c += 32) row[cols[c]] += f[c];
                    }
                }
                if (g >= h) {       // a is a row of J = g of unit (I = h, J = g), columns dofs(I = h)
                    const long long off = X.uoff[(size_t)h * ng + g];
                    if (off >= 0) {
                        const double *f = X.stage + off + (long long)G.dist.gcnt[h * np + me] * nldg + pos * nldh;
                        for (int c = lane; c < nldh; c += 32) row[cols[c]] += f[c];
                    }
                }
            }
        }
        __syncthreads();
    }
    if (use_D && tid == 0) {
        for (int t = dof_ptr[a]; t < dof_ptr[a + 1]; t++) {
            const int K = dof_cells[t] >> 2, pq = dof_cells[t] & 3;
            for (int q = 0; q < 3; q++) {
                const int Jd = P.dofs[(size_t)K * 3 + q];
                if (Jd < 0) continue;
                const int kk = pq <= q ? tri_idx(3, pq, q) : tri_idx(3, q, pq);
                row[Jd] += D[(size_t)K * 6 + kk];
            }
        }
    }
}
