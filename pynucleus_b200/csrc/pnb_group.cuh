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
//                2..5; for units that may hold other pairs it adds the blocks staged by gnear_kernel
//   gnear_kernel (runs first) singular pairs and regular pairs of order > 5 of the near units, warp per
//                (pair, slice); a unit is split into parts (row batches) that are staged separately
// Cell-diagonal blocks (xx / yy of nonlocalOperator_{SCALAR}.pxi:769-789) are staged per (partner group, cell):
// slot Dp[g][c] has exactly one writer, the unit (group(c), g).
#pragma once

struct GroupSched {
    int ngroups, cap, maxld, ldS, ncolors, nparts;
    const int *gptr;     // ngroups+1: first cell slot of a group (multiples of PNB_SB)
    const int *gcells;   // cell id per slot, -1 = padding; batches of PNB_SB slots share no vertex
    const int *gloc;     // packed group-local dof index of the 3 vertices (8 bits each, 0xFF = no dof)
    const int *gdptr;    // ngroups+1
    const int *gdofs;    // group-local dof -> global dof
    double *Dp;          // ngroups x nc x ND
    double *NS;          // near staging [slot][part][nsstride]: block (maxld x maxld), DX (cap x ND), DY (cap x ND)
    size_t nsstride;
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
};

__device__ __forceinline__ double f2_pow(const PowTab *t, const F2Rule &R, double d2)
{
    const int hi = __double2hiint(d2), lo = __double2loint(d2);
    const int E = ((hi >> 20) & 0x7ff) - 1023 + PNB_POW_EOFF;
    if ((unsigned)E > 255u) return kernel_value_slow(t->scal, t->expo, d2);
    const int idx = (hi >> 13) & 0x7f;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double2 it = t->IT[idx];
    const double r = fma(m, it.x, -1.0);
    double p = fma(R.c[7], r, R.c[6]);
    p = fma(p, r, R.c[5]);
    p = fma(p, r, R.c[4]);
    p = fma(p, r, R.c[3]);
    p = fma(p, r, R.c[2]);
    p = fma(p, r, R.c[1]);
    p = fma(p, r, R.c[0]);
    return t->T1[E] * (it.y * p);
}

__device__ __forceinline__ unsigned char *carve(unsigned char *&p, size_t bytes)
{
    unsigned char *r = p;
    p += (bytes + 15) & ~(size_t)15;
    return r;
}

// -------------------------------------------------------------------------------------------------
// uniform order 2
// -------------------------------------------------------------------------------------------------
inline size_t gf2_smem_bytes(int cap, int maxld, int ldS)
{
    size_t b = 0;
    auto add = [&](size_t x) { b += (x + 15) & ~(size_t)15; };
    add(sizeof(PowTab));
    add((size_t)maxld * ldS * 8);
    add((size_t)12 * cap * 8);        // nodes of both sides
    add((size_t)2 * cap * 8);         // vol
    add((size_t)4 * cap * 4);         // cell, loc (both sides)
    add((size_t)2 * 8 * 16 * 6 * 8);  // Yw
    add((size_t)cap * 6 * 8);         // DYs
    return b;
}

__global__ void __launch_bounds__(PNB_THREADS, 2)
gf2_kernel(DProblem P, GroupSched G, const GUnit *__restrict__ units, double *__restrict__ A, int64_t ld, F2Rule R)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *sp = smem_raw;
    const int cap = G.cap, ldS = G.ldS;
    PowTab *pw = reinterpret_cast<PowTab *>(carve(sp, sizeof(PowTab)));
    double *S = reinterpret_cast<double *>(carve(sp, (size_t)G.maxld * ldS * 8));
    double *xi = reinterpret_cast<double *>(carve(sp, (size_t)12 * cap * 8));
    double *yj = xi + 6 * cap;
    double *voli = reinterpret_cast<double *>(carve(sp, (size_t)2 * cap * 8));
    double *volj = voli + cap;
    int *celli = reinterpret_cast<int *>(carve(sp, (size_t)4 * cap * 4));
    int *cellj = celli + cap, *loci = celli + 2 * cap, *locj = celli + 3 * cap;
    double *Yw = reinterpret_cast<double *>(carve(sp, (size_t)2 * 8 * 16 * 6 * 8));
    double *DYs = reinterpret_cast<double *>(carve(sp, (size_t)cap * 6 * 8));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const GUnit u = units[blockIdx.x];
    const int I = u.I, J = u.J;
    const int ibeg = G.gptr[I], nI = G.gptr[I + 1] - ibeg, jbeg = G.gptr[J], nJ = G.gptr[J + 1] - jbeg;
    const int dI = G.gdptr[I], nldI = G.gdptr[I + 1] - dI, dJ = G.gdptr[J], nldJ = G.gdptr[J + 1] - dJ;
    {
        const double *src = reinterpret_cast<const double *>(P.pow_int);
        double *dst = reinterpret_cast<double *>(pw);
        for (int e = tid; e < (int)(sizeof(PowTab) / sizeof(double)); e += PNB_THREADS) dst[e] = src[e];
        for (int e = tid; e < nldI * ldS; e += PNB_THREADS) S[e] = 0.;
        for (int e = tid; e < nJ * 6; e += PNB_THREADS) DYs[e] = 0.;
        for (int e = tid; e < nI + nJ; e += PNB_THREADS) {
            const bool first = e < nI;
            const int s = first ? e : e - nI;
            const int c = G.gcells[(first ? ibeg : jbeg) + s];
            double *nd = first ? xi : yj;
            (first ? celli : cellj)[s] = c;
            (first ? loci : locj)[s] = G.gloc[(first ? ibeg : jbeg) + s];
            if (c >= 0) {
                const double *v = P.simplices + (size_t)c * 6;
                (first ? voli : volj)[s] = P.vol[c];
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    nd[(2 * q) * cap + s] = R.bary[0][q] * v[0] + R.bary[1][q] * v[2] + R.bary[2][q] * v[4];
                    nd[(2 * q + 1) * cap + s] = R.bary[0][q] * v[1] + R.bary[1][q] * v[3] + R.bary[2][q] * v[5];
                }
            }
        }
    }
    __syncthreads();
    const int k1 = tid >> 4, k2 = tid & 15;
    unsigned long long my_pairs = 0;
    int step = 0;
    for (int rb = 0; rb < nI; rb += PNB_SB) {
        const int s1 = rb + k1;
        const int c1 = celli[s1];
        const int l1 = loci[s1];
        double x[3][2];
#pragma unroll
        for (int q = 0; q < 3; q++) { x[q][0] = xi[(2 * q) * cap + s1]; x[q][1] = xi[(2 * q + 1) * cap + s1]; }
        const double v1 = c1 >= 0 ? 2.0 * voli[s1] : 0.;
        double xx[6] = {0., 0., 0., 0., 0., 0.};
        for (int cb = 0; cb < nJ; cb += PNB_SB, step++) {
            const int s2 = cb + k2;
            const int c2 = cellj[s2];
            const int l2 = locj[s2];
            double yy[6] = {0., 0., 0., 0., 0., 0.};
            // a pair is skipped only when neither cell carries a dof (as the reference does)
            const bool live = c1 >= 0 && c2 >= 0 && !((l1 & 0x00FFFFFF) == 0x00FFFFFF && (l2 & 0x00FFFFFF) == 0x00FFFFFF);
            if (live) {
                my_pairs++;
                double g[3][3];
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const double y0 = yj[(2 * j) * cap + s2], y1 = yj[(2 * j + 1) * cap + s2];
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        const double a = x[i][0] - y0, b = x[i][1] - y1;
                        g[i][j] = f2_pow(pw, R, a * a + b * b);
                    }
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
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        const double q = R.wphi[i][a];
                        X[a * 3 + 0] = fma(-q, t0, X[a * 3 + 0]);
                        X[a * 3 + 1] = fma(-q, t1, X[a * 3 + 1]);
                        X[a * 3 + 2] = fma(-q, t2, X[a * 3 + 2]);
                    }
                    r *= sc;
#pragma unroll
                    for (int e = 0; e < 6; e++) xx[e] = fma(R.qq[e][i], r, xx[e]);
                }
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    double c = 0.;
#pragma unroll
                    for (int i = 0; i < 3; i++) c = fma(g[i][j], R.w[i], c);
                    c *= sc;
#pragma unroll
                    for (int e = 0; e < 6; e++) yy[e] = fma(R.qq[e][j], c, yy[e]);
                }
                // conflict free: the 16 row cells share no vertex, neither do the 16 column cells
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const int ra = (l1 >> (8 * a)) & 0xFF;
                    if (ra == 0xFF) continue;
#pragma unroll
                    for (int b = 0; b < 3; b++) {
                        const int cbk = (l2 >> (8 * b)) & 0xFF;
                        if (cbk == 0xFF) continue;
                        S[ra * ldS + cbk] += X[a * 3 + b] * sc;
                    }
                }
            }
            // column-cell blocks: sum over the 16 row cells (two half-warps, then 8 warps through shared memory)
            double *yw = Yw + (size_t)(step & 1) * (8 * 16 * 6);
#pragma unroll
            for (int e = 0; e < 6; e++) yy[e] += __shfl_xor_sync(0xffffffffu, yy[e], 16);
            if (lane < 16) {
#pragma unroll
                for (int e = 0; e < 6; e++) yw[(warp * 16 + k2) * 6 + e] = yy[e];
            }
            __syncthreads();
            if (tid < 96) {
                const int kk2 = tid / 6, e = tid - kk2 * 6;
                double s = 0.;
#pragma unroll
                for (int w = 0; w < 8; w++) s += yw[(w * 16 + kk2) * 6 + e];
                DYs[(cb + kk2) * 6 + e] += s;
            }
        }
        // row-cell blocks: sum over the column cells of the whole group (16 lanes, fixed tree)
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) {
#pragma unroll
            for (int e = 0; e < 6; e++) xx[e] += __shfl_xor_sync(0xffffffffu, xx[e], off);
        }
        if (k2 == 0 && c1 >= 0) {
#pragma unroll
            for (int e = 0; e < 6; e++) G.Dp[((size_t)J * P.nc + c1) * 6 + e] = xx[e];
        }
    }
    __syncthreads();
    for (int e = tid; e < nldI * nldJ; e += PNB_THREADS) {
        const int a = e / nldJ, b = e - a * nldJ;
        A[(size_t)G.gdofs[dI + a] * ld + G.gdofs[dJ + b]] += S[a * ldS + b];
    }
    for (int e = tid; e < nJ * 6; e += PNB_THREADS) {
        const int c2 = cellj[e / 6];
        if (c2 >= 0) G.Dp[((size_t)I * P.nc + c2) * 6 + (e % 6)] = DYs[e];
    }
    for (int off = 16; off > 0; off >>= 1) my_pairs += __shfl_xor_sync(0xffffffffu, my_pairs, off);
    if (lane == 0 && my_pairs) atomicAdd(G.counters, my_pairs);
}

// -------------------------------------------------------------------------------------------------
// mixed units (orders 2..5 by thread-per-pair evaluation, binned by order) and near parts
// -------------------------------------------------------------------------------------------------
struct GMixFixed {
    PowTab pw;
    FarRule far[PNB_FAR_MAX_ORDER - 1];      // orders 2..PNB_FAR_MAX_ORDER
    double dxy[PNB_SB * PNB_SB][12];
    unsigned char slotD[PNB_SB * PNB_SB];
    int list[PNB_SB * PNB_SB];
    int clscnt[(PNB_FAR_MAX_ORDER - 1) * (PNB_THREADS / 32)];
    int warpcnt[PNB_THREADS / 32];
    int nlist, anyD;
};
struct GNearExtra {
    double partial[64][PairDims<2>::NL];     // slice sums of split pairs
    int listpanel[PNB_SB * PNB_SB];
};

inline size_t gmix_smem_bytes(int cap, int maxld, int ldS, bool nearpart)
{
    size_t b = 0;
    auto add = [&](size_t x) { b += (x + 15) & ~(size_t)15; };
    add(sizeof(GMixFixed));
    if (nearpart) add(sizeof(GNearExtra));
    add((size_t)maxld * ldS * 8);
    add((size_t)2 * 6 * cap * 8);     // sx
    add((size_t)2 * 2 * cap * 8);     // cx
    add((size_t)2 * cap * 8);         // vol
    add((size_t)2 * 2 * cap * 4);     // lh, ah
    add((size_t)2 * 2 * cap * 4);     // cell, loc
    add((size_t)cap * 6 * 8);         // DYs
    return b;
}

// NEARPART = false: one CTA per unit, far pairs; NEARPART = true: one CTA per (near unit, part), other pairs
template <bool NEARPART>
__global__ void __launch_bounds__(PNB_THREADS, NEARPART ? 1 : 2)
gmix_kernel(DProblem P, GroupSched G, const GUnit *__restrict__ units, double *__restrict__ A, int64_t ld, int far_mask)
{
    constexpr int NV = 3, NX = 9, ND = 6, NL = PairDims<2>::NL, SB = PNB_SB, NW = PNB_THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *sp = smem_raw;
    const int cap = G.cap, ldS = G.ldS;
    GMixFixed &sm = *reinterpret_cast<GMixFixed *>(carve(sp, sizeof(GMixFixed)));
    GNearExtra &nx = *reinterpret_cast<GNearExtra *>(NEARPART ? carve(sp, sizeof(GNearExtra)) : sp);
    double *S = reinterpret_cast<double *>(carve(sp, (size_t)G.maxld * ldS * 8));
    double *sx = reinterpret_cast<double *>(carve(sp, (size_t)2 * 6 * cap * 8));
    double *cxs = reinterpret_cast<double *>(carve(sp, (size_t)2 * 2 * cap * 8));
    double *vols = reinterpret_cast<double *>(carve(sp, (size_t)2 * cap * 8));
    float *lhs = reinterpret_cast<float *>(carve(sp, (size_t)2 * 2 * cap * 4));
    int *ints = reinterpret_cast<int *>(carve(sp, (size_t)2 * 2 * cap * 4));
    double *DYs = reinterpret_cast<double *>(carve(sp, (size_t)cap * 6 * 8));
    // side 0 = rows (I), side 1 = columns (J)
    double *sxI = sx, *sxJ = sx + 6 * cap, *cxI = cxs, *cxJ = cxs + 2 * cap, *volI = vols, *volJ = vols + cap;
    float *lhI = lhs, *ahI = lhs + cap, *lhJ = lhs + 2 * cap, *ahJ = lhs + 3 * cap;
    int *cellI = ints, *locI = ints + cap, *cellJ = ints + 2 * cap, *locJ = ints + 3 * cap;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const GUnit u = units[NEARPART ? blockIdx.x / G.nparts : blockIdx.x];
    const int part = NEARPART ? blockIdx.x % G.nparts : 0;
    const int I = u.I, J = u.J;
    const bool diag = I == J;
    const int ibeg = G.gptr[I], nI = G.gptr[I + 1] - ibeg, jbeg = G.gptr[J], nJ = G.gptr[J + 1] - jbeg;
    const int dI = G.gdptr[I], nldI = G.gdptr[I + 1] - dI, dJ = G.gdptr[J], nldJ = G.gdptr[J + 1] - dJ;
    const float cf = (float)P.c_int, sf = (float)fmax(-0.5 * (P.sing + 2), 0.);
    unsigned long long my_pairs = 0;
    {
        const double *src = reinterpret_cast<const double *>(P.pow_int);
        double *dst = reinterpret_cast<double *>(&sm.pw);
        for (int e = tid; e < (int)(sizeof(PowTab) / sizeof(double)); e += PNB_THREADS) dst[e] = src[e];
        if (!NEARPART) {
            const double *fs = reinterpret_cast<const double *>(P.far_rules + 2);
            double *fd = reinterpret_cast<double *>(&sm.far[0]);
            for (int e = tid; e < (int)((PNB_FAR_MAX_ORDER - 1) * sizeof(FarRule) / sizeof(double)); e += PNB_THREADS) fd[e] = fs[e];
        }
        for (int e = tid; e < nldI * ldS; e += PNB_THREADS) S[e] = 0.;
        for (int e = tid; e < cap * 6; e += PNB_THREADS) DYs[e] = 0.;
        for (int e = tid; e < nI + nJ; e += PNB_THREADS) {
            const bool first = e < nI;
            const int s = first ? e : e - nI;
            const int c = G.gcells[(first ? ibeg : jbeg) + s];
            (first ? cellI : cellJ)[s] = c;
            const int lc = G.gloc[(first ? ibeg : jbeg) + s];
            (first ? locI : locJ)[s] = lc;
            if (c >= 0) {
                double *d = first ? sxI : sxJ;
#pragma unroll
                for (int k = 0; k < 6; k++) d[k * cap + s] = P.simplices[(size_t)c * 6 + k];
                (first ? cxI : cxJ)[s] = P.centers[(size_t)c * 2];
                (first ? cxI : cxJ)[cap + s] = P.centers[(size_t)c * 2 + 1];
                (first ? volI : volJ)[s] = P.vol[c];
                (first ? lhI : lhJ)[s] = P.lhf[c];
                (first ? ahI : ahJ)[s] = P.ahf[c];
            }
        }
    }
    __syncthreads();
    const PowCtx kv(&sm.pw);
    const int k1 = tid / SB, k2 = tid % SB;

    const double *nsrc = (!NEARPART && u.kind == 2) ? G.NS + (size_t)u.slot * G.nparts * G.nsstride : nullptr;
    const size_t doff = (size_t)G.maxld * G.maxld;
    for (int rb = NEARPART ? part * SB : 0; rb < nI; rb += NEARPART ? G.nparts * SB : SB) {
        double dxacc = 0.;      // threads tid < SB*ND: entry (tid % ND) of the block of row cell rb + tid / ND
        for (int cb = diag ? rb : 0; cb < nJ; cb += SB) {
            // ---- classify every pair of the sub-batch ----
            const int s1 = rb + k1, s2 = cb + k2;
            const int K1 = cellI[s1], K2 = cellJ[s2];
            int todo = 0, cls = 0;
            sm.slotD[tid] = 0;
            if (tid == 0) sm.anyD = 0;
            // diagonal units: every unordered pair once (batches rb <= cb; inside a batch k1 <= k2)
            if (K1 >= 0 && K2 >= 0 && K1 != K2 ? (!diag || rb < cb || k1 < k2) : (K1 >= 0 && K1 == K2 && diag)) {
                if ((locI[s1] & 0x00FFFFFF) != 0x00FFFFFF || (locJ[s2] & 0x00FFFFFF) != 0x00FFFFFF) {
                    int panel;
                    if (K1 == K2) panel = -NV;
                    else {
                        panel = 0;
                        if (u.kind == 2) {
                            int v1[NV], v2[NV];
#pragma unroll
                            for (int m = 0; m < NV; m++) { v1[m] = P.cells[(size_t)K1 * NV + m]; v2[m] = P.cells[(size_t)K2 * NV + m]; }
                            panel = -shared_vertices(v1, NV, v2, NV);
                        }
                        if (panel == 0) {
                            const double a = cxI[s1] - cxJ[s2], b = cxI[cap + s1] - cxJ[cap + s2];
                            panel = fast_order_2d(a * a + b * b, lhI[s1], lhJ[s2], ahI[s1], ahJ[s2], cf, sf);
                            if (panel < 0) {
                                // getPanelType evaluates (c1 <= c2): keep the operand order of the reference
                                const int c1 = min(K1, K2), c2 = max(K1, K2);
                                const double d = center_distance(P.centers + (size_t)c1 * 2, P.centers + (size_t)c2 * 2, 2);
                                panel = quad_order_interior(P, P.h[c1], P.h[c2], d);
                            }
                        }
                    }
                    const bool is_far = panel >= 2 && panel <= PNB_FAR_MAX_ORDER && ((far_mask >> panel) & 1);
                    if (panel > P.max_order) atomicMax(G.err, panel);
                    else if (is_far) cls = NEARPART ? 0 : panel;
                    else {
                        todo = NEARPART ? panel : 0;
                        if (!NEARPART && u.kind != 2) atomicMax(G.err + 1, 1);   // host bound violated (never expected)
                    }
                }
            }
            // ---- ordered binning: far pairs by order, other pairs in slot order ----
            unsigned mybal = 0;
            if (!NEARPART) {
#pragma unroll
                for (int c = 2; c <= PNB_FAR_MAX_ORDER; c++) {
                    const unsigned bc = __ballot_sync(0xffffffffu, cls == c);
                    if (lane == 0) sm.clscnt[(c - 2) * NW + warp] = __popc(bc);
                    if (cls == c) mybal = bc;
                }
            } else {
                mybal = __ballot_sync(0xffffffffu, todo != 0);
                if (lane == 0) sm.warpcnt[warp] = __popc(mybal);
            }
            __syncthreads();    // B1
            if (!NEARPART) {
                const int me = (cls - 2) * NW + warp;
                int pos = 0, tot = 0;
#pragma unroll 4
                for (int q = 0; q < (PNB_FAR_MAX_ORDER - 1) * NW; q++) {
                    const int c = sm.clscnt[q];
                    if (q < me) pos += c;
                    tot += c;
                }
                if (cls != 0) sm.list[pos + __popc(mybal & ((1u << lane) - 1))] = tid | (cls << 12);
                if (tid == 0) sm.nlist = tot;
            } else {
                int pos = 0, tot = 0;
                for (int w = 0; w < NW; w++) {
                    if (w < warp) pos += sm.warpcnt[w];
                    tot += sm.warpcnt[w];
                }
                if (todo != 0) {
                    pos += __popc(mybal & ((1u << lane) - 1));
                    sm.list[pos] = tid;
                    nx.listpanel[pos] = todo;
                }
                if (tid == 0) sm.nlist = tot;
            }
            __syncthreads();    // B2
            const int nlist = sm.nlist;
            if (nlist == 0) continue;     // uniform across the CTA
            // ---- evaluate ----
            if (!NEARPART) {
                if (tid < nlist) {
                    const int item = sm.list[tid];
                    const int slot = item & 0xFF, order = item >> 12;
                    const int a1 = rb + slot / SB, a2 = cb + slot % SB;
                    my_pairs++;
                    double s1v[3][2], s2v[3][2], xx[6], yy[6], xy[9];
#pragma unroll
                    for (int m = 0; m < 3; m++) {
                        s1v[m][0] = sxI[(2 * m) * cap + a1];
                        s1v[m][1] = sxI[(2 * m + 1) * cap + a1];
                        s2v[m][0] = sxJ[(2 * m) * cap + a2];
                        s2v[m][1] = sxJ[(2 * m + 1) * cap + a2];
                    }
                    const double sc = 2.0 * volI[a1] * volJ[a2];
                    far_eval_2d(sm.far[order - 2], s1v, s2v, kv, true, xy, xx, yy);
#pragma unroll
                    for (int k = 0; k < 6; k++) {
                        sm.dxy[slot][k] = xx[k] * sc;
                        sm.dxy[slot][6 + k] = yy[k] * sc;
                    }
                    sm.slotD[slot] = 1;
                    sm.anyD = 1;
                    const int rl = locI[a1], cl = locJ[a2];
#pragma unroll
                    for (int i = 0; i < NV; i++) {
                        const int a = (rl >> (8 * i)) & 0xFF;
                        if (a == 0xFF) continue;
#pragma unroll
                        for (int j = 0; j < NV; j++) {
                            const int b = (cl >> (8 * j)) & 0xFF;
                            if (b == 0xFF) continue;
                            S[a * ldS + b] += xy[i * NV + j] * sc;
                        }
                    }
                }
            } else {
                // One warp per (pair, slice): sub-batches with few queued pairs split every pair into slices of its
                // quadrature nodes so that all warps stay busy; slice sums are combined in fixed order.
                const int Sl = nlist >= 32 ? 1 : (nlist >= 16 ? 2 : (nlist >= 8 ? 4 : 8));
                constexpr int NRr = 2 * NV - 1, NA = NRr * (NRr + 1) / 2;
                for (int pass = 0; pass < (Sl > 1 ? 2 : 1); pass++) {
                    if (pass == 1) __syncthreads();
                    const int nitems = pass == 0 ? nlist * Sl : nlist;
                    for (int it = warp; it < nitems; it += NW) {
                        const int q = pass == 0 ? it / Sl : it, sl = pass == 0 ? it - q * Sl : 0;
                        const int slot = sm.list[q] & 0xFF;
                        const int panel = nx.listpanel[q];
                        const int a1 = rb + slot / SB, a2 = cb + slot % SB;
                        const int Ka = cellI[a1], Kb = cellJ[a2];
                        // reference orientation of singular pairs: smaller cell index first
                        const bool swapped = panel < 0 && Ka > Kb;
                        const int c1 = swapped ? Kb : Ka, c2 = swapped ? Ka : Kb;
                        int p1[3] = {0, 1, 2}, p2[3] = {0, 1, 2};
                        int pan = panel;
                        if (panel < 0) pan = proto_panel(P.cells + (size_t)c1 * NV, NV, P.cells + (size_t)c2 * NV, NV, c1 == c2, p1, p2);
                        double acc[NL];
                        if (pass == 0) {
                            if (lane == 0 && sl == 0) my_pairs++;
                            if (panel >= 1) {
                                lanes_regular_interior<2>(P, Ka, Kb, panel, sl * 32 + lane, 32 * Sl, acc);
                                warp_allreduce<NL>(acc);
                            } else {
                                lanes_singular_interior<2>(P, c1, c2, pan, p1, p2, sl * 32 + lane, 32 * Sl, acc);
                                warp_allreduce<NA>(acc);
                            }
                            if (Sl > 1) {
#pragma unroll
                                for (int k = 0; k < NL; k++)
                                    if (k == lane) nx.partial[it][k] = acc[k];
                                continue;
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < NL; k++) {
                                double v = 0.;
                                for (int ss = 0; ss < Sl; ss++) v += nx.partial[q * Sl + ss][k];
                                acc[k] = v;
                            }
                        }
                        double myv = 0.;         // lane k < NX: entry k of the cross block
                        double myd = 0.;         // lane k < 2*ND: entry k of (dx, dy)
                        if (panel >= 1) {
                            const double sc = 2.0 * P.vol[Ka] * P.vol[Kb];
                            int k = 0;
#pragma unroll
                            for (int II = 0; II < 2 * NV; II++)
#pragma unroll
                                for (int JJ = II; JJ < 2 * NV; JJ++) {
                                    const double v = acc[k] * sc;
                                    if (II < NV && JJ >= NV) { if (lane == II * NV + (JJ - NV)) myv = v; }
                                    else if (JJ < NV) { if (lane == tri_idx(NV, II, JJ)) myd = v; }
                                    else { if (lane == ND + tri_idx(NV, II - NV, JJ - NV)) myd = v; }
                                    k++;
                                }
                        } else {
                            const double sc = (c1 == c2 ? 1.0 : 2.0) * 4.0 * P.vol[c1] * P.vol[c2];
                            const int common = -pan, rows = 2 * NV - common;
                            int k = 0;
#pragma unroll
                            for (int II = 0; II < NRr; II++)
#pragma unroll
                                for (int JJ = II; JJ < NRr; JJ++) {
                                    if (JJ < rows) {
                                        const double v = acc[k] * sc;
                                        int i = II < NV ? p1[II] : NV + p2[II - NV + common];
                                        int j = JJ < NV ? p1[JJ] : NV + p2[JJ - NV + common];
                                        if (j < i) { const int t = i; i = j; j = t; }
                                        // (i,j) in the reference's 2NV x 2NV local numbering of (c1,c2)
                                        if (i < NV && j >= NV) {
                                            const int e = !swapped ? i * NV + (j - NV) : (j - NV) * NV + i;
                                            if (lane == e) myv = v;
                                        } else {
                                            const bool first = j < NV;   // block of c1
                                            const int a = first ? i : i - NV, b = first ? j : j - NV;
                                            const bool to_dx = first != swapped;
                                            if (lane == (to_dx ? 0 : ND) + tri_idx(NV, a, b)) myd = v;
                                        }
                                    }
                                    k++;
                                }
                        }
                        // lanes 0..NX-1 add the cross block (distinct entries), lanes 0..2ND-1 store the diagonal blocks
                        if (lane < NX) {
                            const int i = lane / NV, j = lane - i * NV;
                            const int a = (locI[a1] >> (8 * i)) & 0xFF, b = (locJ[a2] >> (8 * j)) & 0xFF;
                            if (a != 0xFF && b != 0xFF) S[a * ldS + b] += myv;
                        }
                        if (lane < 2 * ND) sm.dxy[slot][lane] = myd;
                        if (lane == 0) { sm.slotD[slot] = 1; sm.anyD = 1; }
                    }
                }
            }
            __syncthreads();    // B3
            // ---- cell-diagonal blocks: reduce over the sub-batch ----
            if (sm.anyD) {
                if (tid < SB * ND) {
                    const int kk1 = tid / ND, comp = tid - kk1 * ND;
                    double sacc = 0.;
                    for (int kk2 = 0; kk2 < SB; kk2++)
                        if (sm.slotD[kk1 * SB + kk2]) sacc += sm.dxy[kk1 * SB + kk2][comp];
                    dxacc += sacc;
                } else if (tid < 2 * SB * ND) {
                    const int t2 = tid - SB * ND;
                    const int kk2 = t2 / ND, comp = t2 - kk2 * ND;
                    double sacc = 0.;
                    for (int kk1 = 0; kk1 < SB; kk1++)
                        if (sm.slotD[kk1 * SB + kk2]) sacc += sm.dxy[kk1 * SB + kk2][ND + comp];
                    DYs[(cb + kk2) * ND + comp] += sacc;
                }
            }
            __syncthreads();    // B4: slotD / dxy reused by the next sub-batch
        }
        if (tid < SB * ND) {
            const int kk1 = tid / ND, comp = tid - kk1 * ND;
            const int c = cellI[rb + kk1];
            if (NEARPART) {
                G.NS[((size_t)u.slot * G.nparts + part) * G.nsstride + doff + (size_t)(rb + kk1) * ND + comp] = dxacc;
            } else if (c >= 0) {
                // near units: the part that owns this row batch staged the sum of the remaining pairs
                if (nsrc) dxacc += nsrc[(size_t)((rb / SB) % G.nparts) * G.nsstride + doff + (size_t)(rb + kk1) * ND + comp];
                G.Dp[((size_t)J * P.nc + c) * ND + comp] = dxacc;
            }
        }
    }
    __syncthreads();
    if (NEARPART) {
        // stage the block and the column-cell sums of this part (row-cell sums were staged per row batch)
        double *dst = G.NS + ((size_t)u.slot * G.nparts + part) * G.nsstride;
        for (int e = tid; e < nldI * nldJ; e += PNB_THREADS) dst[e] = S[(e / nldJ) * ldS + (e % nldJ)];
        double *dd = dst + doff + (size_t)cap * ND;
        for (int e = tid; e < nJ * ND; e += PNB_THREADS) dd[e] = DYs[e];
    } else {
        for (int e = tid; e < nldI * nldJ; e += PNB_THREADS) {
            const int a = e / nldJ, b = e - a * nldJ;
            double v = S[a * ldS + b];
            if (nsrc)
                for (int pp = 0; pp < G.nparts; pp++) v += nsrc[(size_t)pp * G.nsstride + e];
            A[(size_t)G.gdofs[dI + a] * ld + G.gdofs[dJ + b]] += v;
        }
        // column-cell sums; same group on both sides: slot Dp[I][c] takes the row sums (written above) and these
        for (int e = tid; e < nJ * ND; e += PNB_THREADS) {
            const int c = cellJ[e / ND];
            if (c < 0) continue;
            double v = DYs[e];
            if (nsrc)
                for (int pp = 0; pp < G.nparts; pp++) v += nsrc[(size_t)pp * G.nsstride + doff + (size_t)cap * ND + e];
            double *dp = &G.Dp[((size_t)I * P.nc + c) * ND + (e % ND)];
            *dp = diag ? *dp + v : v;
        }
    }
    for (int off = 16; off > 0; off >>= 1) my_pairs += __shfl_xor_sync(0xffffffffu, my_pairs, off);
    if (lane == 0 && my_pairs) atomicAdd(G.counters + (NEARPART ? 1 : 0), my_pairs);
}

// F = U + U^T in place, 32 x 32 tiles; bitwise symmetric by construction
__global__ void __launch_bounds__(256) symmetrize_kernel(double *A, int64_t ld, int N)
{
    __shared__ double T1[32][33], T2[32][33];
    const int r = blockIdx.y, c = blockIdx.x;
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
