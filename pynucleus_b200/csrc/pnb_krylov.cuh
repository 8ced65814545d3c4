// BLAS-1 of the device CG loop, fused (cg_solver.solve, base/PyNucleus_base/solvers.pyx:364-445).  The scalars of the
// iteration (<p, Ap>, <r, z>, <r, r>) stay on the device; every reduction has a fixed shape (PNB_KRYLOV_BLOCKS partial sums
// per quantity, each over a fixed slice of the vector, then one fixed tree), so the iteration is bitwise reproducible.
//   krylov_dot_kernel        partial sums of <a, b>
//   krylov_cg_update_kernel  alpha = <r,z>_old / <p,Ap>;  x += alpha p;  r -= alpha Ap;  z = Minv r (or r);
//                            partial sums of <r, z> and <r, r>
//   krylov_finish_kernel     the fixed tree over the partial sums of up to two quantities
//   krylov_cg_direction_kernel   p = z + (<r,z> / <r,z>_old) p
#pragma once
#define PNB_KRYLOV_BLOCKS 256
#define PNB_KRYLOV_THREADS 256

__device__ __forceinline__ double krylov_block_sum(double v, double *red)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.;
    if (threadIdx.x == 0)
        for (int w = 0; w < PNB_KRYLOV_THREADS / 32; w++) s += red[w];
    return s;      // valid in thread 0
}

__global__ void __launch_bounds__(PNB_KRYLOV_THREADS) krylov_dot_kernel(int64_t n, const double *__restrict__ a, const double *__restrict__ b,
                                                                        double *__restrict__ partial)
{
    __shared__ double red[PNB_KRYLOV_THREADS / 32];
    const int64_t per = (n + PNB_KRYLOV_BLOCKS - 1) / PNB_KRYLOV_BLOCKS;
    const int64_t i0 = blockIdx.x * per, i1 = min(n, i0 + per);
    double s = 0.;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += PNB_KRYLOV_THREADS) s = fma(a[i], b[i], s);
    s = krylov_block_sum(s, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// out[q] = sum of partial[q * PNB_KRYLOV_BLOCKS ...], q < nq
__global__ void __launch_bounds__(PNB_KRYLOV_BLOCKS) krylov_finish_kernel(int nq, const double *__restrict__ partial, double *__restrict__ out)
{
    __shared__ double red[PNB_KRYLOV_BLOCKS / 32];
    for (int q = 0; q < nq; q++) {
        double v = partial[q * PNB_KRYLOV_BLOCKS + threadIdx.x];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.;
            for (int w = 0; w < PNB_KRYLOV_BLOCKS / 32; w++) s += red[w];
            out[q] = s;
        }
    }
}

// scal: [0] <r,z> of the previous iteration, [1] <p,Ap>; partial: 2 x PNB_KRYLOV_BLOCKS
__global__ void __launch_bounds__(PNB_KRYLOV_THREADS) krylov_cg_update_kernel(int64_t n, const double *__restrict__ scal, const double *__restrict__ p,
                                                                              const double *__restrict__ Ap, const double *__restrict__ Minv,
                                                                              double *__restrict__ x, double *r, double *z,
                                                                              double *__restrict__ partial)
{
    __shared__ double red[PNB_KRYLOV_THREADS / 32];
    const double alpha = scal[0] / scal[1];
    const int64_t per = (n + PNB_KRYLOV_BLOCKS - 1) / PNB_KRYLOV_BLOCKS;
    const int64_t i0 = blockIdx.x * per, i1 = min(n, i0 + per);
    double srz = 0., srr = 0.;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += PNB_KRYLOV_THREADS) {
        x[i] = fma(alpha, p[i], x[i]);
        const double ri = fma(-alpha, Ap[i], r[i]);
        r[i] = ri;
        const double zi = Minv ? Minv[i] * ri : ri;
        if (z != r) z[i] = zi;
        srz = fma(ri, zi, srz);
        srr = fma(ri, ri, srr);
    }
    srz = krylov_block_sum(srz, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = srz;
    srr = krylov_block_sum(srr, red);
    if (threadIdx.x == 0) partial[PNB_KRYLOV_BLOCKS + blockIdx.x] = srr;
}

// scal: [0] <r,z> old, [2] <r,z> new
__global__ void __launch_bounds__(PNB_KRYLOV_THREADS) krylov_cg_direction_kernel(int64_t n, const double *__restrict__ scal,
                                                                                 const double *__restrict__ z, double *__restrict__ p)
{
    const double beta = scal[2] / scal[0];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = fma(beta, p[i], z[i]);
}

// ---- C ABI ------------------------------------------------------------------------------------------------------------
// workspace: 3 scalars + 2 x PNB_KRYLOV_BLOCKS partial sums, device doubles
extern "C" int pnb_krylov_workspace_doubles(void) { return 8 + 2 * PNB_KRYLOV_BLOCKS; }

// out (device) = <a, b>
extern "C" int pnb_krylov_dot(int device, int64_t n, const double *a, const double *b, double *work, double *out, void *stream)
{
    if (!a || !b || !work || !out) return fail(PNB_ERR_ARG, "null argument");
    ON_DEVICE(device);
    cudaStream_t st = (cudaStream_t)stream;
    krylov_dot_kernel<<<PNB_KRYLOV_BLOCKS, PNB_KRYLOV_THREADS, 0, st>>>(n, a, b, work + 8);
    krylov_finish_kernel<<<1, PNB_KRYLOV_BLOCKS, 0, st>>>(1, work + 8, out);
    CK(cudaGetLastError());
    return 0;
}

// One CG update with the scalars in work[0..3): work[0] = <r,z> of the previous iteration (input), work[1] = <p,Ap>
// (computed here), then x, r, z are updated and work[2] = <r,z>, work[3] = <r,r> (outputs).  Minv: diagonal preconditioner
// or NULL (then z may be r itself).
extern "C" int pnb_krylov_cg_update(int device, int64_t n, const double *p, const double *Ap, const double *Minv, double *x, double *r,
                                    double *z, double *work, void *stream)
{
    if (!p || !Ap || !x || !r || !z || !work) return fail(PNB_ERR_ARG, "null argument");
    ON_DEVICE(device);
    cudaStream_t st = (cudaStream_t)stream;
    krylov_dot_kernel<<<PNB_KRYLOV_BLOCKS, PNB_KRYLOV_THREADS, 0, st>>>(n, p, Ap, work + 8);
    krylov_finish_kernel<<<1, PNB_KRYLOV_BLOCKS, 0, st>>>(1, work + 8, work + 1);
    krylov_cg_update_kernel<<<PNB_KRYLOV_BLOCKS, PNB_KRYLOV_THREADS, 0, st>>>(n, work, p, Ap, Minv, x, r, z, work + 8);
    krylov_finish_kernel<<<1, PNB_KRYLOV_BLOCKS, 0, st>>>(2, work + 8, work + 2);
    CK(cudaGetLastError());
    return 0;
}

// p = z + (work[2] / work[0]) p, then work[0] = work[2] (the new <r,z> becomes the old one)
__global__ void krylov_shift_kernel(double *work) { work[0] = work[2]; }

extern "C" int pnb_krylov_cg_direction(int device, int64_t n, const double *z, double *p, double *work, void *stream)
{
    if (!z || !p || !work) return fail(PNB_ERR_ARG, "null argument");
    ON_DEVICE(device);
    cudaStream_t st = (cudaStream_t)stream;
    int sms = device_attr(cudaDevAttrMultiProcessorCount, device);
    if (sms <= 0) sms = 148;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)sms * 4));
    krylov_cg_direction_kernel<<<blocks, 256, 0, st>>>(n, work, z, p);
    krylov_shift_kernel<<<1, 1, 0, st>>>(work);
    CK(cudaGetLastError());
    return 0;
}
