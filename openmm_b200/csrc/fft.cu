// fft.cu -- the bespoke grid-resident 3-D FFT for PME (sm_100a), fused with the reciprocal-space convolution.
//
// Replaces cufftExecR2C / cufftExecC2R + reciprocalConvolution + gridEvaluateEnergy of the reference CUDA
// platform (CudaKernels.cpp:826-829,1228-1255; pme.cc:390-505); numerically it restates fftpack_exec_3d +
// pme_reciprocal_convolution (ReferencePME.cpp:409-514, 793-799): unnormalised transforms in both directions,
// forward = exp(-2 pi i jk/n).
//
// Five launches for forward + convolution + inverse; every launch is a batch of independent 1-D lines that live in
// shared memory for the whole transform, so the grid is read and written exactly once per pass and stays in L2:
//   1  k_fft_z_fwd   real-to-complex along z; two real rows are packed into one complex line (half the work)
//   2  k_fft_y       complex along y, 16 adjacent kz columns per CTA (128-byte coalesced segments)
//   3  k_fft_x_conv  forward along x, multiply by the influence function, accumulate the reciprocal energy,
//                    inverse along x -- the k-space grid never leaves shared memory between the three
//   4  k_fft_y       inverse along y
//   5  k_fft_z_inv   complex-to-real along z (Hermitian unpacking, two rows per complex line)
// Round 1 v1 used one CTA per x-slab (3 launches); profiles/r01_launches_bench_nograph.csv showed 56 under-filled
// CTAs taking 27 us each -- the line-batched layout below gives 100-200 CTAs per pass.
//
// 1-D transforms are Stockham autosort, mixed radix with generic radices 2..16 (any n whose prime factors are
// <= 13: 56 = 8*7, 88 = 8*11, 90 = 6*5*3, 128 = 8*4*4 ...), out of place between two shared-memory buffers, twiddles
// staged in shared memory.  The whole reciprocal pipeline is DOUBLE precision (the forces are small differences of a
// smooth potential of magnitude ~1e2: an fp32 grid alone costs ~1e-3 kJ/mol/nm); the B200 FP64
// rate is ample for 1e5..1e6 grid points and these kernels are latency bound.
#include "engine.h"
#include <math.h>


// Thread layout of every FFT kernel: blockDim = (n, LY): threadIdx.x = position p inside a line, threadIdx.y strides
// over the lines of the batch -- no integer division in the hot loops.  Per stage, a packed table entry per position
// (built once per CTA) holds j (source butterfly offset), e (twiddle increment) and dst (Stockham output position).
#define TID (threadIdx.y*blockDim.x + threadIdx.x)
#define NTHR (blockDim.x*blockDim.y)

__device__ __forceinline__ void fft_build_tables(unsigned int* tab, const FftPlanDev& plan) {
    const int n = plan.n;
    for (int i = TID; i < plan.nstages*n; i += NTHR) {
        const int s = i/n, p = i - s*n;
        int Ns = 1;
        for (int a = 0; a < s; a++) Ns *= plan.radix[a];
        const int R = plan.radix[s], nb = n/R;
        const int j = p/R, q = p - j*R, k = j % Ns;
        const int e = (k*(n/(Ns*R)) + q*nb) % n;
        const int dst = (j/Ns)*Ns*R + k + q*Ns;
        tab[i] = (unsigned int) j | ((unsigned int) e << 10) | ((unsigned int) dst << 20);
    }
}

// One Stockham stage of radix R over `nlines` contiguous lines of length n (line l at base + l*n), ONE OUTPUT ELEMENT
// PER THREAD: output q of butterfly j is sum_t x[j + t n/R] * w^(t e), e = k twStep + q n/R -- the stage twiddle and
// the radix-R DFT root collapse into a single table lookup whose index advances by e (mod n).  The loop is rolled
// (same ~40 instructions of code for every radix) and every grid point is an independent work item.
__device__ __forceinline__ void fft_stage(const double2* __restrict__ in, double2* __restrict__ out, int n, int nlines, int R,
                                          const unsigned int* __restrict__ tab, const double2* __restrict__ tw, bool inverse) {
    const int p = threadIdx.x;
    if (p >= n) return;
    const unsigned int u = tab[p];
    const int j = u & 1023u, e = (u >> 10) & 1023u, dst = u >> 20;
    const int nb = n/R;
    const double sgn = inverse ? -1.0 : 1.0;
    for (int line = threadIdx.y; line < nlines; line += blockDim.y) {
        const double2* src = in + line*n + j;
        double2 acc = src[0];
        int idx = e;
#pragma unroll 4
        for (int t = 1; t < R; t++) {
            const double2 x = src[t*nb];
            const double2 wv = tw[idx];
            const double wy = sgn*wv.y;
            acc.x += x.x*wv.x - x.y*wy;
            acc.y += x.x*wy + x.y*wv.x;
            idx += e;
            if (idx >= n) idx -= n;
        }
        out[line*n + dst] = acc;
    }
}

// full 1-D transform of `nlines` contiguous lines; returns the buffer holding the result. Block-wide.
__device__ double2* fft_lines(double2* a, double2* b, const FftPlanDev& plan, int nlines, const unsigned int* tab, const double2* tw, bool inverse) {
    double2* in = a;
    double2* out = b;
    for (int s = 0; s < plan.nstages; s++) {
        fft_stage(in, out, plan.n, nlines, plan.radix[s], tab + s*plan.n, tw, inverse);
        __syncthreads();
        double2* t = in; in = out; out = t;
    }
    return in;
}

// factorisation into radices <= 16 minimising sum(R + 4): R complex MACs per point per generic stage plus a
// synchronisation cost per stage (exhaustive search, n is small)
static int best_cost(int n, int* radix, int depth) {
    if (n == 1) return 0;
    if (depth >= B200MD_MAX_FFT_STAGES) return 1 << 28;
    int best = 1 << 28, sub[B200MD_MAX_FFT_STAGES];
    for (int r = 2; r <= B200MD_MAX_RADIX && r <= n; r++) {
        if (n % r) continue;
        const int stageCost = r + 4;
        int c = stageCost + best_cost(n/r, sub, depth+1);
        if (c < best) {
            best = c;
            radix[0] = r;
            for (int k = 0; k + depth + 1 < B200MD_MAX_FFT_STAGES && k < B200MD_MAX_FFT_STAGES-1; k++) radix[k+1] = sub[k];
        }
    }
    return best;
}

bool fft_make_radices(int n, int* radix, int* nstages) {
    int r[B200MD_MAX_FFT_STAGES+1] = {0};
    if (n < 1) return false;
    if (best_cost(n, r, 0) >= (1 << 28)) return false;
    int ns = 0, rem = n;
    while (rem > 1 && ns < B200MD_MAX_FFT_STAGES) { radix[ns] = r[ns]; rem /= r[ns]; ns++; }
    if (rem != 1) return false;
    // larger radices first: the early stages have the poorest write locality, keep them few
    for (int a = 0; a < ns; a++) for (int b = a+1; b < ns; b++) if (radix[b] > radix[a]) { int t = radix[a]; radix[a] = radix[b]; radix[b] = t; }
    *nstages = ns;
    return true;
}

#define ZROWS 16          // real rows per CTA in the z passes (8 packed complex lines)
#define LINE_BATCH 16     // lines per CTA in the y and x passes
#define FFT_THREADS 256

size_t fft_plane_smem_bytes(int ny, int nz) {       // kept for the engine's capacity check: largest per-CTA need
    size_t z = (2*(size_t) (ZROWS/2)*nz + 3*nz)*sizeof(double2);
    size_t y = (2*(size_t) LINE_BATCH*ny + 3*ny)*sizeof(double2);
    return z > y ? z : y;
}
size_t fft_line_smem_bytes(int nx) { return (2*(size_t) LINE_BATCH*nx + 3*nx)*sizeof(double2); }

// twiddles (n double2) followed by the per-stage position tables (8*n uint32 = 2n double2 of space)
__device__ __forceinline__ unsigned int* stage_twiddles(double2* tws, const FftPlanDev& plan) {
    for (int i = TID; i < plan.n; i += NTHR) tws[i] = plan.tw[i];
    unsigned int* tab = (unsigned int*) (tws + plan.n);
    fft_build_tables(tab, plan);
    return tab;
}

// ---- 1: forward z, real to complex, two rows per complex line ----
__global__ void __launch_bounds__(1024) k_fft_z_fwd(PmeDev pme) {
    extern __shared__ double2 smem[];
    const int nz = pme.nz, nzc = pme.nzc;
    const int nrowsTotal = pme.nx*pme.ny;
    const int row0 = blockIdx.x*ZROWS;
    const int nrows = min(ZROWS, nrowsTotal - row0);
    const int np = (nrows + 1)/2;
    double2* A = smem;
    double2* B = A + (ZROWS/2)*nz;
    double2* tws = B + (ZROWS/2)*nz;
    const unsigned int* tab = stage_twiddles(tws, pme.plan[2]);
    if (pme.gridFixed != nullptr) {
        const long long* base = pme.gridFixed + (size_t) row0*nz;
        const double sc = 1.0/4294967296.0;
        for (int i = TID; i < np*nz; i += NTHR) {
            const int p = i/nz, z = i - p*nz;
            const double re = (double) base[(size_t) (2*p)*nz + z]*sc;
            const double im = (2*p+1 < nrows) ? (double) base[(size_t) (2*p+1)*nz + z]*sc : 0.0;
            A[i] = make_double2(re, im);
        }
    }
    else {
        const double* base = pme.grid + (size_t) row0*nz;
        for (int i = TID; i < np*nz; i += NTHR) {
            const int p = i/nz, z = i - p*nz;
            const double re = base[(size_t) (2*p)*nz + z];
            const double im = (2*p+1 < nrows) ? base[(size_t) (2*p+1)*nz + z] : 0.0;
            A[i] = make_double2(re, im);
        }
    }
    __syncthreads();
    const double2* R = fft_lines(A, B, pme.plan[2], np, tab, tws, false);
    // unpack the two interleaved real transforms straight to global memory
    double2* dst = pme.cgrid + (size_t) row0*nzc;
    for (int i = TID; i < np*nzc; i += NTHR) {
        const int p = i/nzc, k = i - p*nzc;
        const double2 Z = R[p*nz + k];
        double2 Zc = R[p*nz + ((nz - k) % nz)];
        Zc.y = -Zc.y;
        dst[(size_t) (2*p)*nzc + k] = make_double2(0.5*(Z.x + Zc.x), 0.5*(Z.y + Zc.y));
        if (2*p+1 < nrows) {
            const double2 d = make_double2(0.5*(Z.x - Zc.x), 0.5*(Z.y - Zc.y));
            dst[(size_t) (2*p+1)*nzc + k] = make_double2(d.y, -d.x);     // -i*d
        }
    }
}

// ---- 5: inverse z, complex to real ----
__global__ void __launch_bounds__(1024) k_fft_z_inv(PmeDev pme) {
    extern __shared__ double2 smem[];
    const int nz = pme.nz, nzc = pme.nzc;
    const int nrowsTotal = pme.nx*pme.ny;
    const int row0 = blockIdx.x*ZROWS;
    const int nrows = min(ZROWS, nrowsTotal - row0);
    const int np = (nrows + 1)/2;
    double2* A = smem;
    double2* B = A + (ZROWS/2)*nz;
    double2* tws = B + (ZROWS/2)*nz;
    const unsigned int* tab = stage_twiddles(tws, pme.plan[2]);
    const double2* src = pme.cgrid + (size_t) row0*nzc;
    // pack rows (2p, 2p+1) into one complex line using the Hermitian symmetry along z
    for (int i = TID; i < np*nz; i += NTHR) {
        const int p = i/nz, k = i - p*nz;
        const int kk = (k < nzc) ? k : nz - k;
        double2 a = src[(size_t) (2*p)*nzc + kk];
        double2 b = (2*p+1 < nrows) ? src[(size_t) (2*p+1)*nzc + kk] : make_double2(0.0, 0.0);
        if (k >= nzc) { a.y = -a.y; b.y = -b.y; }
        A[i] = make_double2(a.x - b.y, a.y + b.x);       // a + i b
    }
    __syncthreads();
    const double2* Z = fft_lines(A, B, pme.plan[2], np, tab, tws, true);
    double* dst = pme.grid + (size_t) row0*nz;
    for (int i = TID; i < np*nz; i += NTHR) {
        const int p = i/nz, z = i - p*nz;
        const double2 v = Z[i];
        dst[(size_t) (2*p)*nz + z] = v.x;
        if (2*p+1 < nrows) dst[(size_t) (2*p+1)*nz + z] = v.y;
    }
}

// ---- 2 / 4: along y, LINE_BATCH adjacent kz columns of one x per CTA ----
__global__ void __launch_bounds__(1024) k_fft_y(PmeDev pme, int inverse) {
    extern __shared__ double2 smem[];
    const int ny = pme.ny, nzc = pme.nzc;
    const int nbz = (nzc + LINE_BATCH - 1)/LINE_BATCH;
    const int x = blockIdx.x/nbz, bz = blockIdx.x - x*nbz;
    const int kz0 = bz*LINE_BATCH;
    const int nk = min(LINE_BATCH, nzc - kz0);
    double2* A = smem;
    double2* B = A + LINE_BATCH*ny;
    double2* tws = B + LINE_BATCH*ny;
    const unsigned int* tab = stage_twiddles(tws, pme.plan[1]);
    double2* base = pme.cgrid + (size_t) x*ny*nzc + kz0;
    for (int i = TID; i < ny*LINE_BATCH; i += NTHR) {
        const int y = i/LINE_BATCH, l = i - y*LINE_BATCH;
        if (l < nk) A[l*ny + y] = base[(size_t) y*nzc + l];
    }
    __syncthreads();
    const double2* R = fft_lines(A, B, pme.plan[1], nk, tab, tws, inverse != 0);
    for (int i = TID; i < ny*LINE_BATCH; i += NTHR) {
        const int y = i/LINE_BATCH, l = i - y*LINE_BATCH;
        if (l < nk) base[(size_t) y*nzc + l] = R[l*ny + y];
    }
}

// ---- 3: forward x, convolution + energy, inverse x; one batch of (ky,kz) lines per CTA ----
// mode 0: forward + convolution + inverse (PME); mode 1: forward only; mode 2: inverse only (stand-alone FFT)
template <bool ENERGY>
__global__ void __launch_bounds__(1024) k_fft_x_conv(PmeDev pme, double* energyOut, int mode) {
    extern __shared__ double2 smem[];
    const int nx = pme.nx;
    const int plane = pme.ny*pme.nzc;
    double2* A = smem;
    double2* B = A + LINE_BATCH*nx;
    double2* tws = B + LINE_BATCH*nx;
    const unsigned int* tab = stage_twiddles(tws, pme.plan[0]);
    const int m0 = blockIdx.x*LINE_BATCH;
    const int nl = min(LINE_BATCH, plane - m0);
    for (int i = TID; i < nx*LINE_BATCH; i += NTHR) {
        const int x = i/LINE_BATCH, l = i - x*LINE_BATCH;
        if (l < nl) A[l*nx + x] = pme.cgrid[(size_t) x*plane + m0 + l];
    }
    __syncthreads();
    double2* R = A;
    double2* other = B;
    if (mode != 2) {
        R = fft_lines(A, B, pme.plan[0], nl, tab, tws, false);
        other = (R == A) ? B : A;
    }
    if (mode == 0) {
        double esum = 0.0;
        for (int i = TID; i < nx*LINE_BATCH; i += NTHR) {
            const int x = i/LINE_BATCH, l = i - x*LINE_BATCH;
            if (l < nl) {
                const int m = m0 + l;
                const double et = pme.eterm[(size_t) x*plane + m];
                const double2 v = R[l*nx + x];
                if (ENERGY) {
                    const int kz = m % pme.nzc;
                    const double wgt = (kz == 0 || (2*kz == pme.nz)) ? 1.0 : 2.0;    // Hermitian mirror counted here
                    esum += wgt*et*(v.x*v.x + v.y*v.y);
                }
                R[l*nx + x] = make_double2(v.x*et, v.y*et);
            }
        }
        if (ENERGY) {
            __shared__ double red[32];
            for (int off = 16; off > 0; off >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, off);
            if ((TID & 31) == 0) red[TID >> 5] = esum;
            __syncthreads();
            if (TID == 0) {
                double tot = 0.0;
                for (int w = 0; w < (NTHR + 31)/32; w++) tot += red[w];
                atomicAdd(energyOut, 0.5*tot);
            }
        }
        __syncthreads();
    }
    if (mode != 1)
        R = fft_lines(R, other, pme.plan[0], nl, tab, tws, true);
    for (int i = TID; i < nx*LINE_BATCH; i += NTHR) {
        const int x = i/LINE_BATCH, l = i - x*LINE_BATCH;
        if (l < nl) pme.cgrid[(size_t) x*plane + m0 + l] = R[l*nx + x];
    }
}

static void set_smem(const void* f, size_t bytes) {
    if (bytes > 48*1024) cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes);
}

static dim3 fft_block(int n, int maxLines) {
    int ly = FFT_THREADS/n;
    if (ly < 1) ly = 1;
    if (ly > maxLines) ly = maxLines;
    return dim3(n, ly);
}

struct FftLaunch {
    size_t zs, ys, xs;
    int zb, yb, xb;
    dim3 zt, yt, xt;
    FftLaunch(const PmeDev& p) {
        zs = (2*(size_t) (ZROWS/2)*p.nz + 3*p.nz)*sizeof(double2);
        ys = (2*(size_t) LINE_BATCH*p.ny + 3*p.ny)*sizeof(double2);
        xs = (2*(size_t) LINE_BATCH*p.nx + 3*p.nx)*sizeof(double2);
        zt = fft_block(p.nz, ZROWS/2); yt = fft_block(p.ny, LINE_BATCH); xt = fft_block(p.nx, LINE_BATCH);
        zb = (p.nx*p.ny + ZROWS - 1)/ZROWS;
        yb = p.nx*((p.nzc + LINE_BATCH - 1)/LINE_BATCH);
        xb = (p.ny*p.nzc + LINE_BATCH - 1)/LINE_BATCH;
        set_smem((const void*) k_fft_z_fwd, zs);
        set_smem((const void*) k_fft_z_inv, zs);
        set_smem((const void*) k_fft_y, ys);
        set_smem((const void*) k_fft_x_conv<true>, xs);
        set_smem((const void*) k_fft_x_conv<false>, xs);
    }
};

void launch_pme_fft_conv(const NbDev& nb, const PmeDev& pme, bool energy, cudaStream_t s) {
    FftLaunch L(pme);
    k_fft_z_fwd<<<L.zb, L.zt, L.zs, s>>>(pme);
    k_fft_y<<<L.yb, L.yt, L.ys, s>>>(pme, 0);
    if (energy) k_fft_x_conv<true><<<L.xb, L.xt, L.xs, s>>>(pme, nb.energy + EN_RECIP, 0);
    else k_fft_x_conv<false><<<L.xb, L.xt, L.xs, s>>>(pme, nb.energy + EN_RECIP, 0);
    k_fft_y<<<L.yb, L.yt, L.ys, s>>>(pme, 1);
    k_fft_z_inv<<<L.zb, L.zt, L.zs, s>>>(pme);
}

void launch_fft3d_r2c(const PmeDev& pme, cudaStream_t s) {
    FftLaunch L(pme);
    k_fft_z_fwd<<<L.zb, L.zt, L.zs, s>>>(pme);
    k_fft_y<<<L.yb, L.yt, L.ys, s>>>(pme, 0);
    k_fft_x_conv<false><<<L.xb, L.xt, L.xs, s>>>(pme, nullptr, 1);
}

void launch_fft3d_c2r(const PmeDev& pme, cudaStream_t s) {
    FftLaunch L(pme);
    k_fft_x_conv<false><<<L.xb, L.xt, L.xs, s>>>(pme, nullptr, 2);
    k_fft_y<<<L.yb, L.yt, L.ys, s>>>(pme, 1);
    k_fft_z_inv<<<L.zb, L.zt, L.zs, s>>>(pme);
}
