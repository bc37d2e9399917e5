// fft.cu -- the bespoke grid-resident 3-D FFT for PME (sm_100a), fused with the reciprocal-space convolution.
//
// Replaces cufftExecR2C / cufftExecC2R + reciprocalConvolution + gridEvaluateEnergy of the reference CUDA
// platform (CudaKernels.cpp:826-829,1228-1255; pme.cc:390-505); numerically it restates fftpack_exec_3d +
// pme_reciprocal_convolution (ReferencePME.cpp:409-514, 793-799): unnormalised transforms in both directions,
// forward = exp(-2 pi i jk/n).
//
// Slab decomposition, three launches for forward + convolution + inverse:
//   A  k_fft_zy_fwd  one CTA per x-slab: the (y,z) plane lives in shared memory, real-to-complex along z
//                    (two real rows packed in one complex line), complex along y, written once as [x][ky][kz]
//   B  k_fft_x_conv  one CTA per batch of (ky,kz) lines: forward along x, multiply by the influence function,
//                    accumulate the reciprocal energy, inverse along x -- the k-space grid never leaves smem
//   C  k_fft_yz_inv  one CTA per x-slab: inverse along y, complex-to-real along z
// Every grid point is read and written exactly once per launch: 3 x (8H + 8H) + the 4G real read / write.
//
// 1-D transforms are Stockham autosort, mixed radix with generic radices 2..16 (any n whose factors are <= 16,
// so 56 = 8*7, 88 = 8*11, 90 = 10*9, 128 = 16*8 ...), out of place between two shared-memory buffers.
#include "engine.h"
#include <math.h>

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x*b.x - a.y*b.y, a.x*b.y + a.y*b.x); }

// one Stockham stage of radix R over `nlines` lines of length n.
// element (line, i) lives at base[line*lineStride + i*elemStride].
template <int R>
__device__ void fft_stage(const float2* __restrict__ in, float2* __restrict__ out, int n, int nlines, int elemStride, int lineStride,
                          int Ns, const float2* __restrict__ tw, bool inverse) {
    const int nb = n/R;                 // butterflies per line
    const int twStep = n/(Ns*R);
    float2 root[R];
#pragma unroll
    for (int m = 0; m < R; m++) {
        float2 w = __ldg(&tw[m*nb]);
        root[m] = inverse ? make_float2(w.x, -w.y) : w;
    }
    const int total = nlines*nb;
    for (int w = threadIdx.x; w < total; w += blockDim.x) {
        const int line = w/nb;
        const int j = w - line*nb;
        const int k = j % Ns;
        const float2* src = in + line*lineStride;
        float2 v[R];
#pragma unroll
        for (int t = 0; t < R; t++) {
            float2 x = src[(j + t*nb)*elemStride];
            if (t > 0 && k > 0) {
                float2 wv = __ldg(&tw[t*k*twStep]);
                if (inverse) wv.y = -wv.y;
                x = cmul(x, wv);
            }
            v[t] = x;
        }
        float2* dst = out + line*lineStride;
        const int j0 = (j/Ns)*Ns*R + k;
#pragma unroll
        for (int q = 0; q < R; q++) {
            float2 acc = v[0];
#pragma unroll
            for (int t = 1; t < R; t++) {
                const float2 r = root[(q*t) % R];
                acc.x += v[t].x*r.x - v[t].y*r.y;
                acc.y += v[t].x*r.y + v[t].y*r.x;
            }
            dst[(j0 + q*Ns)*elemStride] = acc;
        }
    }
}

// full 1-D transform of all lines; returns the buffer holding the result. Block-wide; ends with __syncthreads.
__device__ float2* fft_lines(float2* a, float2* b, const FftPlanDev& plan, int nlines, int elemStride, int lineStride, bool inverse) {
    int Ns = 1;
    float2* in = a;
    float2* out = b;
    for (int s = 0; s < plan.nstages; s++) {
        const int R = plan.radix[s];
        switch (R) {
            case 2: fft_stage<2>(in, out, plan.n, nlines, elemStride, lineStride, Ns, plan.tw, inverse); break;
            case 3: fft_stage<3>(in, out, plan.n, nlines, elemStride, lineStride, Ns, plan.tw, inverse); break;
            case 4: fft_stage<4>(in, out, plan.n, nlines, elemStride, lineStride, Ns, plan.tw, inverse); break;
            case 5: fft_stage<5>(in, out, plan.n, nlines, elemStride, lineStride, Ns, plan.tw, inverse); break;
            case 6: fft_stage<6>(in, out, plan.n, nlines, elemStride, lineStride, Ns, plan.tw, inverse); break;
            case 7: fft_stage<7>(in, out, plan.n, nlines, elemStride, lineStride, Ns, plan.tw, inverse); break;
            case 8: fft_stage<8>(in, out, plan.n, nlines, elemStride, lineStride, Ns, plan.tw, inverse); break;
            case 9: fft_stage<9>(in, out, plan.n, nlines, elemStride, lineStride, Ns, plan.tw, inverse); break;
            case 10: fft_stage<10>(in, out, plan.n, nlines, elemStride, lineStride, Ns, plan.tw, inverse); break;
            case 11: fft_stage<11>(in, out, plan.n, nlines, elemStride, lineStride, Ns, plan.tw, inverse); break;
            case 12: fft_stage<12>(in, out, plan.n, nlines, elemStride, lineStride, Ns, plan.tw, inverse); break;
            case 13: fft_stage<13>(in, out, plan.n, nlines, elemStride, lineStride, Ns, plan.tw, inverse); break;
            case 14: fft_stage<14>(in, out, plan.n, nlines, elemStride, lineStride, Ns, plan.tw, inverse); break;
            case 15: fft_stage<15>(in, out, plan.n, nlines, elemStride, lineStride, Ns, plan.tw, inverse); break;
            case 16: fft_stage<16>(in, out, plan.n, nlines, elemStride, lineStride, Ns, plan.tw, inverse); break;
            default: break;   // n == 1
        }
        __syncthreads();
        Ns *= R;
        float2* t = in; in = out; out = t;
    }
    return in;
}

// greedy factorisation into radices <= 16, largest first so the stage count is minimal
bool fft_make_radices(int n, int* radix, int* nstages) {
    int ns = 0;
    int rem = n;
    while (rem > 1) {
        int best = 0;
        // prefer a split that leaves a remainder also factorable; greedy largest divisor <= 16 works for all
        // numbers whose prime factors are <= 13
        for (int r = B200MD_MAX_RADIX; r >= 2; r--)
            if (rem % r == 0) { best = r; break; }
        if (best == 0 || ns >= B200MD_MAX_FFT_STAGES) return false;
        radix[ns++] = best;
        rem /= best;
    }
    *nstages = ns;
    return true;
}

size_t fft_plane_smem_bytes(int ny, int nz) {
    int nzc = nz/2 + 1;
    int np = (ny + 1)/2;
    size_t elems = (size_t) ny*nzc;
    if ((size_t) np*nz > elems) elems = (size_t) np*nz;
    return 2*elems*sizeof(float2);
}

#define FFT_LINE_BATCH 16
size_t fft_line_smem_bytes(int nx) { return 2*(size_t) FFT_LINE_BATCH*nx*sizeof(float2); }

// ---- A: forward z (R2C, two rows per complex line) then y, one x-slab per CTA ----
__global__ void __launch_bounds__(512) k_fft_zy_fwd(PmeDev pme) {
    extern __shared__ float2 smem[];
    const int ny = pme.ny, nz = pme.nz, nzc = pme.nzc;
    const int np = (ny + 1)/2;
    size_t elems = (size_t) ny*nzc;
    if ((size_t) np*nz > elems) elems = (size_t) np*nz;
    float2* A = smem;
    float2* B = smem + elems;
    const int x = blockIdx.x;
    const float* plane = pme.grid + (size_t) x*ny*nz;
    for (int i = threadIdx.x; i < np*nz; i += blockDim.x) {
        int p = i/nz, z = i - p*nz;
        float re = plane[(2*p)*nz + z];
        float im = (2*p+1 < ny) ? plane[(2*p+1)*nz + z] : 0.f;
        A[i] = make_float2(re, im);
    }
    __syncthreads();
    float2* R = fft_lines(A, B, pme.plan[2], np, 1, nz, false);
    float2* O = (R == A) ? B : A;
    // unpack the two interleaved real transforms
    for (int i = threadIdx.x; i < np*nzc; i += blockDim.x) {
        int p = i/nzc, k = i - p*nzc;
        float2 Z = R[p*nz + k];
        float2 Zc = R[p*nz + ((nz - k) % nz)];
        Zc.y = -Zc.y;
        float2 a = make_float2(0.5f*(Z.x + Zc.x), 0.5f*(Z.y + Zc.y));
        float2 d = make_float2(0.5f*(Z.x - Zc.x), 0.5f*(Z.y - Zc.y));
        O[(2*p)*nzc + k] = a;
        if (2*p+1 < ny) O[(2*p+1)*nzc + k] = make_float2(d.y, -d.x);     // -i*d
    }
    __syncthreads();
    float2* other = (O == A) ? B : A;
    float2* Y = fft_lines(O, other, pme.plan[1], nzc, nzc, 1, false);
    float2* dst = pme.cgrid + (size_t) x*ny*nzc;
    for (int i = threadIdx.x; i < ny*nzc; i += blockDim.x) dst[i] = Y[i];
}

// ---- C: inverse y then z (C2R), one x-slab per CTA ----
__global__ void __launch_bounds__(512) k_fft_yz_inv(PmeDev pme) {
    extern __shared__ float2 smem[];
    const int ny = pme.ny, nz = pme.nz, nzc = pme.nzc;
    const int np = (ny + 1)/2;
    size_t elems = (size_t) ny*nzc;
    if ((size_t) np*nz > elems) elems = (size_t) np*nz;
    float2* A = smem;
    float2* B = smem + elems;
    const int x = blockIdx.x;
    const float2* src = pme.cgrid + (size_t) x*ny*nzc;
    for (int i = threadIdx.x; i < ny*nzc; i += blockDim.x) A[i] = src[i];
    __syncthreads();
    float2* Y = fft_lines(A, B, pme.plan[1], nzc, nzc, 1, true);
    float2* O = (Y == A) ? B : A;
    // pack rows (2p, 2p+1) into one complex line using the Hermitian symmetry along z
    for (int i = threadIdx.x; i < np*nz; i += blockDim.x) {
        int p = i/nz, k = i - p*nz;
        int kk = (k < nzc) ? k : nz - k;
        float2 a = Y[(2*p)*nzc + kk];
        float2 b = (2*p+1 < ny) ? Y[(2*p+1)*nzc + kk] : make_float2(0.f, 0.f);
        if (k >= nzc) { a.y = -a.y; b.y = -b.y; }
        O[i] = make_float2(a.x - b.y, a.y + b.x);       // a + i b
    }
    __syncthreads();
    float2* other = (O == A) ? B : A;
    float2* Z = fft_lines(O, other, pme.plan[2], np, 1, nz, true);
    float* plane = pme.grid + (size_t) x*ny*nz;
    for (int i = threadIdx.x; i < np*nz; i += blockDim.x) {
        int p = i/nz, z = i - p*nz;
        float2 v = Z[i];
        plane[(2*p)*nz + z] = v.x;
        if (2*p+1 < ny) plane[(2*p+1)*nz + z] = v.y;
    }
}

// ---- B: forward x, convolution + energy, inverse x; one batch of (ky,kz) lines per CTA ----
// mode 0: forward + convolution + inverse (PME); mode 1: forward only; mode 2: inverse only (stand-alone FFT)
template <bool ENERGY>
__global__ void __launch_bounds__(256) k_fft_x_conv(PmeDev pme, double* energyOut, int mode) {
    extern __shared__ float2 smem[];
    const int nx = pme.nx;
    const int plane = pme.ny*pme.nzc;
    float2* A = smem;
    float2* B = smem + FFT_LINE_BATCH*nx;
    const int m0 = blockIdx.x*FFT_LINE_BATCH;
    const int nl = min(FFT_LINE_BATCH, plane - m0);
    for (int i = threadIdx.x; i < nx*FFT_LINE_BATCH; i += blockDim.x) {
        int x = i/FFT_LINE_BATCH, l = i - x*FFT_LINE_BATCH;
        A[l*nx + x] = (l < nl) ? pme.cgrid[(size_t) x*plane + m0 + l] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    float2* R = A;
    float2* other = B;
    if (mode != 2) {
        R = fft_lines(A, B, pme.plan[0], FFT_LINE_BATCH, 1, nx, false);
        other = (R == A) ? B : A;
    }
    if (mode == 0) {
        float esum = 0.f;
        for (int i = threadIdx.x; i < nx*FFT_LINE_BATCH; i += blockDim.x) {
            int x = i/FFT_LINE_BATCH, l = i - x*FFT_LINE_BATCH;
            if (l < nl) {
                const int m = m0 + l;
                const float et = pme.eterm[(size_t) x*plane + m];
                float2 v = R[l*nx + x];
                if (ENERGY) {
                    const int kz = m % pme.nzc;
                    const float wgt = (kz == 0 || (2*kz == pme.nz)) ? 1.f : 2.f;    // Hermitian mirror counted here
                    esum += wgt*et*(v.x*v.x + v.y*v.y);
                }
                R[l*nx + x] = make_float2(v.x*et, v.y*et);
            }
        }
        if (ENERGY) {
            __shared__ float red[8];
            for (int off = 16; off > 0; off >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, off);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = esum;
            __syncthreads();
            if (threadIdx.x == 0) {
                float tot = 0.f;
                for (int w = 0; w < (blockDim.x >> 5); w++) tot += red[w];
                atomicAdd(energyOut, 0.5*(double) tot);
            }
        }
        __syncthreads();
    }
    if (mode != 1)
        R = fft_lines(R, other, pme.plan[0], FFT_LINE_BATCH, 1, nx, true);
    for (int i = threadIdx.x; i < nx*FFT_LINE_BATCH; i += blockDim.x) {
        int x = i/FFT_LINE_BATCH, l = i - x*FFT_LINE_BATCH;
        if (l < nl) pme.cgrid[(size_t) x*plane + m0 + l] = R[l*nx + x];
    }
}

static void set_smem(const void* f, size_t bytes) {
    if (bytes > 48*1024) cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes);
}

void launch_pme_fft_conv(const NbDev& nb, const PmeDev& pme, bool energy, cudaStream_t s) {
    size_t ps = fft_plane_smem_bytes(pme.ny, pme.nz);
    size_t ls = fft_line_smem_bytes(pme.nx);
    set_smem((const void*) k_fft_zy_fwd, ps);
    set_smem((const void*) k_fft_yz_inv, ps);
    set_smem((const void*) k_fft_x_conv<true>, ls);
    set_smem((const void*) k_fft_x_conv<false>, ls);
    int nbatch = (pme.ny*pme.nzc + FFT_LINE_BATCH - 1)/FFT_LINE_BATCH;
    k_fft_zy_fwd<<<pme.nx, 512, ps, s>>>(pme);
    if (energy) k_fft_x_conv<true><<<nbatch, 256, ls, s>>>(pme, nb.energy + EN_RECIP, 0);
    else k_fft_x_conv<false><<<nbatch, 256, ls, s>>>(pme, nb.energy + EN_RECIP, 0);
    k_fft_yz_inv<<<pme.nx, 512, ps, s>>>(pme);
}

void launch_fft3d_r2c(const PmeDev& pme, cudaStream_t s) {
    size_t ps = fft_plane_smem_bytes(pme.ny, pme.nz);
    size_t ls = fft_line_smem_bytes(pme.nx);
    set_smem((const void*) k_fft_zy_fwd, ps);
    set_smem((const void*) k_fft_x_conv<false>, ls);
    int nbatch = (pme.ny*pme.nzc + FFT_LINE_BATCH - 1)/FFT_LINE_BATCH;
    k_fft_zy_fwd<<<pme.nx, 512, ps, s>>>(pme);
    k_fft_x_conv<false><<<nbatch, 256, ls, s>>>(pme, nullptr, 1);
}

void launch_fft3d_c2r(const PmeDev& pme, cudaStream_t s) {
    size_t ps = fft_plane_smem_bytes(pme.ny, pme.nz);
    size_t ls = fft_line_smem_bytes(pme.nx);
    set_smem((const void*) k_fft_yz_inv, ps);
    set_smem((const void*) k_fft_x_conv<false>, ls);
    int nbatch = (pme.ny*pme.nzc + FFT_LINE_BATCH - 1)/FFT_LINE_BATCH;
    k_fft_x_conv<false><<<nbatch, 256, ls, s>>>(pme, nullptr, 2);
    k_fft_yz_inv<<<pme.nx, 512, ps, s>>>(pme);
}
