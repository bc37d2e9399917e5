// comm.cu -- the peer-memory data plane of the multi-GPU engine (sm_100a, NVLink 5 / NVSwitch).
//
// What the reference does here: CudaParallelKernels.cpp:110-121, 177-252 -- every device computes a share of the forces
// on ALL atoms, the host copies the partial force buffers to device 0 through pinned memory, device 0 sums and integrates
// everything, and the new positions are broadcast through the host again; reciprocal space stays on device 0
// (CudaKernels.cpp:715).  Here nothing passes through the host and no library collective is called on the step path:
// the kernels below (and k_integrate, k_pme_*, k_fft_* in their multi-rank form) store straight into the peers' windows
// over NVLink and publish one flag per stage (engine.h: CommDev, comm_wait, comm_signal).
//
//   k_force_push    partial int64 forces of the atoms this rank does not own -> the owners' inboxes (and zero them here)
//   k_integrate     (integrate.cu) owner: total = own partial + inboxes, integrate + constrain, new positions -> every rank
//   k_force_total   compute path only (energies, getState): owners total their atoms and broadcast the totals
//   k_grid_push     this rank's charge-grid contribution, slab by slab -> the slab owners' inboxes
//   k_vel_push      velocities of the owner's atoms -> every rank (before any read of the velocities from the host)
#include "engine.h"
#include <algorithm>
#include "../../include/b200md.h"

__device__ __forceinline__ long long* win_force(const CommDev& cd, int q) { return (long long*) (cd.peer[q] + cd.offForce); }
__device__ __forceinline__ long long* win_finbox(const CommDev& cd, int q, int src, int npad) { return (long long*) (cd.peer[q] + cd.offFinbox) + (size_t) src*3*npad; }

__global__ void __launch_bounds__(256) k_force_push(NbDev nb, CommDev cd) {
    const unsigned long long E = *cd.epoch + 1ull;
    const int stride = gridDim.x*blockDim.x;
    for (int a = blockIdx.x*blockDim.x + threadIdx.x; a < nb.natoms; a += stride) {
        if (a >= cd.atomLo[cd.rank] && a < cd.atomLo[cd.rank + 1]) continue;
        const int q = comm_owner_of_atom(cd, a);
        long long* in = win_finbox(cd, q, cd.rank, nb.npad);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const size_t i = (size_t) c*nb.npad + a;
            in[i] = nb.force[i];          // unconditional: the inbox needs no zeroing between steps
            nb.force[i] = 0;
        }
    }
    comm_signal(cd, CH_FORCE, E, gridDim.x);
}

// compute path: total force of the owner's atoms = own partial + inboxes, written to every rank's force buffer
__global__ void __launch_bounds__(256) k_force_total(NbDev nb, CommDev cd) {
    const unsigned long long E = *cd.epoch + 1ull;
    comm_wait(cd, CH_FORCE, E);
    const int lo = cd.atomLo[cd.rank], hi = cd.atomLo[cd.rank + 1];
    const int stride = gridDim.x*blockDim.x;
    for (int a = lo + blockIdx.x*blockDim.x + threadIdx.x; a < hi; a += stride) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const size_t i = (size_t) c*nb.npad + a;
            long long f = nb.force[i];
            for (int q = 0; q < cd.world; q++) if (q != cd.rank) f += win_finbox(cd, cd.rank, q, nb.npad)[i];
            for (int q = 0; q < cd.world; q++) win_force(cd, q)[i] = f;
        }
    }
    comm_signal(cd, CH_FINAL, E, gridDim.x);
}
// ... and everybody waits for everybody's totals; the exchange is complete, advance the epoch
__global__ void k_final_wait(CommDev cd, int ch) {
    const unsigned long long E = *cd.epoch + 1ull;
    comm_wait(cd, ch, E);
    if (threadIdx.x == 0) *cd.epoch = E;
}

__global__ void __launch_bounds__(256) k_vel_push(NbDev nb, CommDev cd) {
    const unsigned long long E = *cd.epoch + 1ull;
    const int lo = cd.atomLo[cd.rank], hi = cd.atomLo[cd.rank + 1];
    const int stride = gridDim.x*blockDim.x;
    for (int a = lo + blockIdx.x*blockDim.x + threadIdx.x; a < hi; a += stride) {
        const float4 v = nb.velm[a];
        for (int q = 0; q < cd.world; q++) if (q != cd.rank) ((float4*) (cd.peer[q] + cd.offVelm))[a] = v;
    }
    comm_signal(cd, CH_VEL, E, gridDim.x);
}

// wait until the peers' position stores of the last step have landed here (host reads of the replicated state)
__global__ void k_pos_wait(CommDev cd) {
    comm_wait(cd, CH_POS, *cd.posNeed);
}

// this rank's contribution to the charge grid (int64 fixed point, from the atoms it owns), slab by slab into the inbox
// that the slab's owner keeps for this rank.  blockIdx.y = peer (skipping this rank), 16-byte loads and stores.
__global__ void __launch_bounds__(256) k_grid_push(PmeDev pme, CommDev cd) {
    const unsigned long long E = *cd.epoch + 1ull;
    const size_t planeCells = (size_t) pme.ny*pme.nz;
    const size_t inboxStride = (size_t) cd.maxPlanes*planeCells;
    const int q = (int) blockIdx.y + ((int) blockIdx.y >= cd.rank ? 1 : 0);
    const size_t begin = (size_t) cd.xLo[q]*planeCells, cells = (size_t) (cd.xLo[q+1] - cd.xLo[q])*planeCells;
    const long long* src = pme.gridFixed + begin;
    long long* dst = (long long*) (cd.peer[q] + cd.offGridInbox) + (size_t) cd.rank*inboxStride;
    const size_t stride = (size_t) gridDim.x*blockDim.x;
    if ((begin & 1) == 0 && ((((size_t) cd.rank*inboxStride) & 1) == 0)) {      // 16-byte aligned on both sides
        const size_t pairs = cells >> 1;
        const ulonglong2* s2 = (const ulonglong2*) src;
        ulonglong2* d2 = (ulonglong2*) dst;
        for (size_t i = (size_t) blockIdx.x*blockDim.x + threadIdx.x; i < pairs; i += stride) d2[i] = s2[i];
        if ((cells & 1) && blockIdx.x == 0 && threadIdx.x == 0) dst[cells-1] = src[cells-1];
    }
    else
        for (size_t i = (size_t) blockIdx.x*blockDim.x + threadIdx.x; i < cells; i += stride) dst[i] = src[i];
    comm_signal(cd, CH_GRID, E, gridDim.x*gridDim.y);
}

static int sm_count() {
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    return sms;
}

void launch_force_push(const NbDev& nb, const CommDev& cd, cudaStream_t s) {
    if (cd.world <= 1) return;
    k_force_push<<<std::min((nb.natoms + 255)/256, 2*sm_count()), 256, 0, s>>>(nb, cd);
}
void launch_force_total(const NbDev& nb, const CommDev& cd, cudaStream_t s) {
    if (cd.world <= 1) return;
    const int own = cd.atomLo[cd.rank + 1] - cd.atomLo[cd.rank];
    k_force_total<<<std::max(1, std::min((own + 255)/256, 2*sm_count())), 256, 0, s>>>(nb, cd);
    k_final_wait<<<1, 32, 0, s>>>(cd, CH_FINAL);
}
void launch_vel_push(const NbDev& nb, const CommDev& cd, cudaStream_t s) {
    if (cd.world <= 1) return;
    const int own = cd.atomLo[cd.rank + 1] - cd.atomLo[cd.rank];
    k_vel_push<<<std::max(1, std::min((own + 255)/256, 2*sm_count())), 256, 0, s>>>(nb, cd);
    k_final_wait<<<1, 32, 0, s>>>(cd, CH_VEL);
}
void launch_pos_wait(const NbDev& nb, const CommDev& cd, cudaStream_t s) {
    (void) nb;
    if (cd.world <= 1) return;
    k_pos_wait<<<1, 32, 0, s>>>(cd);
}
void launch_grid_push(const PmeDev& pme, const CommDev& cd, cudaStream_t s) {
    if (cd.world <= 1) return;
    const size_t slab = (size_t) cd.maxPlanes*pme.ny*pme.nz/2;          // 16-byte elements per peer
    const int bx = (int) std::max<size_t>(1, std::min<size_t>((slab + 255)/256/4, (size_t) (4*sm_count()/(cd.world - 1) + 1)));
    k_grid_push<<<dim3(bx, cd.world - 1), 256, 0, s>>>(pme, cd);
}
