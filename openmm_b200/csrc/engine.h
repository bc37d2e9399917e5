// engine.h -- internal declarations shared by the CUDA translation units of libb200md.so.
// Public boundary: include/b200md.h.  Data layout and kernel inventory: DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#define B200MD_TILE 32
#define B200MD_PME_ORDER 5           // ReferenceLJCoulombIxn.cpp:243 (pme_init(..., 5, 1)); same on every reference platform
#define B200MD_FORCE_SCALE 4294967296.0   // 2^32 fixed point, as the reference GPU platforms (nonbonded.cu:301-316)
#define B200MD_ONE_4PI_EPS0 138.93545764438198    // SimTKOpenMMRealType.h:89
#define B200MD_BOLTZ 0.00831446261815324  // kJ/mol/K, SimTKOpenMMRealType.h:76-80 (CODATA 2018)
#define B200MD_MAX_RADIX 16
#define B200MD_MAX_FFT_STAGES 8

#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw std::runtime_error(std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

// Periodic box, reduced lower-triangular form (ContextImpl.cpp:267-275): a=(ax,0,0) b=(bx,by,0) c=(cx,cy,cz).
struct BoxDev {
    float ax, bx, by, cx, cy, cz;
    float invAx, invBy, invCz;
    int triclinic;
    int periodic;
    double dax, dby, dcz;   // box lengths in double (orthorhombic single-image wrap of the pair kernel)
    double recip[9];     // recipBoxVectors[i][j] at recip[3*i+j], as invert_box_vectors (ReferencePME.cpp:196-204)
    double volume;
};

// precision of the spectral pipeline (charge grid -> FFT -> convolution -> potential grid); the input grid is int64
// fixed point either way.  fp32 is sufficient: a double pipeline (make dbl) changes the reciprocal-space force error by
// nothing on DHFR, ApoA1 and the 894-ion fixture.  What DID matter is the precision of the B-spline weights in spreading
// and interpolation, which are double (pme.cu; profiles/r02_parity_probe.md).
#ifdef B200MD_REAL_DOUBLE      // experiment build (make dbl -> libb200md_dbl.so): double spectral pipeline, for error attribution only
typedef double real;
typedef double2 real2;
#else
typedef float real;
typedef float2 real2;
#endif

struct FftPlanDev {      // 1-D mixed-radix Stockham plan for one grid dimension
    int n;
    int nstages;
    int radix[B200MD_MAX_FFT_STAGES];
    const real2* tw;     // tw[k] = exp(-2 pi i k / n), k < n
};

// Everything the force kernels need, passed by value.
// One neighbour list: the sorted copy of the atoms, the 32-atom blocks and the 32x32 tiles.
struct ListDev {
    float4* sposq;               // sorted positions (exact user coordinates, refreshed every step) + charge
    float2* ssigeps;
    float4* swrap;               // sorted positions wrapped into the anchored cell at the build (list build only)
    int* sorig;                  // sorted slot -> user atom (-1 for padding)
    float4* blockCenter;
    float4* blockHalf;
    float4* superCenter;         // bounding boxes of superblocks of 32 consecutive blocks (first level of the candidate search)
    float4* superHalf;
    int* tileI;                  // [maxTiles]
    int* tileJ;                  // [maxTiles*32]
    int* tileMask;               // [maxTiles] index into maskPool or -1
    unsigned int* maskPool;      // [maxTiles*32]
    int* lc;                     // per-list counters, LC_* below
};
#define TILE_REGIONS 32
// ListDev::lc: tile and mask-tile counts of the TILE_REGIONS slot pools, max block half extent (float bits), tiles in use
enum { LC_TILES = 0, LC_MASKS = TILE_REGIONS, LC_MAXHALF = 2*TILE_REGIONS, LC_USED = 2*TILE_REGIONS + 1, LC_STRIDE = 128 };

// indices into NbDev::counters
enum { CT_PAIRSTART = 0,      // SM partition: CTAs of the tile kernel that have started (see k_pair)
       CT_REBUILD = 2, CT_OVERFLOW = 3, CT_BUILDS = 4, CT_PAIRS = 5, CT_LASTBLOCK = 8, CT_CUR = 10, CT_SOFT = 11, CT_PENDING = 12, CT_STALE = 13, CT_CURSOR = 14,
       CT_BTDONE = 7, CT_BAR = 9, CT_BAREXIT = 15 };      // k_build_tiles completion count; grid barrier of k_list_prep

struct NbDev {
    BoxDev box;
    int natoms, npad, nblocks;
    int method;                  // B200MD_NB_*
    int useSwitch;
    float cutoff, cutoff2, paddedCutoff2, switchDist;
    float alpha;                 // Ewald alpha
    float krf, crf;              // reaction field
    float dispDummy;
    // user-order state
    float4* posq;                // xyz + charge*sqrt(ONE_4PI_EPS0)
    float4* velm;                // v + 1/mass
    float2* sigeps;              // (sigma/2, 2 sqrt(eps))  (ReferenceKernels.cpp:1093-1097)
    const double* chargeD;       // the same parameters in double (user order): the double-precision close-pair path
    const double2* sigepsD;
    long long* force;            // [3][npad] fixed point, user order
    double* energy;              // [B200MD_NUM_ENERGY] accumulators
    // sorted (nonbonded) copies, blocks and tiles: two complete lists.  counters[CT_CUR] names the one the tile kernel reads;
    // a rebuild always fills the other one and flips (at once when the current list is no longer valid, at the end of the
    // step when the successor was built beside the step: see k_check_gather)
    ListDev list[2];
    int* sortedOf;               // user atom -> sorted slot (of the list built last; list construction only)
    float4* refPos;              // user-order positions at the last list build
    // binning
    int ncell[3];
    int ncells;
    const int* cellRank;         // linear cell -> rank along the space filling curve
    int* cellCount;              // [ncells+1] -> after scan: start offsets
    int* cellFill;               // [ncells]
    int* atomCell;               // [natoms] rank of the atom's cell
    int* tmpSorted;              // [npad]
    float4* atomShift;           // [natoms]
    int maxTiles;
    int* counters;               // [2]=rebuild flag [3]=overflow [4]=list builds [5]=pairs(diag) [6]=nan flag
    // exclusions, CSR in user order
    const int* exclStart;
    const int* exclList;
    // whole molecules that have diffused more than one box length out of the primary cell are moved back by lattice vectors
    // at list builds (fp32 coordinates lose ~1e-7 nm of resolution per nm of magnitude: a water walks ~100 nm per
    // microsecond); cellOffset remembers the lattice vectors so that the user keeps seeing a continuous trajectory
    int nmol;                    // 0: wrapping off (non-periodic, or more than one rank)
    const int* molStart;         // CSR over molAtoms
    const int* molAtoms;
    int* cellOffset;             // [3][npad] lattice vector counts to ADD when reporting positions
    float halfPad2;              // (padding/2)^2: beyond this displacement the current list is invalid
    float softPad2;              // displacement^2 at which the successor list is built beside the step (3e38: never)
    // multi-GPU sharding of the tile list / PME atoms
    int rank, world;
    // origin of the primary periodic cell used for binning (chosen at set_positions so that a structure centred anywhere,
    // e.g. a PDB centred on 0, is binned WITHOUT lattice shifts: a shift costs one fp32 rounding of the coordinate)
    double origin[3];
    // CUDA-graph conditional node that holds the list-rebuild kernels (0 = none: rebuild kernels are gated on counters[CT_REBUILD])
    unsigned long long condHandle;
    unsigned long long condAsync;   // second IF node: build of the successor list on a side stream
    int packCull;                // k_build_tiles: exact cull on full warps (B200MD_BT_PACK)
    int pairDynamic;             // tile kernel fetches tiles from a cursor instead of a static stride
    // SM partition: the tile kernel's CTAs that land on an SM whose bit is set here return at once, so those SMs stay free for
    // the reciprocal-space chain (spread -> FFT -> gather) that runs beside it; the other CTAs (one persistent wave) share ALL
    // tiles through the cursor.  All zero: no partition.
    unsigned long long pmeSmMask[4];
    float closeCut2;             // pairs closer than this (squared) are evaluated in double from the exact coordinates (0: off)
    int useRational;             // B200MD_PAIR_RATIONAL=1: rational Ewald kernel in the force-only tile loop (1 MUFU less, lower accuracy)
};

enum { EN_NB = 0, EN_RECIP = 1, EN_BOND = 2, EN_ANGLE = 3, EN_TORSION = 4, EN_EXC = 5, EN_KE = 6, B200MD_NUM_ENERGY = 8 };

struct PmeDev {
    int nx, ny, nz, nzc;
    real* grid;                  // real [nx][ny][nz] (output of the inverse transform, input of the gather)
    long long* gridFixed;        // real [nx][ny][nz], 2^32 fixed point: deterministic charge spreading (pme.cc:78-89 option)
    real2* cgrid;                // complex [nx][ny][nzc]
    real* eterm;                 // [nx][ny][nzc] influence function (no ONE_4PI_EPS0: charges carry sqrt of it)
    const double* moduli[3];
    FftPlanDev plan[3];          // x, y, z
    double alpha;
};

struct BondedDev {
    int nbonds, nangles, ntorsions, nexc;
    const int2* bondAtoms; const double2* bondParams;           // (r0, k)
    const int4* angleAtoms; const double2* angleParams;         // (theta0, k)
    const int4* torsionAtoms; const double4* torsionParams;     // (k, phase, n, 0)
    const int2* excAtoms; const double4* excParams;             // (qq14*ONE_4PI_EPS0, sigma, 4 eps, 0)
    int excPeriodic;
    // force group of every bond / angle / torsion (several Force objects of one class may sit in different groups,
    // ContextImpl::calcForcesAndEnergy groups, ContextImpl.cpp:293-308); an element is evaluated iff bit `group` of groupMask is set
    const unsigned char* bondGroup; const unsigned char* angleGroup; const unsigned char* torsionGroup;
    unsigned int groupMask;
};

// One integration unit = a SETTLE water, a SHAKE cluster (centre + <=3 H) or a free atom.
struct UnitDev {
    int nunits;
    const int4* unitAtoms;       // atoms (unused = -1); SETTLE: (O,H1,H2,-1)
    const int* unitType;         // 0 free, 1 SETTLE, 2 SHAKE
    const float4* unitParams;    // SETTLE: (dOH, dHH, 0, 0); SHAKE: (d1, d2, d3, 0)
};

struct IntegDev {
    int kind;
    float dt, vscale, fscale, noisescale;   // Langevin constants (ReferenceStochasticDynamics.cpp:94-99)
    float kT;
    float tol;
    unsigned int seed;
    unsigned long long stepIndex;            // not used on device when graphs are active: see stepCounter
    unsigned long long* stepCounter;         // device counter, incremented by the integrate kernel
    int fused;                               // step path: zero forces, fused CM removal, last block advances the counter
    int cmEveryStep;                         // CMMotionRemover with frequency 1
    double* cmScratch;                       // [3][4] rotating momentum/mass accumulators
    unsigned int* blocksDone;
};

// ---------------------------------------------------------------------------------------------------------------
// Multi-GPU data plane (one process per GPU): every rank maps every other rank's WINDOW (one cudaMalloc, exported with
// cudaIpcGetMemHandle) and the kernels themselves move the data with plain stores over NVLink, tile by tile, and
// publish a flag per (channel, source rank) when a stage is complete; consumers spin on their LOCAL flag copy.  No
// NCCL call on the step path.  Ownership: rank q owns the user atoms [atomLo[q], atomLo[q+1]) (cut at integration-unit
// boundaries): it reduces their forces, integrates them and pushes their new positions to everybody.  Reciprocal space
// is slab-decomposed: rank q owns the x planes [xLo[q], xLo[q+1]) for the (y,z) transforms and the (ky,kz) lines
// [q*lineChunk, ...) for the x transform; the two transposes are the stores of the FFT kernels themselves.
#define B200MD_MAX_RANKS 8
enum { CH_POS = 0,      // new positions of the owner's atoms are in everybody's posq        (k_integrate)
       CH_FORCE = 1,    // partial forces are in the owners' inboxes                         (k_force_push)
       CH_FINAL = 2,    // total forces of the owner's atoms are in everybody's force buffer (k_force_total; compute path only)
       CH_GRID = 3,     // charge-grid contributions are in the slab owners' inboxes         (k_grid_push)
       CH_FWD = 4,      // (y,z)-transformed planes are in the line owners' buffers          (k_fft_slab_fwd)
       CH_INV = 5,      // x-transformed, convolved lines are back in the slab owners' buffers (k_fft_x_conv)
       CH_POT = 6,      // potential planes are in everybody's grid                          (k_fft_slab_inv)
       CH_VEL = 7,      // velocities of the owner's atoms are in everybody's velm           (k_vel_push; state reads only)
       CH_COUNT = 8 };
struct CommDev {
    int rank, world;                     // world == 1: no peer traffic, every wait/signal is skipped
    int atomLo[B200MD_MAX_RANKS + 1];
    int unitLo[B200MD_MAX_RANKS + 1];
    int xLo[B200MD_MAX_RANKS + 1];       // x planes of the PME grid
    int lineChunk;                       // (ky,kz) lines per rank (multiple of the x-pass batch; the last rank may get fewer)
    int maxPlanes;                       // max x planes per rank
    char* peer[B200MD_MAX_RANKS];        // window base of every rank as mapped HERE (peer[rank] = own window)
    // offsets inside a window (identical on every rank)
    size_t offFlags, offPosq, offVelm, offForce, offFinbox, offCm, offGridInbox, offLineBuf, offPlaneBuf, offGrid;
    unsigned long long* epoch;           // local: number of completed exchanges; every evaluation uses E = *epoch + 1
    unsigned long long* posNeed;         // local: CH_POS value the next evaluation must wait for (0: positions were set by the host)
    unsigned int* done;                  // local: [CH_COUNT] block-completion counters of the signalling kernels
    int* errFlag;                        // local: sticky error (NbDev::counters + CT_OVERFLOW): a wait that times out raises 3
    int posByPush;                       // 1: k_integrate writes positions locally only, k_pos_push (TMA) publishes them and CH_POS
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long* comm_flag(const CommDev& cd, int onRank, int ch, int src) {
    return (unsigned long long*) (cd.peer[onRank] + cd.offFlags) + ch*B200MD_MAX_RANKS + src;
}
// Block-wide: wait until every peer has published `need` on channel ch.  Bounded (~1 s): a lost peer raises the sticky
// error flag instead of hanging the device.
__device__ __forceinline__ void comm_wait(const CommDev& cd, int ch, unsigned long long need) {
    if (cd.world > 1 && need != 0ull) {
        if (threadIdx.x < cd.world && (int) threadIdx.x != cd.rank && threadIdx.y == 0 && threadIdx.z == 0) {
            const unsigned long long* f = comm_flag(cd, cd.rank, ch, threadIdx.x);
            long spins = 0;
            while (ld_acquire_sys(f) < need) {
                if (*((volatile int*) cd.errFlag) == 3) break;        // a peer was lost earlier: fail fast, the host raises at its next sync
                __nanosleep(64);
                if (++spins > (1L << 22)) { *cd.errFlag = 3; break; }
            }
        }
    }
    __syncthreads();
}
// Block-wide: this block's stores to peer memory are complete.  Returns true (to all threads) in the LAST block of the grid,
// which then publishes the stage with comm_publish (possibly after a little more work of its own).
__device__ __forceinline__ bool comm_arrive(const CommDev& cd, int ch, unsigned int nblocks) {
    __shared__ int lastBlock;
    __syncthreads();                   // the block's stores happen-before thread 0's fence (cumulativity): ONE system fence per CTA
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
        __threadfence_system();
        lastBlock = (atomicAdd(&cd.done[ch], 1u) == nblocks - 1u);
        if (lastBlock) { cd.done[ch] = 0u; __threadfence_system(); }
    }
    __syncthreads();
    return lastBlock != 0;
}
__device__ __forceinline__ void comm_publish(const CommDev& cd, int ch, unsigned long long value) {       // one thread
    __threadfence_system();
    for (int k = 1; k < cd.world; k++) { const int q = (cd.rank + k) % cd.world; st_release_sys(comm_flag(cd, q, ch, cd.rank), value); }
}
__device__ __forceinline__ bool comm_signal(const CommDev& cd, int ch, unsigned long long value, unsigned int nblocks) {
    const bool last = comm_arrive(cd, ch, nblocks);
    if (last && threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) comm_publish(cd, ch, value);
    return last;
}
__device__ __forceinline__ int comm_owner_of_atom(const CommDev& cd, int a) {
    int q = 0;
#pragma unroll
    for (int k = 1; k < B200MD_MAX_RANKS; k++) q += (k < cd.world && a >= cd.atomLo[k]) ? 1 : 0;
    return q;
}
#endif

// 2^32 fixed point <-> fp32 without the 64-bit conversion instructions (I2F.S64 / F2I.S64 are multi-pass on the XU pipe and
// showed up as the hottest instructions of the FFT load loop and of the integrator in the round-1 profiles)
#ifdef __CUDACC__
__device__ __forceinline__ float fixed_to_float(long long v) {
    const int hi = (int) (v >> 32);
    const unsigned int lo = (unsigned int) v;
    return __int2float_rn(hi) + __uint2float_rn(lo)*2.3283064365386963e-10f;
}
__device__ __forceinline__ long long float_to_fixed(float f) {           // |f| < 2^31
    const float fl = floorf(f);
    const int hi = __float2int_rd(f);
    const unsigned int lo = __float2uint_rz((f - fl)*4294967296.0f);
    return ((long long) hi << 32) | (long long) lo;
}
#endif

// General constraint networks (CCMA, constraints.cu): one CTA per connected component of the constraint graph.
struct CcmaDev {
    int ncomp, ncon, natomsC;
    const int* compConStart;     // [ncomp+1] constraints of a component are contiguous
    const int* compAtomStart;    // [ncomp+1] so are its atoms (positions in `atoms`)
    const int2* conAtoms;        // [ncon] user atom indices
    const float* conDist;        // [ncon]
    const float* conRedMass;     // [ncon] 0.5/(1/mi + 1/mj)
    const int* rowStart; const int* col; const float* val;      // approximate inverse of the coupling matrix, CSR over constraints
    const int* atoms;            // [natomsC] user atom index
    const int* aStart;           // [natomsC+1] constraints of an atom: +(k+1) if it is the first atom of constraint k, -(k+1) if the second
    const int* aCon;
    float4* rij; float* delta1; float* delta2;                   // [ncon] scratch
    float4* xold; float4* xunc;                                  // [npad] scratch (user atom order)
    int maxIter;
};

#ifdef __CUDACC__
// ---------------------------------------------------------------- Philox4x32-10 (Salmon et al., SC'11)
__device__ __forceinline__ uint4 philox(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u*c.x;
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u*c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}

// three independent N(0,1) for (atom, step)
__device__ __forceinline__ float3 gauss3(unsigned int seed, int atom, unsigned long long step) {
    uint4 r = philox(make_uint4((unsigned int) atom, (unsigned int) step, (unsigned int) (step >> 32), 0x5eed5eedu), make_uint2(seed, 0xb200b200u));
    const float u1 = ((r.x >> 8) + 1u)*(1.0f/16777216.0f);      // (0,1]
    const float u2 = (r.y >> 8)*(1.0f/16777216.0f);
    const float u3 = ((r.z >> 8) + 1u)*(1.0f/16777216.0f);
    const float u4 = (r.w >> 8)*(1.0f/16777216.0f);
    const float m1 = sqrtf(-2.0f*logf(u1)), m2 = sqrtf(-2.0f*logf(u3));
    float s1, c1, s2, c2;
    sincospif(2.0f*u2, &s1, &c1);
    sincospif(2.0f*u4, &s2, &c2);
    (void) s2;
    return make_float3(m1*c1, m1*s1, m2*c2);
}

#endif

// ---- launchers (defined in the .cu files) ----
void launch_check_displacement(const NbDev& nb, const CommDev& cd, cudaStream_t s);
bool list_build_merged();        // list build = 2 gated launches (k_list_prep with grid barriers + k_build_tiles); B200MD_LIST_MERGED=0: 7
void launch_list_build(const NbDev& nb, cudaStream_t s, int mode = 0);   // all list kernels, gated on counters[CT_REBUILD] (mode 0) or counters[CT_SOFT] (mode 1)
void launch_pair(const NbDev& nb, bool energy, cudaStream_t s);
void launch_count_pairs(const NbDev& nb, cudaStream_t s);
int choose_pme_sms(int reserve, unsigned long long mask[4]);     // SM partition of the tile kernel (NbDev::pmeSmMask)
int  list_build_launch_count();

void launch_pme_eterm(const NbDev& nb, const PmeDev& pme, cudaStream_t s);
void launch_pme_spread(const NbDev& nb, const PmeDev& pme, const CommDev& cd, cudaStream_t s);
void launch_pme_fft_conv(const NbDev& nb, const PmeDev& pme, const CommDev& cd, bool energy, cudaStream_t s);
void launch_pme_gather(const NbDev& nb, const PmeDev& pme, const CommDev& cd, cudaStream_t s);
void launch_grid_push(const PmeDev& pme, const CommDev& cd, cudaStream_t s);
void launch_fft3d_r2c(const PmeDev& pme, cudaStream_t s);         // grid -> cgrid
void launch_fft3d_c2r(const PmeDev& pme, cudaStream_t s);         // cgrid -> grid
size_t fft_plane_smem_bytes(int ny, int nz);
size_t fft_line_smem_bytes(int nx);
bool fft_make_radices(int n, int* radix, int* nstages);
int pme_fft_launch_count(const PmeDev& pme);
void fft_set_compact(int on);                   // smaller FFT CTAs (the chain shares the GPU with the tile kernel)
bool fft_slab_path(const PmeDev& pme);          // the 3-launch slab pipeline is usable for this grid (precondition of the multi-GPU FFT)

void launch_bonded(const NbDev& nb, const BondedDev& bd, int terms, bool energy, cudaStream_t s);

void launch_integrate(const NbDev& nb, const UnitDev& units, const IntegDev& integ, const CommDev& cd, cudaStream_t s);
void launch_force_push(const NbDev& nb, const CommDev& cd, cudaStream_t s);          // partial forces of foreign atoms -> owners' inboxes
void launch_pos_push(const NbDev& nb, const CommDev& cd, const IntegDev& in, cudaStream_t s);   // owned positions -> every peer (TMA); publishes CH_POS
bool pos_push_available();
void launch_force_total(const NbDev& nb, const CommDev& cd, cudaStream_t s);         // compute path: owners total and broadcast, everybody waits
void launch_vel_push(const NbDev& nb, const CommDev& cd, cudaStream_t s);            // owners' velocities -> everybody (state reads)
void launch_pos_wait(const NbDev& nb, const CommDev& cd, cudaStream_t s);            // wait for the peers' position stores (state reads)
void launch_constrain_positions(const NbDev& nb, const UnitDev& units, float tol, cudaStream_t s);
void launch_constrain_velocities(const NbDev& nb, const UnitDev& units, float tol, cudaStream_t s);
void launch_kinetic_energy(const NbDev& nb, const UnitDev& units, const IntegDev& integ, float shiftDt, cudaStream_t s);
void launch_remove_cm(const NbDev& nb, double* scratch, cudaStream_t s);
void launch_ccma_step(const NbDev& nb, const CcmaDev& cc, const IntegDev& in, cudaStream_t s);      // before launch_integrate in a step
void launch_ccma_apply(const NbDev& nb, const CcmaDev& cc, bool velocities, float tol, cudaStream_t s);
void launch_ccma_kinetic(const NbDev& nb, const CcmaDev& cc, float shiftDt, float tol, cudaStream_t s);
void launch_cm_prime(const NbDev& nb, const IntegDev& integ, const CommDev& cd, cudaStream_t s);
