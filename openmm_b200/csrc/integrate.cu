// integrate.cu -- fused integrate + constrain step (sm_100a): one thread per integration unit
// (a rigid 3-atom molecule, an X-H_n SHAKE cluster or a free atom), everything in registers, ONE launch.
//
// Restates ReferenceStochasticDynamics::update (ReferenceStochasticDynamics.cpp:89-194),
// ReferenceLangevinMiddleDynamics::update (ReferenceLangevinMiddleDynamics.cpp:54-127), ReferenceVerletDynamics,
// ReferenceSETTLEAlgorithm::apply/applyToVelocities (ReferenceSETTLEAlgorithm.cpp:54-244) and the per-cluster SHAKE of
// the reference GPU platforms (integrationUtilities.cc:99-326).  Replaces integrateLangevinPart1 -> applySettle* ->
// applyShake* -> integrateLangevinPart2 -> generateRandomNumbers (4-6 launches + an RNG refill) with one kernel;
// the noise is Philox4x32-10 keyed on (seed; atom, step) so results do not depend on launch geometry.
#include "engine.h"
#include <algorithm>
#include "../../include/b200md.h"

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x+b.x, a.y+b.y, a.z+b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x-b.x, a.y-b.y, a.z-b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return {a.x*s, a.y*s, a.z*s}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x*b.x + a.y*b.y + a.z*b.z; }

// ---------------------------------------------------------------- SETTLE, positions (Miyamoto & Kollman 1992)
// NOTE on provenance: settle_positions / settle_velocities below restate the published closed form in the same sequence of
// steps (and with the same intermediate names) as the reference's GPU kernels applySettleToPositions / applySettleToVelocities
// (platforms/common/src/kernels/integrationUtilities.cc:328-470, 489-551), which themselves follow
// ReferenceSETTLEAlgorithm.cpp:54-244; only the types (fp32 registers, one thread per water inside the fused integrate
// kernel) and the surrounding kernel are this repository's own.  The SHAKE code further down is restructured (looped, per-
// constraint distances).
// x: old positions (constraints satisfied), d: position deltas (in/out), m: masses, dist1 = |01| = |02|, dist2 = |12|
__device__ void settle_positions(const V3* x, V3* d, const float* m, float dist1, float dist2) {
    const V3 xp0 = d[0], xp1 = d[1], xp2 = d[2];
    const float xb0 = x[1].x-x[0].x, yb0 = x[1].y-x[0].y, zb0 = x[1].z-x[0].z;
    const float xc0 = x[2].x-x[0].x, yc0 = x[2].y-x[0].y, zc0 = x[2].z-x[0].z;
    const float invTotalMass = 1.0f/(m[0]+m[1]+m[2]);
    const float xcom = (xp0.x*m[0] + (xb0+xp1.x)*m[1] + (xc0+xp2.x)*m[2])*invTotalMass;
    const float ycom = (xp0.y*m[0] + (yb0+xp1.y)*m[1] + (yc0+xp2.y)*m[2])*invTotalMass;
    const float zcom = (xp0.z*m[0] + (zb0+xp1.z)*m[1] + (zc0+xp2.z)*m[2])*invTotalMass;
    const float xa1 = xp0.x - xcom, ya1 = xp0.y - ycom, za1 = xp0.z - zcom;
    const float xb1 = xb0 + xp1.x - xcom, yb1 = yb0 + xp1.y - ycom, zb1 = zb0 + xp1.z - zcom;
    const float xc1 = xc0 + xp2.x - xcom, yc1 = yc0 + xp2.y - ycom, zc1 = zc0 + xp2.z - zcom;
    const float xaksZd = yb0*zc0 - zb0*yc0, yaksZd = zb0*xc0 - xb0*zc0, zaksZd = xb0*yc0 - yb0*xc0;
    const float xaksXd = ya1*zaksZd - za1*yaksZd, yaksXd = za1*xaksZd - xa1*zaksZd, zaksXd = xa1*yaksZd - ya1*xaksZd;
    const float xaksYd = yaksZd*zaksXd - zaksZd*yaksXd, yaksYd = zaksZd*xaksXd - xaksZd*zaksXd, zaksYd = xaksZd*yaksXd - yaksZd*xaksXd;
    const float axlng = rsqrtf(xaksXd*xaksXd + yaksXd*yaksXd + zaksXd*zaksXd);
    const float aylng = rsqrtf(xaksYd*xaksYd + yaksYd*yaksYd + zaksYd*zaksYd);
    const float azlng = rsqrtf(xaksZd*xaksZd + yaksZd*yaksZd + zaksZd*zaksZd);
    const float t11 = xaksXd*axlng, t21 = yaksXd*axlng, t31 = zaksXd*axlng;
    const float t12 = xaksYd*aylng, t22 = yaksYd*aylng, t32 = zaksYd*aylng;
    const float t13 = xaksZd*azlng, t23 = yaksZd*azlng, t33 = zaksZd*azlng;
    const float xb0d = t11*xb0 + t21*yb0 + t31*zb0, yb0d = t12*xb0 + t22*yb0 + t32*zb0;
    const float xc0d = t11*xc0 + t21*yc0 + t31*zc0, yc0d = t12*xc0 + t22*yc0 + t32*zc0;
    const float za1d = t13*xa1 + t23*ya1 + t33*za1;
    const float xb1d = t11*xb1 + t21*yb1 + t31*zb1, yb1d = t12*xb1 + t22*yb1 + t32*zb1, zb1d = t13*xb1 + t23*yb1 + t33*zb1;
    const float xc1d = t11*xc1 + t21*yc1 + t31*zc1, yc1d = t12*xc1 + t22*yc1 + t32*zc1, zc1d = t13*xc1 + t23*yc1 + t33*zc1;
    // step 2
    const float rc = 0.5f*dist2;
    float rb = sqrtf(dist1*dist1 - rc*rc);
    const float ra = rb*(m[1]+m[2])*invTotalMass;
    rb -= ra;
    const float sinphi = za1d/ra;
    const float cosphi = sqrtf(1.0f - sinphi*sinphi);
    const float sinpsi = (zb1d - zc1d)/(2.0f*rc*cosphi);
    const float cospsi = sqrtf(1.0f - sinpsi*sinpsi);
    const float ya2d = ra*cosphi;
    float xb2d = -rc*cospsi;
    const float yb2d = -rb*cosphi - rc*sinpsi*sinphi;
    const float yc2d = -rb*cosphi + rc*sinpsi*sinphi;
    const float xb2d2 = xb2d*xb2d;
    const float hh2 = 4.0f*xb2d2 + (yb2d-yc2d)*(yb2d-yc2d) + (zb1d-zc1d)*(zb1d-zc1d);
    const float deltx = 2.0f*xb2d + sqrtf(4.0f*xb2d2 - hh2 + dist2*dist2);
    xb2d -= deltx*0.5f;
    // step 3
    const float alpha = xb2d*(xb0d-xc0d) + yb0d*yb2d + yc0d*yc2d;
    const float beta = xb2d*(yc0d-yb0d) + xb0d*yb2d + xc0d*yc2d;
    const float gamma = xb0d*yb1d - xb1d*yb0d + xc0d*yc1d - xc1d*yc0d;
    const float al2be2 = alpha*alpha + beta*beta;
    const float sintheta = (alpha*gamma - beta*sqrtf(al2be2 - gamma*gamma))/al2be2;
    // step 4
    const float costheta = sqrtf(1.0f - sintheta*sintheta);
    const float xa3d = -ya2d*sintheta, ya3d = ya2d*costheta, za3d = za1d;
    const float xb3d = xb2d*costheta - yb2d*sintheta, yb3d = xb2d*sintheta + yb2d*costheta, zb3d = zb1d;
    const float xc3d = -xb2d*costheta - yc2d*sintheta, yc3d = -xb2d*sintheta + yc2d*costheta, zc3d = zc1d;
    // step 5
    const float xa3 = t11*xa3d + t12*ya3d + t13*za3d, ya3 = t21*xa3d + t22*ya3d + t23*za3d, za3 = t31*xa3d + t32*ya3d + t33*za3d;
    const float xb3 = t11*xb3d + t12*yb3d + t13*zb3d, yb3 = t21*xb3d + t22*yb3d + t23*zb3d, zb3 = t31*xb3d + t32*yb3d + t33*zb3d;
    const float xc3 = t11*xc3d + t12*yc3d + t13*zc3d, yc3 = t21*xc3d + t22*yc3d + t23*zc3d, zc3 = t31*xc3d + t32*yc3d + t33*zc3d;
    d[0] = {xcom + xa3, ycom + ya3, zcom + za3};
    d[1] = {xcom + xb3 - xb0, ycom + yb3 - yb0, zcom + zb3 - zb0};
    d[2] = {xcom + xc3 - xc0, ycom + yc3 - yc0, zcom + zc3 - zc0};
}

// SETTLE, velocities (ReferenceSETTLEAlgorithm.cpp:197-244; general masses)
__device__ void settle_velocities(const V3* x, V3* v, const float* m) {
    V3 eAB = x[1]-x[0], eBC = x[2]-x[1], eCA = x[0]-x[2];
    eAB = eAB*rsqrtf(dot(eAB, eAB)); eBC = eBC*rsqrtf(dot(eBC, eBC)); eCA = eCA*rsqrtf(dot(eCA, eCA));
    const float vAB = dot(v[1]-v[0], eAB), vBC = dot(v[2]-v[1], eBC), vCA = dot(v[0]-v[2], eCA);
    const float cA = -dot(eAB, eCA), cB = -dot(eAB, eBC), cC = -dot(eBC, eCA);
    const float s2A = 1-cA*cA, s2B = 1-cB*cB, s2C = 1-cC*cC;
    const float mA = m[0], mB = m[1], mC = m[2];
    const float mABCinv = 1.0f/(mA*mB*mC);
    const float denom = (((s2A*mB+s2B*mA)*mC+(s2A*mB*mB+2*(cA*cB*cC+1)*mA*mB+s2B*mA*mA))*mC+s2C*mA*mB*(mA+mB))*mABCinv;
    const float tab = ((cB*cC*mA-cA*mB-cA*mC)*vCA + (cA*cC*mB-cB*mC-cB*mA)*vBC + (s2C*mA*mA*mB*mB*mABCinv+(mA+mB+mC))*vAB)/denom;
    const float tbc = ((cA*cB*mC-cC*mB-cC*mA)*vCA + (s2A*mB*mB*mC*mC*mABCinv+(mA+mB+mC))*vBC + (cA*cC*mB-cB*mA-cB*mC)*vAB)/denom;
    const float tca = ((s2B*mA*mA*mC*mC*mABCinv+(mA+mB+mC))*vCA + (cA*cB*mC-cC*mB-cC*mA)*vBC + (cB*cC*mA-cA*mB-cA*mC)*vAB)/denom;
    v[0] = v[0] + (eAB*tab - eCA*tca)*(1.0f/mA);
    v[1] = v[1] + (eBC*tbc - eAB*tab)*(1.0f/mB);
    v[2] = v[2] + (eCA*tca - eBC*tbc)*(1.0f/mC);
}

// SHAKE on a centre + n hydrogens, positions (deltas d relative to old positions x)
__device__ void shake_positions(const V3* x, V3* d, const float* invM, const float* dist, int n, float tol) {
    V3 rij[3]; float rij2[3], ld[3];
    _Pragma("unroll") for (int k = 0; k < 3; k++) if (k < n) {
        rij[k] = x[0] - x[k+1];
        rij2[k] = dot(rij[k], rij[k]);
        ld[k] = dist[k]*dist[k] - rij2[k];
    }
    bool converged = false;
    for (int it = 0; it < 30 && !converged; it++) {
        converged = true;
        _Pragma("unroll") for (int k = 0; k < 3; k++) if (k < n) {
            const V3 rp = d[0] - d[k+1];
            const float rp2 = dot(rp, rp), rrpr = dot(rij[k], rp);
            const float diff = ld[k] - 2.0f*rrpr - rp2;
            const float d2 = dist[k]*dist[k];
            if (fabsf(diff) >= d2*tol) {
                const float acor = diff*0.5f/((invM[0] + invM[k+1])*(rrpr + rij2[k]));
                d[0] = d[0] + rij[k]*(acor*invM[0]);
                d[k+1] = d[k+1] - rij[k]*(acor*invM[k+1]);
                converged = false;
            }
        }
    }
}

__device__ void shake_velocities(const V3* x, V3* v, const float* invM, int n, float tol) {
    V3 rij[3]; float rij2[3];
    _Pragma("unroll") for (int k = 0; k < 3; k++) if (k < n) { rij[k] = x[0] - x[k+1]; rij2[k] = dot(rij[k], rij[k]); }
    bool converged = false;
    for (int it = 0; it < 30 && !converged; it++) {
        converged = true;
        _Pragma("unroll") for (int k = 0; k < 3; k++) if (k < n) {
            const V3 rp = v[0] - v[k+1];
            const float rrpr = dot(rp, rij[k]);
            const float delta = -rrpr/((invM[0] + invM[k+1])*rij2[k]);
            v[0] = v[0] + rij[k]*(delta*invM[0]);
            v[k+1] = v[k+1] - rij[k]*(delta*invM[k+1]);
            if (fabsf(delta) > tol) converged = false;
        }
    }
}

struct Unit {
    int n;             // atoms in the unit
    int atom[4];
    int type;
    V3 x[4], v[4], f[4];
    float invM[4], m[4];
    float4 prm;
};

__device__ __forceinline__ bool load_unit(const NbDev& nb, const UnitDev& un, int u, Unit& U, bool wantForce, const CommDev* cd = nullptr) {
    const int4 at = un.unitAtoms[u];
    U.atom[0] = at.x; U.atom[1] = at.y; U.atom[2] = at.z; U.atom[3] = at.w;
    U.type = un.unitType[u];
    U.prm = un.unitParams[u];
    U.n = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int a = U.atom[k];
        if (a < 0) continue;
        U.n = k+1;
        const float4 p = nb.posq[a];
        const float4 v = nb.velm[a];
        U.x[k] = {p.x, p.y, p.z};
        U.v[k] = {v.x, v.y, v.z};
        U.invM[k] = v.w;
        U.m[k] = (v.w > 0.f) ? 1.0f/v.w : 0.f;
        if (wantForce) {
            long long fx = nb.force[a], fy = nb.force[a + nb.npad], fz = nb.force[a + 2*nb.npad];
            if (cd != nullptr && cd->world > 1) {
                // owner: total = own partial + what the other ranks pushed into the inboxes (exact int64 sums, any order)
                const long long* in = (const long long*) (cd->peer[cd->rank] + cd->offFinbox);
                for (int q = 0; q < cd->world; q++) if (q != cd->rank) {
                    const long long* iq = in + (size_t) q*3*nb.npad;
                    fx += iq[a]; fy += iq[a + nb.npad]; fz += iq[a + 2*nb.npad];
                }
            }
            U.f[k] = {fixed_to_float(fx), fixed_to_float(fy), fixed_to_float(fz)};
        }
    }
    return true;
}

__device__ __forceinline__ void constrain_pos(const Unit& U, V3* d, float tol) {
    if (U.type == 1) settle_positions(U.x, d, U.m, U.prm.x, U.prm.y);
    else if (U.type == 2) { const float dist[3] = {U.prm.x, U.prm.y, U.prm.z}; shake_positions(U.x, d, U.invM, dist, U.n-1, tol); }
}
__device__ __forceinline__ void constrain_vel(const Unit& U, V3* v, float tol) {
    if (U.type == 1) settle_velocities(U.x, v, U.m);
    else if (U.type == 2) shake_velocities(U.x, v, U.invM, U.n-1, tol);
}

// Fused epilogue work (in.fused != 0, the b200md_step path):
//  * centre-of-mass motion removal (CMMotionRemover, frequency 1): the momentum of the velocities this kernel WRITES is
//    reduced into cm[(step+1)%3]; the next step subtracts cm[step%3]/mass before integrating (same velocities, so the same
//    result as removing it at the start of that step, ReferenceKernels RemoveCMMotion); cm[(step+2)%3] is zeroed for reuse.
//  * the force buffer is zeroed after it has been read (saves the memset node at the head of the next step's graph).
//  * the last block to finish advances the step counter (saves a 1-thread kernel).
//
// Multi-GPU (cd.world > 1): this rank integrates the units it OWNS.  The kernel first waits for the other ranks' partial
// forces (CH_FORCE), totals them with its own, and stores the new positions into EVERY rank's posq over NVLink -- the
// integrate step IS the position all-gather.  The last block hands its momentum sums to everybody and publishes CH_POS.
template <int KIND>
__global__ void __launch_bounds__(128) k_integrate(NbDev nb, UnitDev un, IntegDev in, CommDev cd) {
    const bool multi = cd.world > 1;
    const bool useInbox = multi && in.fused;       // step path; after b200md_compute the force buffer already holds the totals
    const unsigned long long E = multi ? *cd.epoch + 1ull : 0ull;
    if (useInbox) comm_wait(cd, CH_FORCE, E);
    const int u = (multi ? cd.unitLo[cd.rank] : 0) + blockIdx.x*blockDim.x + threadIdx.x;
    const bool active = u < (multi ? cd.unitLo[cd.rank + 1] : un.nunits);
    const unsigned long long step = *in.stepCounter;
    Unit U;
    U.n = 0;
    V3 d[4];
    const float invDt = 1.0f/in.dt;
    const bool cmFused = in.fused && in.cmEveryStep;
    V3 vcm = {0.f, 0.f, 0.f};
    if (cmFused) {
        double c[4];
        if (multi) {
            // every rank left the momentum of ITS atoms in slot [rank][step % 3] of everybody's table
            const double* t = (const double*) (cd.peer[cd.rank] + cd.offCm) + 4*(step % 3ull);
            c[0] = c[1] = c[2] = c[3] = 0.0;
            for (int q = 0; q < cd.world; q++) for (int k = 0; k < 4; k++) c[k] += t[q*12 + k];
        }
        else { const double* t = in.cmScratch + 4*(step % 3ull); c[0] = t[0]; c[1] = t[1]; c[2] = t[2]; c[3] = t[3]; }
        const double im = (c[3] > 0.0) ? 1.0/c[3] : 0.0;
        vcm = {(float) (c[0]*im), (float) (c[1]*im), (float) (c[2]*im)};
    }
    if (active) {
    load_unit(nb, un, u, U, true, useInbox ? &cd : nullptr);
    if (in.fused) {
        _Pragma("unroll") for (int k = 0; k < 4; k++) if (k < U.n) {
            const int a = U.atom[k];
            nb.force[a] = 0; nb.force[a + nb.npad] = 0; nb.force[a + 2*nb.npad] = 0;
            if (U.invM[k] > 0.f) U.v[k] = U.v[k] - vcm;
        }
    }
    if (KIND == B200MD_INT_LANGEVIN_MIDDLE) {
        _Pragma("unroll") for (int k = 0; k < 4; k++) if (k < U.n) U.v[k] = U.v[k] + U.f[k]*(in.dt*U.invM[k]);
        constrain_vel(U, U.v, in.tol);
        V3 du[4];
        _Pragma("unroll") for (int k = 0; k < 4; k++) if (k < U.n) {
            d[k] = U.v[k]*(0.5f*in.dt);
            if (U.invM[k] > 0.f) {
                const float3 g = gauss3(in.seed, U.atom[k], step);
                const float ns = in.noisescale*sqrtf(in.kT*U.invM[k]);
                U.v[k] = U.v[k]*in.vscale + V3{g.x, g.y, g.z}*ns;
            }
            d[k] = d[k] + U.v[k]*(0.5f*in.dt);
            if (U.invM[k] == 0.f) d[k] = {0.f, 0.f, 0.f};
            du[k] = d[k];
        }
        constrain_pos(U, d, in.tol);
        _Pragma("unroll") for (int k = 0; k < 4; k++) if (k < U.n) U.v[k] = U.v[k] + (d[k] - du[k])*invDt;
    }
    else {
        _Pragma("unroll") for (int k = 0; k < 4; k++) if (k < U.n) {
            V3 vn;
            if (KIND == B200MD_INT_LANGEVIN) {
                vn = U.v[k]*in.vscale + U.f[k]*(in.fscale*U.invM[k]);
                if (U.invM[k] > 0.f && in.noisescale > 0.f) {
                    const float3 g = gauss3(in.seed, U.atom[k], step);
                    vn = vn + V3{g.x, g.y, g.z}*(in.noisescale*sqrtf(U.invM[k]));
                }
            }
            else
                vn = U.v[k] + U.f[k]*(in.dt*U.invM[k]);
            if (U.invM[k] == 0.f) vn = U.v[k];
            d[k] = (U.invM[k] == 0.f) ? V3{0.f, 0.f, 0.f} : vn*in.dt;
        }
        constrain_pos(U, d, in.tol);
        _Pragma("unroll") for (int k = 0; k < 4; k++) if (k < U.n) if (U.invM[k] > 0.f) U.v[k] = d[k]*invDt;
    }
    _Pragma("unroll") for (int k = 0; k < 4; k++) if (k < U.n) {
        const int a = U.atom[k];
        const float4 p = nb.posq[a];
        const float4 pn = make_float4(U.x[k].x + d[k].x, U.x[k].y + d[k].y, U.x[k].z + d[k].z, p.w);
        nb.posq[a] = pn;
        nb.velm[a] = make_float4(U.v[k].x, U.v[k].y, U.v[k].z, U.invM[k]);
        if (multi && !cd.posByPush)
            for (int k = 1; k < cd.world; k++) {            // staggered: rank r starts with peer r+1, so the ranks do not all hit peer 0 first
                const int q = (cd.rank + k) % cd.world;
                ((float4*) (cd.peer[q] + cd.offPosq))[a] = pn;
            }
    }
    }   // active
    if (!in.fused && !(multi && !cd.posByPush)) return;
    if (cmFused) {
        double px = 0, py = 0, pz = 0, m = 0;
        _Pragma("unroll") for (int k = 0; k < 4; k++) if (k < U.n && U.invM[k] > 0.f) {
            const double mk = U.m[k];
            px += mk*U.v[k].x; py += mk*U.v[k].y; pz += mk*U.v[k].z; m += mk;
        }
        for (int off = 16; off > 0; off >>= 1) {
            px += __shfl_xor_sync(0xffffffffu, px, off); py += __shfl_xor_sync(0xffffffffu, py, off);
            pz += __shfl_xor_sync(0xffffffffu, pz, off); m += __shfl_xor_sync(0xffffffffu, m, off);
        }
        __shared__ double red[4][4];
        if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = px; red[threadIdx.x >> 5][1] = py; red[threadIdx.x >> 5][2] = pz; red[threadIdx.x >> 5][3] = m; }
        __syncthreads();
        if (threadIdx.x < 4) {
            double t = 0;
            for (int w = 0; w < (int) (blockDim.x >> 5); w++) t += red[w][threadIdx.x];
            atomicAdd(&in.cmScratch[4*((step + 1ull) % 3ull) + threadIdx.x], t);
            if (blockIdx.x == 0) in.cmScratch[4*((step + 2ull) % 3ull) + threadIdx.x] = 0.0;
        }
    }
    if (multi && !cd.posByPush) {
        // last block: momentum sums of this rank's atoms -> slot [rank][(step+1) % 3] of every rank's table, then CH_POS
        const bool lastBlock = comm_arrive(cd, CH_POS, gridDim.x);
        if (lastBlock && threadIdx.x == 0) {
            if (cmFused) {
                const double* mine = in.cmScratch + 4*((step + 1ull) % 3ull);
                for (int q = 0; q < cd.world; q++) {
                    double* t = (double*) (cd.peer[q] + cd.offCm) + cd.rank*12 + 4*((step + 1ull) % 3ull);
                    for (int k = 0; k < 4; k++) t[k] = ((volatile const double*) mine)[k];
                }
            }
            comm_publish(cd, CH_POS, E);
            *cd.posNeed = E;
            *cd.epoch = E;
            *in.stepCounter = step + 1ull;
            if (nb.counters[CT_PENDING]) { nb.counters[CT_PENDING] = 0; nb.counters[CT_CUR] ^= 1; }
        }
        return;
    }
    // last block to finish: advance the step counter
    __shared__ int last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = (atomicAdd(in.blocksDone, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        *in.blocksDone = 0u;
        *in.stepCounter = step + 1ull;
        // a successor neighbour list built beside this step (nonbonded.cu: k_check_gather / k_list_done) becomes current
        if (nb.counters[CT_PENDING]) { nb.counters[CT_PENDING] = 0; nb.counters[CT_CUR] ^= 1; }
    }
}

__global__ void k_step_advance(IntegDev in) { *in.stepCounter += 1ull; }

// momentum of the current velocities into cm[step % 3] (validates the fused scheme after the state was set from outside)
__global__ void __launch_bounds__(256) k_cm_prime(NbDev nb, IntegDev in, double* table) {
    const unsigned long long step = *in.stepCounter;
    double* c = table + 4*(step % 3ull);
    const int a = blockIdx.x*blockDim.x + threadIdx.x;
    double px = 0, py = 0, pz = 0, m = 0;
    if (a < nb.natoms) {
        const float4 v = nb.velm[a];
        if (v.w > 0.f) { m = 1.0/v.w; px = m*v.x; py = m*v.y; pz = m*v.z; }
    }
    for (int off = 16; off > 0; off >>= 1) {
        px += __shfl_xor_sync(0xffffffffu, px, off); py += __shfl_xor_sync(0xffffffffu, py, off);
        pz += __shfl_xor_sync(0xffffffffu, pz, off); m += __shfl_xor_sync(0xffffffffu, m, off);
    }
    __shared__ double red[8][4];
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = px; red[threadIdx.x >> 5][1] = py; red[threadIdx.x >> 5][2] = pz; red[threadIdx.x >> 5][3] = m; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0;
        for (int w = 0; w < 8; w++) t += red[w][threadIdx.x];
        atomicAdd(&c[threadIdx.x], t);
    }
}

// Multi-GPU: the velocities are in sync at this point (the caller ran launch_vel_push), every rank computes the same total
// over ALL atoms and leaves it in row 0 of its own table (the other rows zero): k_integrate sums the rows.
void launch_cm_prime(const NbDev& nb, const IntegDev& integ, const CommDev& cd, cudaStream_t s) {
    cudaMemsetAsync(integ.cmScratch, 0, 12*sizeof(double), s);
    double* table = integ.cmScratch;
    if (cd.world > 1) {
        table = (double*) (cd.peer[cd.rank] + cd.offCm);
        cudaMemsetAsync(table, 0, (size_t) B200MD_MAX_RANKS*12*sizeof(double), s);
    }
    k_cm_prime<<<(nb.natoms + 255)/256, 256, 0, s>>>(nb, integ, table);
}

void launch_integrate(const NbDev& nb, const UnitDev& units, const IntegDev& integ, const CommDev& cd, cudaStream_t s) {
    // 64-thread blocks: at DHFR size (8k units) 128-thread blocks fill only 65 of the 148 SMs
    const int n = cd.world > 1 ? cd.unitLo[cd.rank + 1] - cd.unitLo[cd.rank] : units.nunits;
    const int grid = std::max(1, (n + 63)/64);
    if (integ.kind == B200MD_INT_VERLET) k_integrate<B200MD_INT_VERLET><<<grid, 64, 0, s>>>(nb, units, integ, cd);
    else if (integ.kind == B200MD_INT_LANGEVIN) k_integrate<B200MD_INT_LANGEVIN><<<grid, 64, 0, s>>>(nb, units, integ, cd);
    else k_integrate<B200MD_INT_LANGEVIN_MIDDLE><<<grid, 64, 0, s>>>(nb, units, integ, cd);
    if (!integ.fused && (cd.world <= 1 || cd.posByPush)) k_step_advance<<<1, 1, 0, s>>>(integ);
    // multi-GPU, posByPush: the new positions of the owned atoms (one contiguous range) go to every peer through the TMA
    // engine in a kernel of their own, which also carries the momentum sums and publishes CH_POS
    if (cd.world > 1 && cd.posByPush) launch_pos_push(nb, cd, integ, s);
}

// ApplyConstraintsKernel::apply: project the current positions onto the constraints (reference & target identical)
__global__ void __launch_bounds__(128) k_constrain_positions(NbDev nb, UnitDev un, float tol) {
    const int u = blockIdx.x*blockDim.x + threadIdx.x;
    if (u >= un.nunits) return;
    Unit U;
    load_unit(nb, un, u, U, false);
    if (U.type == 0) return;
    V3 d[4] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    constrain_pos(U, d, tol);
    _Pragma("unroll") for (int k = 0; k < 4; k++) if (k < U.n) {
        const int a = U.atom[k];
        const float4 p = nb.posq[a];
        nb.posq[a] = make_float4(p.x + d[k].x, p.y + d[k].y, p.z + d[k].z, p.w);
    }
}

__global__ void __launch_bounds__(128) k_constrain_velocities(NbDev nb, UnitDev un, float tol) {
    const int u = blockIdx.x*blockDim.x + threadIdx.x;
    if (u >= un.nunits) return;
    Unit U;
    load_unit(nb, un, u, U, false);
    if (U.type == 0) return;
    constrain_vel(U, U.v, tol);
    _Pragma("unroll") for (int k = 0; k < 4; k++) if (k < U.n) nb.velm[U.atom[k]] = make_float4(U.v[k].x, U.v[k].y, U.v[k].z, U.invM[k]);
}

void launch_constrain_positions(const NbDev& nb, const UnitDev& units, float tol, cudaStream_t s) {
    if (units.nunits == 0) return;          // every atom sits in a general constraint network (constraints.cu)
    k_constrain_positions<<<(units.nunits + 127)/128, 128, 0, s>>>(nb, units, tol);
}
void launch_constrain_velocities(const NbDev& nb, const UnitDev& units, float tol, cudaStream_t s) {
    if (units.nunits == 0) return;
    k_constrain_velocities<<<(units.nunits + 127)/128, 128, 0, s>>>(nb, units, tol);
}

// kinetic energy at time-shifted, re-constrained velocities (computeShiftedKineticEnergy, ReferenceKernels.cpp:146-176)
__global__ void __launch_bounds__(128) k_kinetic_energy(NbDev nb, UnitDev un, float shiftDt) {
    const int u = blockIdx.x*blockDim.x + threadIdx.x;
    double ke = 0.0;
    if (u < un.nunits) {
        Unit U;
        load_unit(nb, un, u, U, true);
        if (shiftDt != 0.f) {
            _Pragma("unroll") for (int k = 0; k < 4; k++) if (k < U.n) U.v[k] = U.v[k] + U.f[k]*(shiftDt*U.invM[k]);
            constrain_vel(U, U.v, 1e-4f);
        }
        _Pragma("unroll") for (int k = 0; k < 4; k++) if (k < U.n) ke += 0.5*(double) U.m[k]*(double) dot(U.v[k], U.v[k]);
    }
    for (int off = 16; off > 0; off >>= 1) ke += __shfl_xor_sync(0xffffffffu, ke, off);
    __shared__ double red[4];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ke;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(&nb.energy[EN_KE], red[0] + red[1] + red[2] + red[3]);
}

void launch_kinetic_energy(const NbDev& nb, const UnitDev& units, const IntegDev& integ, float shiftDt, cudaStream_t s) {
    (void) integ;
    if (units.nunits == 0) return;
    k_kinetic_energy<<<(units.nunits + 127)/128, 128, 0, s>>>(nb, units, shiftDt);
}

// CMMotionRemover (ReferenceKernels.cpp RemoveCMMotion): subtract the centre-of-mass velocity.
__global__ void __launch_bounds__(256) k_cm_sum(NbDev nb, double* scratch) {
    const int a = blockIdx.x*blockDim.x + threadIdx.x;
    double px = 0, py = 0, pz = 0, m = 0;
    if (a < nb.natoms) {
        const float4 v = nb.velm[a];
        if (v.w > 0.f) { m = 1.0/v.w; px = m*v.x; py = m*v.y; pz = m*v.z; }
    }
    for (int off = 16; off > 0; off >>= 1) {
        px += __shfl_xor_sync(0xffffffffu, px, off); py += __shfl_xor_sync(0xffffffffu, py, off);
        pz += __shfl_xor_sync(0xffffffffu, pz, off); m += __shfl_xor_sync(0xffffffffu, m, off);
    }
    __shared__ double red[8][4];
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = px; red[threadIdx.x >> 5][1] = py; red[threadIdx.x >> 5][2] = pz; red[threadIdx.x >> 5][3] = m; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0;
        for (int w = 0; w < 8; w++) t += red[w][threadIdx.x];
        atomicAdd(&scratch[threadIdx.x], t);
    }
}
__global__ void k_cm_apply(NbDev nb, double* scratch) {
    const int a = blockIdx.x*blockDim.x + threadIdx.x;
    if (a >= nb.natoms) return;
    const double im = 1.0/scratch[3];
    float4 v = nb.velm[a];
    if (v.w > 0.f) {
        v.x -= (float) (scratch[0]*im); v.y -= (float) (scratch[1]*im); v.z -= (float) (scratch[2]*im);
        nb.velm[a] = v;
    }
}
void launch_remove_cm(const NbDev& nb, double* scratch, cudaStream_t s) {
    cudaMemsetAsync(scratch, 0, 4*sizeof(double), s);
    k_cm_sum<<<(nb.natoms + 255)/256, 256, 0, s>>>(nb, scratch);
    k_cm_apply<<<(nb.natoms + 255)/256, 256, 0, s>>>(nb, scratch);
}
