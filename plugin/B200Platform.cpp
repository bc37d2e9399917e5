// B200Platform.cpp -- libOpenMMB200.so: the OpenMM Platform plugin for the B200-native hot path.
//
// A thin C++ adapter: every KernelImpl below forwards one abstract kernel interface of the reference
// (olla/include/openmm/kernels.h) to the C-ABI of libb200md.so (include/b200md.h), where all the CUDA lives.
// Loaded by Platform::loadPluginLibrary / loadPluginsFromDirectory (olla/src/Platform.cpp:221-320) through the
// extern "C" registerPlatforms() entry point (PluginInitializer.h:47), exactly like platforms/cuda
// (CudaPlatform.cpp:59-61).  Compiled against the reference's headers where they lie; no reference source is copied.
//
// Supported: NonbondedForce (NoCutoff, CutoffNonPeriodic, CutoffPeriodic, PME; parameter offsets), HarmonicBondForce,
// HarmonicAngleForce, PeriodicTorsionForce (any number of objects, each in its own force group), CMMotionRemover;
// Verlet / Langevin / LangevinMiddle integrators; SETTLE + X-H_n SHAKE constraints.  A Force class without a kernel here
// makes Platform::supportsKernels() false; an unsupported OPTION of a supported class is rejected in contextCreated()
// (validateSystem), which is the only place from which ContextImpl falls back to the next platform (ContextImpl.cpp:152-166).
//
// The hot loop.  Integrator::step(n) of the reference is, per step, updateContextState() -> calcForcesAndEnergy(true,
// false, groups) -> Integrate*StepKernel::execute (LangevinIntegrator.cpp:74-82).  Here a forces-only evaluation is
// LAZY: finishComputation records what was asked for and returns; when the integrator kernel's execute() follows and the
// request covered the whole force field, ONE b200md_step(1) replays the captured step graph (list check, tile kernel,
// reciprocal space, bonded terms, fused integrate + constraints + CM removal).  Anything that needs the forces before
// that (getForces, kinetic energy, energies) flushes the pending evaluation with b200md_compute first.
#include "openmm/Platform.h"
#include "openmm/KernelFactory.h"
#include "openmm/kernels.h"
#include "openmm/OpenMMException.h"
#include "openmm/System.h"
#include "openmm/Context.h"
#include "openmm/NonbondedForce.h"
#include "openmm/HarmonicBondForce.h"
#include "openmm/HarmonicAngleForce.h"
#include "openmm/PeriodicTorsionForce.h"
#include "openmm/CMMotionRemover.h"
#include "openmm/VerletIntegrator.h"
#include "openmm/LangevinIntegrator.h"
#include "openmm/LangevinMiddleIntegrator.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/internal/NonbondedForceImpl.h"
#include "../include/b200md.h"
#include <map>
#include <string>
#include <vector>
#include <sstream>
#include <cstdlib>

using namespace OpenMM;
using namespace std;

namespace {

// next grid size whose prime factors are <= 13 (what the bespoke FFT handles; the reference CUDA platform rounds to
// 2,3,5,7-smooth sizes the same way, CudaKernels.cpp:698-700)
int fftFriendly(int n) {
    for (;; n++) {
        int r = n;
        for (int p : {2, 3, 5, 7, 11, 13}) while (r % p == 0) r /= p;
        if (r == 1) return n;
    }
}

// per-Context state (ContextImpl::setPlatformData)
struct PlatformData {
    b200md_ctx* ctx = nullptr;
    int numParticles = 0;
    bool finalized = false;
    int pendingTerms = 0;
    bool includeEnergy = false;
    // lazy forces-only evaluation (see the header comment)
    bool lazyForces = false;
    int lazyTerms = 0;
    unsigned int lazyGroups = 0;
    int systemTerms = 0;            // every term the System has a kernel for
    unsigned int bondedGroupsUsed = 0;
    int cmFrequency = 0;
    bool cmRequested = false;       // RemoveCMMotionKernel::execute seen since the last integrator step
    bool useFusedStep = true;       // B200MD_PLUGIN_FUSED=0: always compute + integrate_only (debugging)
    vector<int> bondG, angG, torG;  // force group of every bonded element
    // bonded terms are gathered over all force objects of a kind and sent at finalize
    vector<int> bondI, bondJ; vector<double> bondR0, bondK;
    vector<int> angI, angJ, angK; vector<double> angT0, angKK;
    vector<int> torI, torJ, torK, torL, torN; vector<double> torPhase, torKK;
    map<string, string> props;
    int integratorKind = -1;
    double dt = 0, temperature = 0, friction = 0, tol = 0;
    int seed = 0;
    void check(int rc) const {
        if (rc != 0) throw OpenMMException(string("B200 platform: ") + b200md_last_error(ctx));
    }
    void ensureFinalized() {
        if (finalized) return;
        if (!bondI.empty()) check(b200md_set_bonds(ctx, (int) bondI.size(), bondI.data(), bondJ.data(), bondR0.data(), bondK.data()));
        if (!angI.empty()) check(b200md_set_angles(ctx, (int) angI.size(), angI.data(), angJ.data(), angK.data(), angT0.data(), angKK.data()));
        if (!torI.empty()) check(b200md_set_torsions(ctx, (int) torI.size(), torI.data(), torJ.data(), torK.data(), torL.data(), torN.data(), torPhase.data(), torKK.data()));
        if (!bondG.empty()) check(b200md_set_bonded_groups(ctx, 0, (int) bondG.size(), bondG.data()));
        if (!angG.empty()) check(b200md_set_bonded_groups(ctx, 1, (int) angG.size(), angG.data()));
        if (!torG.empty()) check(b200md_set_bonded_groups(ctx, 2, (int) torG.size(), torG.data()));
        check(b200md_finalize(ctx));
        finalized = true;
    }
    // a pending forces-only evaluation becomes real (someone reads the forces before an integrator step consumes it)
    void flushForces() {
        if (!lazyForces) return;
        lazyForces = false;
        check(b200md_compute_groups(ctx, lazyTerms, lazyGroups, 1, nullptr));
    }
    void dropForces() { lazyForces = false; }
    // RemoveCMMotionKernel::execute is deferred too: the fused step removes the centre-of-mass motion itself
    void flushCm() {
        if (!cmRequested) return;
        cmRequested = false;
        if (cmFrequency > 0 && b200md_get_step_count(ctx) % cmFrequency == 0) check(b200md_remove_cm_motion(ctx));
    }
};

PlatformData& getData(ContextImpl& context) { return *reinterpret_cast<PlatformData*>(context.getPlatformData()); }
const PlatformData& getData(const ContextImpl& context) { return *reinterpret_cast<const PlatformData*>(const_cast<ContextImpl&>(context).getPlatformData()); }

// ------------------------------------------------------------------------------------------------ mandatory kernels
class B200CalcForcesAndEnergyKernel : public CalcForcesAndEnergyKernel {
public:
    B200CalcForcesAndEnergyKernel(string name, const Platform& platform) : CalcForcesAndEnergyKernel(name, platform) {}
    void initialize(const System& system) {}
    void beginComputation(ContextImpl& context, bool includeForce, bool includeEnergy, int groups) {
        PlatformData& d = getData(context);
        d.ensureFinalized();
        d.dropForces();             // superseded by this evaluation
        d.pendingTerms = 0;
        d.includeEnergy = includeEnergy;
    }
    double finishComputation(ContextImpl& context, bool includeForce, bool includeEnergy, int groups, bool& valid) {
        // every Calc*ForceKernel::execute only recorded its term; one engine call evaluates them all on the device
        PlatformData& d = getData(context);
        double energy = 0;
        valid = true;
        if (includeForce && !includeEnergy) {
            d.lazyForces = true; d.lazyTerms = d.pendingTerms; d.lazyGroups = (unsigned int) groups;
            return 0.0;
        }
        d.check(b200md_compute_groups(d.ctx, d.pendingTerms, (unsigned int) groups, includeForce ? 1 : 0, includeEnergy ? &energy : nullptr));
        return energy;
    }
};

class B200UpdateStateDataKernel : public UpdateStateDataKernel {
public:
    B200UpdateStateDataKernel(string name, const Platform& platform) : UpdateStateDataKernel(name, platform) {}
    void initialize(const System& system) {
        // masses and constraints belong to the System, not to a Force: hand them over here (once per Context)
        // (done in B200Platform::contextCreated, which has the ContextImpl)
    }
    double getTime(const ContextImpl& context) const { return b200md_get_time(getData(context).ctx); }
    void setTime(ContextImpl& context, double time) { b200md_set_time(getData(context).ctx, time); }
    void getPositions(ContextImpl& context, vector<Vec3>& positions) {
        PlatformData& d = getData(context);
        d.ensureFinalized();
        vector<double> x(3*d.numParticles);
        d.check(b200md_get_positions(d.ctx, x.data()));
        positions.resize(d.numParticles);
        for (int i = 0; i < d.numParticles; i++) positions[i] = Vec3(x[3*i], x[3*i+1], x[3*i+2]);
    }
    void setPositions(ContextImpl& context, const vector<Vec3>& positions) {
        PlatformData& d = getData(context);
        d.ensureFinalized();
        d.dropForces();
        vector<double> x(3*d.numParticles);
        for (int i = 0; i < d.numParticles; i++) for (int k = 0; k < 3; k++) x[3*i+k] = positions[i][k];
        d.check(b200md_set_positions(d.ctx, x.data()));
    }
    void getVelocities(ContextImpl& context, vector<Vec3>& velocities) {
        PlatformData& d = getData(context);
        d.ensureFinalized();
        d.flushCm();
        vector<double> x(3*d.numParticles);
        d.check(b200md_get_velocities(d.ctx, x.data()));
        velocities.resize(d.numParticles);
        for (int i = 0; i < d.numParticles; i++) velocities[i] = Vec3(x[3*i], x[3*i+1], x[3*i+2]);
    }
    void setVelocities(ContextImpl& context, const vector<Vec3>& velocities) {
        PlatformData& d = getData(context);
        d.ensureFinalized();
        d.cmRequested = false;
        vector<double> x(3*d.numParticles);
        for (int i = 0; i < d.numParticles; i++) for (int k = 0; k < 3; k++) x[3*i+k] = velocities[i][k];
        d.check(b200md_set_velocities(d.ctx, x.data()));
    }
    void getForces(ContextImpl& context, vector<Vec3>& forces) {
        PlatformData& d = getData(context);
        d.ensureFinalized();
        d.flushForces();
        vector<double> x(3*d.numParticles);
        d.check(b200md_get_forces(d.ctx, x.data()));
        forces.resize(d.numParticles);
        for (int i = 0; i < d.numParticles; i++) forces[i] = Vec3(x[3*i], x[3*i+1], x[3*i+2]);
    }
    void getEnergyParameterDerivatives(ContextImpl& context, map<string, double>& derivs) {}
    void getPeriodicBoxVectors(ContextImpl& context, Vec3& a, Vec3& b, Vec3& c) const {
        double x[3], y[3], z[3];
        b200md_get_box(getData(context).ctx, x, y, z);
        a = Vec3(x[0], x[1], x[2]); b = Vec3(y[0], y[1], y[2]); c = Vec3(z[0], z[1], z[2]);
    }
    void setPeriodicBoxVectors(ContextImpl& context, const Vec3& a, const Vec3& b, const Vec3& c) {
        PlatformData& d = getData(context);
        const double x[3] = {a[0], a[1], a[2]}, y[3] = {b[0], b[1], b[2]}, z[3] = {c[0], c[1], c[2]};
        d.dropForces();
        d.check(b200md_set_box(d.ctx, x, y, z));
    }
    void createCheckpoint(ContextImpl& context, ostream& stream) {
        PlatformData& d = getData(context);
        d.ensureFinalized();
        d.flushCm();
        int64_t n = b200md_checkpoint_save(d.ctx, nullptr, 0);
        vector<char> buf(n);
        if (b200md_checkpoint_save(d.ctx, buf.data(), n) != n) d.check(-1);
        stream.write((const char*) &n, sizeof(n));
        stream.write(buf.data(), n);
    }
    void loadCheckpoint(ContextImpl& context, istream& stream) {
        PlatformData& d = getData(context);
        d.ensureFinalized();
        d.dropForces(); d.cmRequested = false;
        int64_t n = 0;
        stream.read((char*) &n, sizeof(n));
        if (n <= 0 || n != b200md_checkpoint_save(d.ctx, nullptr, 0)) throw OpenMMException("B200 platform: checkpoint does not match this Context");
        vector<char> buf(n);
        stream.read(buf.data(), n);
        d.check(b200md_checkpoint_load(d.ctx, buf.data(), n));
    }
};

class B200ApplyConstraintsKernel : public ApplyConstraintsKernel {
public:
    B200ApplyConstraintsKernel(string name, const Platform& platform) : ApplyConstraintsKernel(name, platform) {}
    void initialize(const System& system) {}
    void apply(ContextImpl& context, double tol) { PlatformData& d = getData(context); d.ensureFinalized(); d.dropForces(); d.check(b200md_apply_constraints(d.ctx, tol)); }
    void applyToVelocities(ContextImpl& context, double tol) { PlatformData& d = getData(context); d.ensureFinalized(); d.flushCm(); d.check(b200md_apply_velocity_constraints(d.ctx, tol)); }
};

class B200VirtualSitesKernel : public VirtualSitesKernel {
public:
    B200VirtualSitesKernel(string name, const Platform& platform) : VirtualSitesKernel(name, platform) {}
    void initialize(const System& system) {
        for (int i = 0; i < system.getNumParticles(); i++)
            if (system.isVirtualSite(i)) throw OpenMMException("B200 platform: virtual sites are not supported");
    }
    void computePositions(ContextImpl& context) {}
};

// ------------------------------------------------------------------------------------------------ forces
class B200CalcNonbondedForceKernel : public CalcNonbondedForceKernel {
public:
    B200CalcNonbondedForceKernel(string name, const Platform& platform, ContextImpl& context) : CalcNonbondedForceKernel(name, platform), context(context), alpha(0) {
        grid[0] = grid[1] = grid[2] = 0;
    }
    // base parameters + offsets (NonbondedForce::addParticleParameterOffset / addExceptionParameterOffset)
    struct Offset { string param; int index; double q, sig, eps; };
    void readForce(const NonbondedForce& force) {
        const int n = force.getNumParticles();
        baseQ.resize(n); baseSig.resize(n); baseEps.resize(n);
        for (int i = 0; i < n; i++) force.getParticleParameters(i, baseQ[i], baseSig[i], baseEps[i]);
        const int ne = force.getNumExceptions();
        ei.resize(ne); ej.resize(ne); baseEqq.resize(ne); baseEsig.resize(ne); baseEeps.resize(ne);
        for (int e = 0; e < ne; e++) force.getExceptionParameters(e, ei[e], ej[e], baseEqq[e], baseEsig[e], baseEeps[e]);
        particleOffsets.clear(); exceptionOffsets.clear(); paramNames.clear();
        for (int i = 0; i < force.getNumParticleParameterOffsets(); i++) {
            Offset o; force.getParticleParameterOffset(i, o.param, o.index, o.q, o.sig, o.eps);
            particleOffsets.push_back(o); paramNames[o.param] = 0.0;
        }
        for (int i = 0; i < force.getNumExceptionParameterOffsets(); i++) {
            Offset o; force.getExceptionParameterOffset(i, o.param, o.index, o.q, o.sig, o.eps);
            exceptionOffsets.push_back(o); paramNames[o.param] = 0.0;
        }
    }
    // computeParameters of the reference (ReferenceKernels.cpp:1077-1121): parameter = base + sum(scale * global value)
    void effective(const map<string, double>& value, vector<double>& q, vector<double>& sig, vector<double>& eps,
                   vector<double>& eqq, vector<double>& esig, vector<double>& eeps) const {
        q = baseQ; sig = baseSig; eps = baseEps; eqq = baseEqq; esig = baseEsig; eeps = baseEeps;
        for (const Offset& o : particleOffsets) { const double v = value.at(o.param); q[o.index] += v*o.q; sig[o.index] += v*o.sig; eps[o.index] += v*o.eps; }
        for (const Offset& o : exceptionOffsets) { const double v = value.at(o.param); eqq[o.index] += v*o.q; esig[o.index] += v*o.sig; eeps[o.index] += v*o.eps; }
    }
    void initialize(const System& system, const NonbondedForce& force) {
        PlatformData& d = getData(context);
        if (d.finalized) throw OpenMMException("B200 platform: NonbondedForce initialised after the Context was finalised");
        b200md_nonbonded_desc nd;
        nd.method = (int) force.getNonbondedMethod();
        if (nd.method == B200MD_NB_EWALD || nd.method == B200MD_NB_LJPME)
            throw OpenMMException("B200 platform: only NoCutoff, CutoffNonPeriodic, CutoffPeriodic and PME are supported");
        nd.cutoff = force.getCutoffDistance();
        nd.use_switch = force.getUseSwitchingFunction() ? 1 : 0;
        nd.switch_distance = force.getSwitchingDistance();
        nd.rf_dielectric = force.getReactionFieldDielectric();
        nd.ewald_alpha = 0; nd.grid[0] = nd.grid[1] = nd.grid[2] = 0;
        if (nd.method == B200MD_NB_PME) {
            NonbondedForceImpl::calcPMEParameters(system, force, alpha, grid[0], grid[1], grid[2], false);
            for (int k = 0; k < 3; k++) grid[k] = fftFriendly(grid[k]);
            nd.ewald_alpha = alpha;
            for (int k = 0; k < 3; k++) nd.grid[k] = grid[k];
        }
        // platform-independent static helper of the reference: call it, don't rewrite it
        nd.dispersion_coefficient = force.getUseDispersionCorrection() ? NonbondedForceImpl::calcDispersionCorrection(system, force) : 0.0;
        nd.exceptions_periodic = force.getExceptionsUsePeriodicBoundaryConditions() ? 1 : 0;
        readForce(force);
        // the Context's parameter map does not exist yet (ContextImpl.cpp:120-131 fills it after the kernels are
        // initialised): start from the defaults, execute() re-uploads whenever a value differs
        for (int i = 0; i < force.getNumGlobalParameters(); i++)
            if (paramNames.count(force.getGlobalParameterName(i))) paramNames[force.getGlobalParameterName(i)] = force.getGlobalParameterDefaultValue(i);
        vector<double> q, sig, eps, eqq, esig, eeps;
        effective(paramNames, q, sig, eps, eqq, esig, eeps);
        d.check(b200md_set_nonbonded(d.ctx, &nd, q.data(), sig.data(), eps.data()));
        if (!ei.empty()) d.check(b200md_set_exceptions(d.ctx, (int) ei.size(), ei.data(), ej.data(), eqq.data(), esig.data(), eeps.data()));
        dispersion = nd.dispersion_coefficient;
        d.systemTerms |= B200MD_TERM_NB_DIRECT | B200MD_TERM_NB_RECIP;
    }
    void upload(PlatformData& d) {
        vector<double> q, sig, eps, eqq, esig, eeps;
        effective(paramNames, q, sig, eps, eqq, esig, eeps);
        d.check(b200md_update_nonbonded_params(d.ctx, q.data(), sig.data(), eps.data(), (int) ei.size(), eqq.data(), esig.data(), eeps.data(), dispersion));
    }
    double execute(ContextImpl& context, bool includeForces, bool includeEnergy, bool includeDirect, bool includeReciprocal) {
        PlatformData& d = getData(context);
        if (!paramNames.empty()) {
            bool changed = false;
            for (auto& pv : paramNames) {
                const double v = context.getParameter(pv.first);
                if (v != pv.second) { pv.second = v; changed = true; }
            }
            if (changed) upload(d);
        }
        if (includeDirect) d.pendingTerms |= B200MD_TERM_NB_DIRECT;
        if (includeReciprocal) d.pendingTerms |= B200MD_TERM_NB_RECIP;
        return 0.0;     // the energy comes back through finishComputation
    }
    void copyParametersToContext(ContextImpl& context, const NonbondedForce& force) {
        PlatformData& d = getData(context);
        d.ensureFinalized();
        d.dropForces();
        if (force.getNumParticles() != (int) baseQ.size()) throw OpenMMException("updateParametersInContext: The number of particles has changed");
        if (force.getNumExceptions() != (int) ei.size()) throw OpenMMException("updateParametersInContext: The number of exceptions has changed");
        const vector<int> oi = ei, oj = ej;
        const map<string, double> old = paramNames;
        readForce(force);
        for (size_t e = 0; e < ei.size(); e++)
            if (ei[e] != oi[e] || ej[e] != oj[e]) throw OpenMMException("updateParametersInContext: The set of particles in an exception has changed");
        for (auto& pv : paramNames) pv.second = context.getParameter(pv.first);
        dispersion = force.getUseDispersionCorrection() ? NonbondedForceImpl::calcDispersionCorrection(context.getSystem(), force) : 0.0;
        upload(d);
    }
    void getPMEParameters(double& a, int& nx, int& ny, int& nz) const { a = alpha; nx = grid[0]; ny = grid[1]; nz = grid[2]; }
    void getLJPMEParameters(double& a, int& nx, int& ny, int& nz) const { throw OpenMMException("B200 platform: LJPME is not supported"); }
private:
    ContextImpl& context;
    double alpha, dispersion = 0;
    int grid[3];
    vector<double> baseQ, baseSig, baseEps, baseEqq, baseEsig, baseEeps;
    vector<int> ei, ej;
    vector<Offset> particleOffsets, exceptionOffsets;
    map<string, double> paramNames;      // global parameters used by offsets -> value the device parameters were computed with
};

// Bonded forces: any number of Force objects per class.  Every object appends its terms (tagged with its force group) to
// the per-Context arrays; the engine evaluates a term iff its class was executed in this evaluation AND its group is in
// the `groups` mask of finishComputation (ForceImpl::calcForcesAndEnergy only calls execute for objects whose group is in it).
class B200CalcHarmonicBondForceKernel : public CalcHarmonicBondForceKernel {
public:
    B200CalcHarmonicBondForceKernel(string name, const Platform& platform, ContextImpl& context) : CalcHarmonicBondForceKernel(name, platform), context(context) {}
    void initialize(const System& system, const HarmonicBondForce& force) {
        PlatformData& d = getData(context);
        if (d.finalized) throw OpenMMException("B200 platform: HarmonicBondForce initialised after the Context was finalised");
        first = (int) d.bondI.size(); count = force.getNumBonds();
        for (int i = 0; i < count; i++) {
            int a, b; double r0, k;
            force.getBondParameters(i, a, b, r0, k);
            d.bondI.push_back(a); d.bondJ.push_back(b); d.bondR0.push_back(r0); d.bondK.push_back(k); d.bondG.push_back(force.getForceGroup() | (force.usesPeriodicBoundaryConditions() ? 0x80 : 0));
        }
        d.systemTerms |= B200MD_TERM_BONDS; d.bondedGroupsUsed |= 1u << force.getForceGroup();
    }
    double execute(ContextImpl& context, bool includeForces, bool includeEnergy) { getData(context).pendingTerms |= B200MD_TERM_BONDS; return 0.0; }
    void copyParametersToContext(ContextImpl& context, const HarmonicBondForce& force) {
        PlatformData& d = getData(context);
        d.ensureFinalized();
        d.dropForces();
        if (force.getNumBonds() != count) throw OpenMMException("updateParametersInContext: The number of bonds has changed");
        for (int i = 0; i < count; i++) {
            int p, q;
            force.getBondParameters(i, p, q, d.bondR0[first+i], d.bondK[first+i]);
            if (p != d.bondI[first+i] || q != d.bondJ[first+i]) throw OpenMMException("updateParametersInContext: The set of particles in a bond has changed");
        }
        d.check(b200md_update_bonded_params(d.ctx, 0, (int) d.bondI.size(), d.bondR0.data(), d.bondK.data(), nullptr));
    }
private:
    ContextImpl& context;
    int first = 0, count = 0;
};

class B200CalcHarmonicAngleForceKernel : public CalcHarmonicAngleForceKernel {
public:
    B200CalcHarmonicAngleForceKernel(string name, const Platform& platform, ContextImpl& context) : CalcHarmonicAngleForceKernel(name, platform), context(context) {}
    void initialize(const System& system, const HarmonicAngleForce& force) {
        PlatformData& d = getData(context);
        if (d.finalized) throw OpenMMException("B200 platform: HarmonicAngleForce initialised after the Context was finalised");
        first = (int) d.angI.size(); count = force.getNumAngles();
        for (int i = 0; i < count; i++) {
            int a, b, c; double t0, k;
            force.getAngleParameters(i, a, b, c, t0, k);
            d.angI.push_back(a); d.angJ.push_back(b); d.angK.push_back(c); d.angT0.push_back(t0); d.angKK.push_back(k); d.angG.push_back(force.getForceGroup() | (force.usesPeriodicBoundaryConditions() ? 0x80 : 0));
        }
        d.systemTerms |= B200MD_TERM_ANGLES; d.bondedGroupsUsed |= 1u << force.getForceGroup();
    }
    double execute(ContextImpl& context, bool includeForces, bool includeEnergy) { getData(context).pendingTerms |= B200MD_TERM_ANGLES; return 0.0; }
    void copyParametersToContext(ContextImpl& context, const HarmonicAngleForce& force) {
        PlatformData& d = getData(context);
        d.ensureFinalized();
        d.dropForces();
        if (force.getNumAngles() != count) throw OpenMMException("updateParametersInContext: The number of angles has changed");
        for (int i = 0; i < count; i++) {
            int p, q, r;
            force.getAngleParameters(i, p, q, r, d.angT0[first+i], d.angKK[first+i]);
            if (p != d.angI[first+i] || q != d.angJ[first+i] || r != d.angK[first+i]) throw OpenMMException("updateParametersInContext: The set of particles in an angle has changed");
        }
        d.check(b200md_update_bonded_params(d.ctx, 1, (int) d.angI.size(), d.angT0.data(), d.angKK.data(), nullptr));
    }
private:
    ContextImpl& context;
    int first = 0, count = 0;
};

class B200CalcPeriodicTorsionForceKernel : public CalcPeriodicTorsionForceKernel {
public:
    B200CalcPeriodicTorsionForceKernel(string name, const Platform& platform, ContextImpl& context) : CalcPeriodicTorsionForceKernel(name, platform), context(context) {}
    void initialize(const System& system, const PeriodicTorsionForce& force) {
        PlatformData& d = getData(context);
        if (d.finalized) throw OpenMMException("B200 platform: PeriodicTorsionForce initialised after the Context was finalised");
        first = (int) d.torI.size(); count = force.getNumTorsions();
        for (int i = 0; i < count; i++) {
            int a, b, c, e, n; double phase, k;
            force.getTorsionParameters(i, a, b, c, e, n, phase, k);
            d.torI.push_back(a); d.torJ.push_back(b); d.torK.push_back(c); d.torL.push_back(e); d.torN.push_back(n); d.torPhase.push_back(phase); d.torKK.push_back(k);
            d.torG.push_back(force.getForceGroup() | (force.usesPeriodicBoundaryConditions() ? 0x80 : 0));
        }
        d.systemTerms |= B200MD_TERM_TORSIONS; d.bondedGroupsUsed |= 1u << force.getForceGroup();
    }
    double execute(ContextImpl& context, bool includeForces, bool includeEnergy) { getData(context).pendingTerms |= B200MD_TERM_TORSIONS; return 0.0; }
    void copyParametersToContext(ContextImpl& context, const PeriodicTorsionForce& force) {
        PlatformData& d = getData(context);
        d.ensureFinalized();
        d.dropForces();
        if (force.getNumTorsions() != count) throw OpenMMException("updateParametersInContext: The number of torsions has changed");
        for (int i = 0; i < count; i++) {
            int p, q, r, t;
            force.getTorsionParameters(i, p, q, r, t, d.torN[first+i], d.torPhase[first+i], d.torKK[first+i]);
            if (p != d.torI[first+i] || q != d.torJ[first+i] || r != d.torK[first+i] || t != d.torL[first+i]) throw OpenMMException("updateParametersInContext: The set of particles in a torsion has changed");
        }
        d.check(b200md_update_bonded_params(d.ctx, 2, (int) d.torI.size(), d.torPhase.data(), d.torKK.data(), d.torN.data()));
    }
private:
    ContextImpl& context;
    int first = 0, count = 0;
};

// RemoveCMMotionKernel::execute is called by CMMotionRemoverImpl::updateContextState in EVERY step; the frequency test is
// the kernel's (ReferenceKernels.cpp:2712-2714).  The removal itself is deferred to the integrator step that follows (the
// fused step graph removes the centre-of-mass motion itself) or to the next read of the velocities (PlatformData::flushCm).
class B200RemoveCMMotionKernel : public RemoveCMMotionKernel {
public:
    B200RemoveCMMotionKernel(string name, const Platform& platform, ContextImpl& context) : RemoveCMMotionKernel(name, platform), context(context) {}
    void initialize(const System& system, const CMMotionRemover& force) {
        PlatformData& d = getData(context);
        d.cmFrequency = force.getFrequency();
        d.check(b200md_set_cm_remover(d.ctx, d.cmFrequency));
    }
    void execute(ContextImpl& context) { PlatformData& d = getData(context); d.ensureFinalized(); d.cmRequested = true; }
private:
    ContextImpl& context;
};

// ------------------------------------------------------------------------------------------------ integrators
// The reference's Integrator::step drives updateContextState -> calcForcesAndEnergy -> kernel.execute per step
// (LangevinIntegrator.cpp:74-82); execute() is the integrate+constrain half, enqueued with no host sync.
void configureIntegrator(ContextImpl& context, int kind, double dt, double temperature, double friction, int seed, double tol) {
    PlatformData& d = getData(context);
    if (d.integratorKind == kind && d.dt == dt && d.temperature == temperature && d.friction == friction && d.tol == tol && d.seed == seed) return;
    d.check(b200md_set_integrator(d.ctx, kind, dt, temperature, friction, seed, tol));
    d.integratorKind = kind; d.dt = dt; d.temperature = temperature; d.friction = friction; d.tol = tol; d.seed = seed;
}

// The integrate + constrain half of a step.  When the forces-only evaluation that precedes it in Integrator::step is still
// pending and covered the whole force field, forces + integration run as ONE replay of the captured step graph.
void integrateStep(ContextImpl& context) {
    PlatformData& d = getData(context);
    d.ensureFinalized();
    const bool whole = d.lazyForces && d.lazyTerms == d.systemTerms && (d.bondedGroupsUsed & ~d.lazyGroups) == 0;
    if (whole && d.useFusedStep) {
        d.lazyForces = false; d.cmRequested = false;         // b200md_step removes the centre-of-mass motion at its own frequency
        d.check(b200md_step(d.ctx, 1));
        return;
    }
    d.flushCm();
    d.flushForces();
    d.check(b200md_integrate_only(d.ctx));
}
double kineticEnergy(ContextImpl& context) {
    PlatformData& d = getData(context);
    d.ensureFinalized();
    d.flushCm();
    d.flushForces();         // the leapfrog integrators report the kinetic energy at a half-step-shifted velocity (ReferenceKernels.cpp:146-176)
    double ke = 0; d.check(b200md_kinetic_energy(d.ctx, &ke)); return ke;
}

class B200IntegrateVerletStepKernel : public IntegrateVerletStepKernel {
public:
    B200IntegrateVerletStepKernel(string name, const Platform& platform) : IntegrateVerletStepKernel(name, platform) {}
    void initialize(const System& system, const VerletIntegrator& integrator) {}
    void configure(ContextImpl& context, const VerletIntegrator& in) { configureIntegrator(context, B200MD_INT_VERLET, in.getStepSize(), 0.0, 0.0, 0, in.getConstraintTolerance()); }
    void execute(ContextImpl& context, const VerletIntegrator& integrator) { configure(context, integrator); integrateStep(context); }
    double computeKineticEnergy(ContextImpl& context, const VerletIntegrator& integrator) { configure(context, integrator); return kineticEnergy(context); }
};

class B200IntegrateLangevinStepKernel : public IntegrateLangevinStepKernel {
public:
    B200IntegrateLangevinStepKernel(string name, const Platform& platform) : IntegrateLangevinStepKernel(name, platform) {}
    void initialize(const System& system, const LangevinIntegrator& integrator) {}
    void configure(ContextImpl& context, const LangevinIntegrator& in) {
        configureIntegrator(context, B200MD_INT_LANGEVIN, in.getStepSize(), in.getTemperature(), in.getFriction(), in.getRandomNumberSeed(), in.getConstraintTolerance());
    }
    void execute(ContextImpl& context, const LangevinIntegrator& integrator) { configure(context, integrator); integrateStep(context); }
    double computeKineticEnergy(ContextImpl& context, const LangevinIntegrator& integrator) { configure(context, integrator); return kineticEnergy(context); }
};

class B200IntegrateLangevinMiddleStepKernel : public IntegrateLangevinMiddleStepKernel {
public:
    B200IntegrateLangevinMiddleStepKernel(string name, const Platform& platform) : IntegrateLangevinMiddleStepKernel(name, platform) {}
    void initialize(const System& system, const LangevinMiddleIntegrator& integrator) {}
    void configure(ContextImpl& context, const LangevinMiddleIntegrator& in) {
        configureIntegrator(context, B200MD_INT_LANGEVIN_MIDDLE, in.getStepSize(), in.getTemperature(), in.getFriction(), in.getRandomNumberSeed(), in.getConstraintTolerance());
    }
    void execute(ContextImpl& context, const LangevinMiddleIntegrator& integrator) { configure(context, integrator); integrateStep(context); }
    double computeKineticEnergy(ContextImpl& context, const LangevinMiddleIntegrator& integrator) { configure(context, integrator); return kineticEnergy(context); }
};

// ------------------------------------------------------------------------------------------------ factory + platform
class B200KernelFactory : public KernelFactory {
public:
    KernelImpl* createKernelImpl(string name, const Platform& platform, ContextImpl& context) const {
        if (name == CalcForcesAndEnergyKernel::Name()) return new B200CalcForcesAndEnergyKernel(name, platform);
        if (name == UpdateStateDataKernel::Name()) return new B200UpdateStateDataKernel(name, platform);
        if (name == ApplyConstraintsKernel::Name()) return new B200ApplyConstraintsKernel(name, platform);
        if (name == VirtualSitesKernel::Name()) return new B200VirtualSitesKernel(name, platform);
        if (name == CalcNonbondedForceKernel::Name()) return new B200CalcNonbondedForceKernel(name, platform, context);
        if (name == CalcHarmonicBondForceKernel::Name()) return new B200CalcHarmonicBondForceKernel(name, platform, context);
        if (name == CalcHarmonicAngleForceKernel::Name()) return new B200CalcHarmonicAngleForceKernel(name, platform, context);
        if (name == CalcPeriodicTorsionForceKernel::Name()) return new B200CalcPeriodicTorsionForceKernel(name, platform, context);
        if (name == RemoveCMMotionKernel::Name()) return new B200RemoveCMMotionKernel(name, platform, context);
        if (name == IntegrateVerletStepKernel::Name()) return new B200IntegrateVerletStepKernel(name, platform);
        if (name == IntegrateLangevinStepKernel::Name()) return new B200IntegrateLangevinStepKernel(name, platform);
        if (name == IntegrateLangevinMiddleStepKernel::Name()) return new B200IntegrateLangevinMiddleStepKernel(name, platform);
        throw OpenMMException((string("Tried to create kernel with illegal kernel name '") + name + "'").c_str());
    }
};

class B200Platform : public Platform {
public:
    B200Platform() {
        B200KernelFactory* factory = new B200KernelFactory();
        for (const string& n : {CalcForcesAndEnergyKernel::Name(), UpdateStateDataKernel::Name(), ApplyConstraintsKernel::Name(), VirtualSitesKernel::Name(),
                                CalcNonbondedForceKernel::Name(), CalcHarmonicBondForceKernel::Name(), CalcHarmonicAngleForceKernel::Name(),
                                CalcPeriodicTorsionForceKernel::Name(), RemoveCMMotionKernel::Name(), IntegrateVerletStepKernel::Name(),
                                IntegrateLangevinStepKernel::Name(), IntegrateLangevinMiddleStepKernel::Name()})
            registerKernelFactory(n, factory);
        platformProperties.push_back(DeviceIndex());
        platformProperties.push_back(Precision());
        setPropertyDefaultValue(DeviceIndex(), "0");
        setPropertyDefaultValue(Precision(), "single");
    }
    static const string& DeviceIndex() { static const string key = "DeviceIndex"; return key; }
    static const string& Precision() { static const string key = "Precision"; return key; }
    const string& getName() const { static const string name = "B200"; return name; }
    double getSpeed() const { return 200; }        // CUDA = 100 (CudaPlatform.cpp:149-151): win auto-selection; what this platform cannot run is refused in contextCreated
    bool supportsDoublePrecision() const { return false; }
    const string& getPropertyValue(const Context& context, const string& property) const {
        const ContextImpl& impl = getContextImpl(context);
        const PlatformData& d = getData(impl);
        map<string, string>::const_iterator it = d.props.find(property);
        if (it != d.props.end()) return it->second;
        return Platform::getPropertyValue(context, property);
    }
    void setPropertyValue(Context& context, const string& property, const string& value) const {
        // no property of this platform can change once the Context exists (DeviceIndex and Precision are fixed at creation)
        throw OpenMMException("B200 platform: property '" + property + "' cannot be changed after the Context was created");
    }
    // Everything the kernels would refuse LATER must be refused HERE: ContextImpl only falls back to the next platform
    // when contextCreated() throws (ContextImpl.cpp:152-166); a throw from Kernel::initialize or from the first
    // setPositions would instead make Context creation fail for a System that CUDA / CPU can run.
    static void validateSystem(const System& system) {
        for (int i = 0; i < system.getNumParticles(); i++)
            if (system.isVirtualSite(i)) throw OpenMMException("B200 platform: virtual sites are not supported");
        int numNonbonded = 0;
        for (int f = 0; f < system.getNumForces(); f++) {
            const Force& force = system.getForce(f);
            if (const NonbondedForce* nb = dynamic_cast<const NonbondedForce*>(&force)) {
                if (++numNonbonded > 1) throw OpenMMException("B200 platform: only one NonbondedForce per System is supported");
                const NonbondedForce::NonbondedMethod m = nb->getNonbondedMethod();
                if (m == NonbondedForce::Ewald || m == NonbondedForce::LJPME)
                    throw OpenMMException("B200 platform: only NoCutoff, CutoffNonPeriodic, CutoffPeriodic and PME are supported");
                if (nb->getNumParticles() != system.getNumParticles()) throw OpenMMException("NonbondedForce must have exactly as many particles as the System it belongs to.");
            }
            if (force.getForceGroup() < 0 || force.getForceGroup() > 31) throw OpenMMException("B200 platform: force group out of range");
        }
        const int nc = system.getNumConstraints();
        if (nc > 0) {
            vector<int> ci(nc), cj(nc); vector<double> cd(nc), mass(system.getNumParticles());
            for (int k = 0; k < nc; k++) system.getConstraintParameters(k, ci[k], cj[k], cd[k]);
            for (int i = 0; i < system.getNumParticles(); i++) mass[i] = system.getParticleMass(i);
            char msg[256];
            if (b200md_check_constraints(system.getNumParticles(), mass.data(), nc, ci.data(), cj.data(), cd.data(), msg, sizeof(msg)) != 0)
                throw OpenMMException(string("B200 platform: ") + msg);
        }
    }
    void contextCreated(ContextImpl& context, const map<string, string>& properties) const {
        validateSystem(context.getSystem());
        PlatformData* d = new PlatformData();
        if (getenv("B200MD_PLUGIN_FUSED")) d->useFusedStep = atoi(getenv("B200MD_PLUGIN_FUSED")) != 0;
        try {
            string dev = properties.count(DeviceIndex()) ? properties.at(DeviceIndex()) : getPropertyDefaultValue(DeviceIndex());
            string prec = properties.count(Precision()) ? properties.at(Precision()) : getPropertyDefaultValue(Precision());
            if (prec != "single") throw OpenMMException("B200 platform: only Precision=single is implemented");
            const System& system = context.getSystem();
            d->numParticles = system.getNumParticles();
            if (b200md_create(&d->ctx, atoi(dev.c_str()), d->numParticles) != 0)
                throw OpenMMException(string("B200 platform: ") + b200md_last_error(nullptr));
            d->props[DeviceIndex()] = dev;
            d->props[Precision()] = prec;
            vector<double> mass(d->numParticles);
            for (int i = 0; i < d->numParticles; i++) mass[i] = system.getParticleMass(i);
            d->check(b200md_set_masses(d->ctx, mass.data()));
            const int nc = system.getNumConstraints();
            if (nc > 0) {
                vector<int> ci(nc), cj(nc); vector<double> cd(nc);
                for (int k = 0; k < nc; k++) system.getConstraintParameters(k, ci[k], cj[k], cd[k]);
                d->check(b200md_set_constraints(d->ctx, nc, ci.data(), cj.data(), cd.data()));
            }
        } catch (...) {
            if (d->ctx) b200md_destroy(d->ctx);
            delete d;
            throw;
        }
        context.setPlatformData(d);
    }
    void contextDestroyed(ContextImpl& context) const {
        PlatformData* d = reinterpret_cast<PlatformData*>(context.getPlatformData());
        if (d) { if (d->ctx) b200md_destroy(d->ctx); delete d; }
    }
};

} // namespace

extern "C" __attribute__((visibility("default"))) void registerPlatforms() {
    Platform::registerPlatform(new B200Platform());
}

extern "C" __attribute__((visibility("default"))) void registerKernelFactories() {
}
