// B200Platform.cpp -- libOpenMMB200.so: the OpenMM Platform plugin for the B200-native hot path.
//
// A thin C++ adapter: every KernelImpl below forwards one abstract kernel interface of the reference
// (olla/include/openmm/kernels.h) to the C-ABI of libb200md.so (include/b200md.h), where all the CUDA lives.
// Loaded by Platform::loadPluginLibrary / loadPluginsFromDirectory (olla/src/Platform.cpp:221-320) through the
// extern "C" registerPlatforms() entry point (PluginInitializer.h:47), exactly like platforms/cuda
// (CudaPlatform.cpp:59-61).  Compiled against the reference's headers where they lie; no reference source is copied.
//
// Supported: NonbondedForce (NoCutoff, CutoffNonPeriodic, CutoffPeriodic, PME), HarmonicBondForce, HarmonicAngleForce,
// PeriodicTorsionForce, CMMotionRemover; Verlet / Langevin / LangevinMiddle integrators; SETTLE + X-H_n SHAKE
// constraints.  Anything else makes Platform::supportsKernels() false, so ContextImpl picks another platform.
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
        check(b200md_finalize(ctx));
        finalized = true;
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
        d.pendingTerms = 0;
        d.includeEnergy = includeEnergy;
    }
    double finishComputation(ContextImpl& context, bool includeForce, bool includeEnergy, int groups, bool& valid) {
        // every Calc*ForceKernel::execute only recorded its term; one engine call evaluates them all on the device
        PlatformData& d = getData(context);
        double energy = 0;
        d.check(b200md_compute(d.ctx, d.pendingTerms, includeForce ? 1 : 0, includeEnergy ? &energy : nullptr));
        valid = true;
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
        vector<double> x(3*d.numParticles);
        for (int i = 0; i < d.numParticles; i++) for (int k = 0; k < 3; k++) x[3*i+k] = positions[i][k];
        d.check(b200md_set_positions(d.ctx, x.data()));
    }
    void getVelocities(ContextImpl& context, vector<Vec3>& velocities) {
        PlatformData& d = getData(context);
        d.ensureFinalized();
        vector<double> x(3*d.numParticles);
        d.check(b200md_get_velocities(d.ctx, x.data()));
        velocities.resize(d.numParticles);
        for (int i = 0; i < d.numParticles; i++) velocities[i] = Vec3(x[3*i], x[3*i+1], x[3*i+2]);
    }
    void setVelocities(ContextImpl& context, const vector<Vec3>& velocities) {
        PlatformData& d = getData(context);
        d.ensureFinalized();
        vector<double> x(3*d.numParticles);
        for (int i = 0; i < d.numParticles; i++) for (int k = 0; k < 3; k++) x[3*i+k] = velocities[i][k];
        d.check(b200md_set_velocities(d.ctx, x.data()));
    }
    void getForces(ContextImpl& context, vector<Vec3>& forces) {
        PlatformData& d = getData(context);
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
        d.check(b200md_set_box(d.ctx, x, y, z));
    }
    void createCheckpoint(ContextImpl& context, ostream& stream) {
        PlatformData& d = getData(context);
        d.ensureFinalized();
        int64_t n = b200md_checkpoint_save(d.ctx, nullptr, 0);
        vector<char> buf(n);
        if (b200md_checkpoint_save(d.ctx, buf.data(), n) != n) d.check(-1);
        stream.write((const char*) &n, sizeof(n));
        stream.write(buf.data(), n);
    }
    void loadCheckpoint(ContextImpl& context, istream& stream) {
        PlatformData& d = getData(context);
        d.ensureFinalized();
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
    void apply(ContextImpl& context, double tol) { PlatformData& d = getData(context); d.ensureFinalized(); d.check(b200md_apply_constraints(d.ctx, tol)); }
    void applyToVelocities(ContextImpl& context, double tol) { PlatformData& d = getData(context); d.ensureFinalized(); d.check(b200md_apply_velocity_constraints(d.ctx, tol)); }
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
    void gather(const NonbondedForce& force, vector<double>& q, vector<double>& sig, vector<double>& eps,
                vector<int>& ei, vector<int>& ej, vector<double>& eqq, vector<double>& esig, vector<double>& eeps) {
        const int n = force.getNumParticles();
        q.resize(n); sig.resize(n); eps.resize(n);
        for (int i = 0; i < n; i++) force.getParticleParameters(i, q[i], sig[i], eps[i]);
        const int ne = force.getNumExceptions();
        ei.resize(ne); ej.resize(ne); eqq.resize(ne); esig.resize(ne); eeps.resize(ne);
        for (int e = 0; e < ne; e++) force.getExceptionParameters(e, ei[e], ej[e], eqq[e], esig[e], eeps[e]);
    }
    void initialize(const System& system, const NonbondedForce& force) {
        PlatformData& d = getData(context);
        if (d.finalized) throw OpenMMException("B200 platform: NonbondedForce initialised after the Context was finalised");
        if (force.getNumParticleParameterOffsets() > 0 || force.getNumExceptionParameterOffsets() > 0)
            throw OpenMMException("B200 platform: NonbondedForce parameter offsets are not supported");
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
        vector<double> q, sig, eps, eqq, esig, eeps; vector<int> ei, ej;
        gather(force, q, sig, eps, ei, ej, eqq, esig, eeps);
        d.check(b200md_set_nonbonded(d.ctx, &nd, q.data(), sig.data(), eps.data()));
        if (!ei.empty()) d.check(b200md_set_exceptions(d.ctx, (int) ei.size(), ei.data(), ej.data(), eqq.data(), esig.data(), eeps.data()));
    }
    double execute(ContextImpl& context, bool includeForces, bool includeEnergy, bool includeDirect, bool includeReciprocal) {
        PlatformData& d = getData(context);
        if (includeDirect) d.pendingTerms |= B200MD_TERM_NB_DIRECT;
        if (includeReciprocal) d.pendingTerms |= B200MD_TERM_NB_RECIP;
        return 0.0;     // the energy comes back through finishComputation
    }
    void copyParametersToContext(ContextImpl& context, const NonbondedForce& force) {
        PlatformData& d = getData(context);
        d.ensureFinalized();
        vector<double> q, sig, eps, eqq, esig, eeps; vector<int> ei, ej;
        gather(force, q, sig, eps, ei, ej, eqq, esig, eeps);
        double disp = force.getUseDispersionCorrection() ? NonbondedForceImpl::calcDispersionCorrection(context.getSystem(), force) : 0.0;
        d.check(b200md_update_nonbonded_params(d.ctx, q.data(), sig.data(), eps.data(), (int) ei.size(), eqq.data(), esig.data(), eeps.data(), disp));
    }
    void getPMEParameters(double& a, int& nx, int& ny, int& nz) const { a = alpha; nx = grid[0]; ny = grid[1]; nz = grid[2]; }
    void getLJPMEParameters(double& a, int& nx, int& ny, int& nz) const { throw OpenMMException("B200 platform: LJPME is not supported"); }
private:
    ContextImpl& context;
    double alpha;
    int grid[3];
};

class B200CalcHarmonicBondForceKernel : public CalcHarmonicBondForceKernel {
public:
    B200CalcHarmonicBondForceKernel(string name, const Platform& platform, ContextImpl& context) : CalcHarmonicBondForceKernel(name, platform), context(context) {}
    void initialize(const System& system, const HarmonicBondForce& force) {
        PlatformData& d = getData(context);
        if (force.usesPeriodicBoundaryConditions()) throw OpenMMException("B200 platform: periodic bonded forces are not supported");
        for (int i = 0; i < force.getNumBonds(); i++) {
            int a, b; double r0, k;
            force.getBondParameters(i, a, b, r0, k);
            d.bondI.push_back(a); d.bondJ.push_back(b); d.bondR0.push_back(r0); d.bondK.push_back(k);
        }
    }
    double execute(ContextImpl& context, bool includeForces, bool includeEnergy) { getData(context).pendingTerms |= B200MD_TERM_BONDS; return 0.0; }
    void copyParametersToContext(ContextImpl& context, const HarmonicBondForce& force) {
        PlatformData& d = getData(context);
        d.ensureFinalized();
        const int n = force.getNumBonds();
        if (n != (int) d.bondI.size()) throw OpenMMException("B200 platform: updateParametersInContext needs exactly one HarmonicBondForce with an unchanged number of bonds");
        vector<double> a(n), b(n);
        for (int i = 0; i < n; i++) { int p, q; force.getBondParameters(i, p, q, a[i], b[i]); if (p != d.bondI[i] || q != d.bondJ[i]) throw OpenMMException("updateParametersInContext: The set of particles in a bond has changed"); }
        d.check(b200md_update_bonded_params(d.ctx, 0, n, a.data(), b.data(), nullptr));
    }
private:
    ContextImpl& context;
};

class B200CalcHarmonicAngleForceKernel : public CalcHarmonicAngleForceKernel {
public:
    B200CalcHarmonicAngleForceKernel(string name, const Platform& platform, ContextImpl& context) : CalcHarmonicAngleForceKernel(name, platform), context(context) {}
    void initialize(const System& system, const HarmonicAngleForce& force) {
        PlatformData& d = getData(context);
        if (force.usesPeriodicBoundaryConditions()) throw OpenMMException("B200 platform: periodic bonded forces are not supported");
        for (int i = 0; i < force.getNumAngles(); i++) {
            int a, b, c; double t0, k;
            force.getAngleParameters(i, a, b, c, t0, k);
            d.angI.push_back(a); d.angJ.push_back(b); d.angK.push_back(c); d.angT0.push_back(t0); d.angKK.push_back(k);
        }
    }
    double execute(ContextImpl& context, bool includeForces, bool includeEnergy) { getData(context).pendingTerms |= B200MD_TERM_ANGLES; return 0.0; }
    void copyParametersToContext(ContextImpl& context, const HarmonicAngleForce& force) {
        PlatformData& d = getData(context);
        d.ensureFinalized();
        const int n = force.getNumAngles();
        if (n != (int) d.angI.size()) throw OpenMMException("B200 platform: updateParametersInContext needs exactly one HarmonicAngleForce with an unchanged number of angles");
        vector<double> a(n), b(n);
        for (int i = 0; i < n; i++) { int p, q, r; force.getAngleParameters(i, p, q, r, a[i], b[i]); if (p != d.angI[i] || q != d.angJ[i] || r != d.angK[i]) throw OpenMMException("updateParametersInContext: The set of particles in an angle has changed"); }
        d.check(b200md_update_bonded_params(d.ctx, 1, n, a.data(), b.data(), nullptr));
    }
private:
    ContextImpl& context;
};

class B200CalcPeriodicTorsionForceKernel : public CalcPeriodicTorsionForceKernel {
public:
    B200CalcPeriodicTorsionForceKernel(string name, const Platform& platform, ContextImpl& context) : CalcPeriodicTorsionForceKernel(name, platform), context(context) {}
    void initialize(const System& system, const PeriodicTorsionForce& force) {
        PlatformData& d = getData(context);
        if (force.usesPeriodicBoundaryConditions()) throw OpenMMException("B200 platform: periodic bonded forces are not supported");
        for (int i = 0; i < force.getNumTorsions(); i++) {
            int a, b, c, e, n; double phase, k;
            force.getTorsionParameters(i, a, b, c, e, n, phase, k);
            d.torI.push_back(a); d.torJ.push_back(b); d.torK.push_back(c); d.torL.push_back(e); d.torN.push_back(n); d.torPhase.push_back(phase); d.torKK.push_back(k);
        }
    }
    double execute(ContextImpl& context, bool includeForces, bool includeEnergy) { getData(context).pendingTerms |= B200MD_TERM_TORSIONS; return 0.0; }
    void copyParametersToContext(ContextImpl& context, const PeriodicTorsionForce& force) {
        PlatformData& d = getData(context);
        d.ensureFinalized();
        const int n = force.getNumTorsions();
        if (n != (int) d.torI.size()) throw OpenMMException("B200 platform: updateParametersInContext needs exactly one PeriodicTorsionForce with an unchanged number of torsions");
        vector<double> a(n), b(n); vector<int> per(n);
        for (int i = 0; i < n; i++) { int p, q, r, t; force.getTorsionParameters(i, p, q, r, t, per[i], a[i], b[i]); if (p != d.torI[i] || q != d.torJ[i] || r != d.torK[i] || t != d.torL[i]) throw OpenMMException("updateParametersInContext: The set of particles in a torsion has changed"); }
        d.check(b200md_update_bonded_params(d.ctx, 2, n, a.data(), b.data(), per.data()));
    }
private:
    ContextImpl& context;
};

class B200RemoveCMMotionKernel : public RemoveCMMotionKernel {
public:
    B200RemoveCMMotionKernel(string name, const Platform& platform) : RemoveCMMotionKernel(name, platform) {}
    void initialize(const System& system, const CMMotionRemover& force) {}
    void execute(ContextImpl& context) { PlatformData& d = getData(context); d.ensureFinalized(); d.check(b200md_remove_cm_motion(d.ctx)); }
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

class B200IntegrateVerletStepKernel : public IntegrateVerletStepKernel {
public:
    B200IntegrateVerletStepKernel(string name, const Platform& platform) : IntegrateVerletStepKernel(name, platform) {}
    void initialize(const System& system, const VerletIntegrator& integrator) {}
    void execute(ContextImpl& context, const VerletIntegrator& integrator) {
        PlatformData& d = getData(context);
        configureIntegrator(context, B200MD_INT_VERLET, integrator.getStepSize(), 0.0, 0.0, 0, integrator.getConstraintTolerance());
        d.check(b200md_integrate_only(d.ctx));
    }
    double computeKineticEnergy(ContextImpl& context, const VerletIntegrator& integrator) {
        PlatformData& d = getData(context);
        configureIntegrator(context, B200MD_INT_VERLET, integrator.getStepSize(), 0.0, 0.0, 0, integrator.getConstraintTolerance());
        double ke = 0; d.check(b200md_kinetic_energy(d.ctx, &ke)); return ke;
    }
};

class B200IntegrateLangevinStepKernel : public IntegrateLangevinStepKernel {
public:
    B200IntegrateLangevinStepKernel(string name, const Platform& platform) : IntegrateLangevinStepKernel(name, platform) {}
    void initialize(const System& system, const LangevinIntegrator& integrator) {}
    void configure(ContextImpl& context, const LangevinIntegrator& in) {
        configureIntegrator(context, B200MD_INT_LANGEVIN, in.getStepSize(), in.getTemperature(), in.getFriction(), in.getRandomNumberSeed(), in.getConstraintTolerance());
    }
    void execute(ContextImpl& context, const LangevinIntegrator& integrator) {
        PlatformData& d = getData(context);
        configure(context, integrator);
        d.check(b200md_integrate_only(d.ctx));
    }
    double computeKineticEnergy(ContextImpl& context, const LangevinIntegrator& integrator) {
        PlatformData& d = getData(context);
        configure(context, integrator);
        double ke = 0; d.check(b200md_kinetic_energy(d.ctx, &ke)); return ke;
    }
};

class B200IntegrateLangevinMiddleStepKernel : public IntegrateLangevinMiddleStepKernel {
public:
    B200IntegrateLangevinMiddleStepKernel(string name, const Platform& platform) : IntegrateLangevinMiddleStepKernel(name, platform) {}
    void initialize(const System& system, const LangevinMiddleIntegrator& integrator) {}
    void configure(ContextImpl& context, const LangevinMiddleIntegrator& in) {
        configureIntegrator(context, B200MD_INT_LANGEVIN_MIDDLE, in.getStepSize(), in.getTemperature(), in.getFriction(), in.getRandomNumberSeed(), in.getConstraintTolerance());
    }
    void execute(ContextImpl& context, const LangevinMiddleIntegrator& integrator) {
        PlatformData& d = getData(context);
        configure(context, integrator);
        d.check(b200md_integrate_only(d.ctx));
    }
    double computeKineticEnergy(ContextImpl& context, const LangevinMiddleIntegrator& integrator) {
        PlatformData& d = getData(context);
        configure(context, integrator);
        double ke = 0; d.check(b200md_kinetic_energy(d.ctx, &ke)); return ke;
    }
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
        if (name == RemoveCMMotionKernel::Name()) return new B200RemoveCMMotionKernel(name, platform);
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
    double getSpeed() const { return 200; }        // CUDA = 100 (CudaPlatform.cpp:149-151): win auto-selection
    bool supportsDoublePrecision() const { return false; }
    const string& getPropertyValue(const Context& context, const string& property) const {
        const ContextImpl& impl = getContextImpl(context);
        const PlatformData& d = getData(impl);
        map<string, string>::const_iterator it = d.props.find(property);
        if (it != d.props.end()) return it->second;
        return Platform::getPropertyValue(context, property);
    }
    void setPropertyValue(Context& context, const string& property, const string& value) const {}
    void contextCreated(ContextImpl& context, const map<string, string>& properties) const {
        PlatformData* d = new PlatformData();
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
