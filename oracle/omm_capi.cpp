// oracle/omm_capi.cpp -- TEST INFRASTRUCTURE, not product code.
//
// A small flat C API over the *unmodified* reference OpenMM C++ API (openmmapi/include/OpenMM.h), so that the
// Python test-suite and bench.py (ctypes) can build a System from numpy arrays and run it on any registered
// Platform ("Reference", "CPU", and our plugin "B200") through the reference's own public classes.
// (The reference's SWIG/C wrappers cannot be built here: no swig / doxygen.)  Compiled by oracle/Makefile
// against the headers where they lie in /root/reference; links oracle/_ref/libOpenMM.so.
#include "OpenMM.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/serialization/XmlSerializer.h"
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <sstream>
#include <fstream>

using namespace OpenMM;

static thread_local std::string g_err;
#define OMM_TRY try {
#define OMM_CATCH(ret) } catch (std::exception& e) { g_err = e.what(); return ret; }

extern "C" {

const char* omm_last_error() { return g_err.c_str(); }

int omm_load_plugin(const char* path) {
    OMM_TRY
    Platform::loadPluginLibrary(path);
    return 0;
    OMM_CATCH(-1)
}

int omm_num_platforms() { return Platform::getNumPlatforms(); }
const char* omm_platform_name(int i) {
    static thread_local std::string s;
    s = Platform::getPlatform(i).getName();
    return s.c_str();
}

// ---------------------------------------------------------------- System
void* omm_system_create(int n, const double* masses) {
    System* s = new System();
    for (int i = 0; i < n; i++) s->addParticle(masses[i]);
    return s;
}
void omm_system_destroy(void* s) { delete (System*) s; }
void omm_system_set_box(void* s, const double* a, const double* b, const double* c) {
    ((System*) s)->setDefaultPeriodicBoxVectors(Vec3(a[0], a[1], a[2]), Vec3(b[0], b[1], b[2]), Vec3(c[0], c[1], c[2]));
}
void omm_system_add_constraints(void* s, int n, const int* i, const int* j, const double* d) {
    for (int k = 0; k < n; k++) ((System*) s)->addConstraint(i[k], j[k], d[k]);
}
int omm_system_add_force(void* s, void* f) { return ((System*) s)->addForce((Force*) f); }
int omm_system_serialize(void* s, const char* path) {
    OMM_TRY
    std::ofstream out(path);
    XmlSerializer::serialize<System>((System*) s, "System", out);
    return 0;
    OMM_CATCH(-1)
}

// ---------------------------------------------------------------- NonbondedForce
// method: 0 NoCutoff, 1 CutoffNonPeriodic, 2 CutoffPeriodic, 3 Ewald, 4 PME, 5 LJPME (NonbondedForce.h:114-143)
void* omm_nonbonded_create(int n, const double* q, const double* sigma, const double* eps) {
    NonbondedForce* f = new NonbondedForce();
    for (int i = 0; i < n; i++) f->addParticle(q[i], sigma[i], eps[i]);
    return f;
}
void omm_nonbonded_add_exceptions(void* f, int n, const int* i, const int* j, const double* qq, const double* sigma, const double* eps) {
    for (int k = 0; k < n; k++) ((NonbondedForce*) f)->addException(i[k], j[k], qq[k], sigma[k], eps[k]);
}
// NonbondedForce::createExceptionsFromBonds (NonbondedForce.cpp:207-248) on the real reference object, then read back
void omm_nonbonded_create_exceptions_from_bonds(void* f, int n, const int* i, const int* j, double coulomb14, double lj14) {
    std::vector<std::pair<int, int> > bonds(n);
    for (int k = 0; k < n; k++) bonds[k] = std::make_pair(i[k], j[k]);
    ((NonbondedForce*) f)->createExceptionsFromBonds(bonds, coulomb14, lj14);
}
int omm_nonbonded_num_exceptions(void* f) { return ((NonbondedForce*) f)->getNumExceptions(); }
void omm_nonbonded_get_exceptions(void* f, int* i, int* j, double* qq, double* sigma, double* eps) {
    NonbondedForce* nb = (NonbondedForce*) f;
    for (int k = 0; k < nb->getNumExceptions(); k++) nb->getExceptionParameters(k, i[k], j[k], qq[k], sigma[k], eps[k]);
}
// global parameters and parameter offsets (NonbondedForce.h addGlobalParameter / addParticleParameterOffset / addExceptionParameterOffset)
int omm_nonbonded_add_global(void* f, const char* name, double defaultValue) { return ((NonbondedForce*) f)->addGlobalParameter(name, defaultValue); }
int omm_nonbonded_add_particle_offset(void* f, const char* name, int particle, double dq, double dsigma, double deps) {
    return ((NonbondedForce*) f)->addParticleParameterOffset(name, particle, dq, dsigma, deps);
}
int omm_nonbonded_add_exception_offset(void* f, const char* name, int exception, double dqq, double dsigma, double deps) {
    return ((NonbondedForce*) f)->addExceptionParameterOffset(name, exception, dqq, dsigma, deps);
}
void omm_force_destroy(void* f) { delete (Force*) f; }
void omm_nonbonded_set_method(void* f, int method, double cutoff, double ewaldTol) {
    NonbondedForce* nb = (NonbondedForce*) f;
    nb->setNonbondedMethod((NonbondedForce::NonbondedMethod) method);
    nb->setCutoffDistance(cutoff);
    nb->setEwaldErrorTolerance(ewaldTol);
}
void omm_nonbonded_set_pme(void* f, double alpha, int nx, int ny, int nz) { ((NonbondedForce*) f)->setPMEParameters(alpha, nx, ny, nz); }
void omm_nonbonded_set_switch(void* f, int use, double dist) {
    ((NonbondedForce*) f)->setUseSwitchingFunction(use != 0);
    ((NonbondedForce*) f)->setSwitchingDistance(dist);
}
void omm_nonbonded_set_dispersion(void* f, int use) { ((NonbondedForce*) f)->setUseDispersionCorrection(use != 0); }
void omm_nonbonded_set_rf_dielectric(void* f, double d) { ((NonbondedForce*) f)->setReactionFieldDielectric(d); }
void omm_nonbonded_set_recip_group(void* f, int g) { ((NonbondedForce*) f)->setReciprocalSpaceForceGroup(g); }
void omm_nonbonded_set_exceptions_periodic(void* f, int p) { ((NonbondedForce*) f)->setExceptionsUsePeriodicBoundaryConditions(p != 0); }
void omm_force_set_group(void* f, int g) { ((Force*) f)->setForceGroup(g); }

// ---------------------------------------------------------------- bonded
void* omm_bonds_create(int n, const int* i, const int* j, const double* r0, const double* k) {
    HarmonicBondForce* f = new HarmonicBondForce();
    for (int b = 0; b < n; b++) f->addBond(i[b], j[b], r0[b], k[b]);
    return f;
}
void* omm_angles_create(int n, const int* i, const int* j, const int* k, const double* th0, const double* kk) {
    HarmonicAngleForce* f = new HarmonicAngleForce();
    for (int b = 0; b < n; b++) f->addAngle(i[b], j[b], k[b], th0[b], kk[b]);
    return f;
}
void* omm_torsions_create(int n, const int* i, const int* j, const int* k, const int* l, const int* per, const double* phase, const double* kk) {
    PeriodicTorsionForce* f = new PeriodicTorsionForce();
    for (int b = 0; b < n; b++) f->addTorsion(i[b], j[b], k[b], l[b], per[b], phase[b], kk[b]);
    return f;
}
// kind: 0 HarmonicBondForce, 1 HarmonicAngleForce, 2 PeriodicTorsionForce (each class has its own, non-virtual setter)
void omm_bonded_set_periodic(void* f, int kind, int p) {
    if (kind == 0) ((HarmonicBondForce*) f)->setUsesPeriodicBoundaryConditions(p != 0);
    else if (kind == 1) ((HarmonicAngleForce*) f)->setUsesPeriodicBoundaryConditions(p != 0);
    else ((PeriodicTorsionForce*) f)->setUsesPeriodicBoundaryConditions(p != 0);
}
void* omm_cmmotion_create(int freq) { return new CMMotionRemover(freq); }

// ---------------------------------------------------------------- Integrators
// kind: 0 Verlet, 1 Langevin, 2 LangevinMiddle
void* omm_integrator_create(int kind, double temperature, double friction, double dt, int seed, double constraintTol) {
    Integrator* integ;
    if (kind == 0) integ = new VerletIntegrator(dt);
    else if (kind == 1) { LangevinIntegrator* l = new LangevinIntegrator(temperature, friction, dt); l->setRandomNumberSeed(seed); integ = l; }
    else { LangevinMiddleIntegrator* l = new LangevinMiddleIntegrator(temperature, friction, dt); l->setRandomNumberSeed(seed); integ = l; }
    integ->setConstraintTolerance(constraintTol);
    return integ;
}
void omm_integrator_destroy(void* i) { delete (Integrator*) i; }
int omm_integrator_step(void* i, int n) {
    OMM_TRY
    ((Integrator*) i)->step(n);
    return 0;
    OMM_CATCH(-1)
}

// ---------------------------------------------------------------- Context
// props: "key=value;key=value"
void* omm_context_create(void* sys, void* integ, const char* platformName, const char* props) {
    OMM_TRY
    Platform& p = Platform::getPlatformByName(platformName);
    std::map<std::string, std::string> pm;
    if (props != NULL) {
        std::stringstream ss(props);
        std::string item;
        while (std::getline(ss, item, ';')) {
            size_t eq = item.find('=');
            if (eq != std::string::npos) pm[item.substr(0, eq)] = item.substr(eq+1);
        }
    }
    return new Context(*(System*) sys, *(Integrator*) integ, p, pm);
    OMM_CATCH(NULL)
}
void omm_context_destroy(void* c) { delete (Context*) c; }
int omm_context_set_positions(void* c, int n, const double* x) {
    OMM_TRY
    std::vector<Vec3> v(n);
    for (int i = 0; i < n; i++) v[i] = Vec3(x[3*i], x[3*i+1], x[3*i+2]);
    ((Context*) c)->setPositions(v);
    return 0;
    OMM_CATCH(-1)
}
int omm_context_set_velocities(void* c, int n, const double* x) {
    OMM_TRY
    std::vector<Vec3> v(n);
    for (int i = 0; i < n; i++) v[i] = Vec3(x[3*i], x[3*i+1], x[3*i+2]);
    ((Context*) c)->setVelocities(v);
    return 0;
    OMM_CATCH(-1)
}
int omm_context_set_box(void* c, const double* a, const double* b, const double* cc) {
    OMM_TRY
    ((Context*) c)->setPeriodicBoxVectors(Vec3(a[0], a[1], a[2]), Vec3(b[0], b[1], b[2]), Vec3(cc[0], cc[1], cc[2]));
    return 0;
    OMM_CATCH(-1)
}
int omm_context_set_velocities_to_temperature(void* c, double T, int seed) {
    OMM_TRY
    ((Context*) c)->setVelocitiesToTemperature(T, seed);
    return 0;
    OMM_CATCH(-1)
}
int omm_context_apply_constraints(void* c, double tol) {
    OMM_TRY
    ((Context*) c)->applyConstraints(tol);
    return 0;
    OMM_CATCH(-1)
}
// Any of pos/vel/frc may be NULL. energies[0] = potential, energies[1] = kinetic. groups = force-group mask.
int omm_context_get_state(void* c, int n, double* pos, double* vel, double* frc, double* energies, int groups, int enforcePeriodic) {
    OMM_TRY
    int types = 0;
    if (pos) types |= State::Positions;
    if (vel) types |= State::Velocities;
    if (frc) types |= State::Forces;
    if (energies) types |= State::Energy;
    State s = ((Context*) c)->getState(types, enforcePeriodic != 0, groups);
    if (pos) { const std::vector<Vec3>& v = s.getPositions(); for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) pos[3*i+k] = v[i][k]; }
    if (vel) { const std::vector<Vec3>& v = s.getVelocities(); for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) vel[3*i+k] = v[i][k]; }
    if (frc) { const std::vector<Vec3>& v = s.getForces(); for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) frc[3*i+k] = v[i][k]; }
    if (energies) { energies[0] = s.getPotentialEnergy(); energies[1] = s.getKineticEnergy(); }
    return 0;
    OMM_CATCH(-1)
}
int omm_context_get_pme(void* c, void* nbforce, double* alpha, int* nx, int* ny, int* nz) {
    OMM_TRY
    ((NonbondedForce*) nbforce)->getPMEParametersInContext(*(Context*) c, *alpha, *nx, *ny, *nz);
    return 0;
    OMM_CATCH(-1)
}
int omm_context_set_parameter(void* c, const char* name, double value) {
    OMM_TRY
    ((Context*) c)->setParameter(name, value);
    return 0;
    OMM_CATCH(-1)
}
double omm_context_get_time(void* c) { return ((Context*) c)->getState(0).getTime(); }
const char* omm_context_platform(void* c) {
    static thread_local std::string s;
    s = ((Context*) c)->getPlatform().getName();
    return s.c_str();
}
int omm_context_checkpoint_roundtrip(void* c) {
    OMM_TRY
    std::stringstream ss(std::ios_base::in | std::ios_base::out | std::ios_base::binary);
    ((Context*) c)->createCheckpoint(ss);
    ((Context*) c)->loadCheckpoint(ss);
    return 0;
    OMM_CATCH(-1)
}

} // extern "C"
