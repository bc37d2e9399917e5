// B200PmePlugin.cpp -- libOpenMMB200Pme.so: the bespoke reciprocal-space PME as a drop-in "CalcPmeReciprocalForce" kernel
// for the reference's OTHER platforms.
//
// The reference lets a platform outsource reciprocal space to a kernel of that name supplied by a plugin: the CPU platform
// uses it whenever one is registered (CpuKernels.cpp:620-690), the CUDA platform under UseCpuPme (CudaKernels.cpp:746-760);
// plugins/cpupme (CpuPmeKernelFactory.cpp:36-45) is the stock provider.  This library is the same shape: loading it
// (Platform::loadPluginLibrary) registers a factory on every platform except B200 itself, whose kernel forwards
// CalcPmeReciprocalForceKernel (olla/include/openmm/kernels.h:1493-1557) to b200md_pme_create / b200md_pme_exec.
// It is a SEPARATE library from libOpenMMB200.so on purpose: loading the Platform must never change what the CPU platform
// computes (bench.py times the untouched CPU platform as the baseline).
#include "openmm/Platform.h"
#include "openmm/KernelFactory.h"
#include "openmm/kernels.h"
#include "openmm/OpenMMException.h"
#include "../include/b200md.h"
#include <atomic>
#include <cstdlib>
#include <string>
#include <vector>

using namespace OpenMM;
using namespace std;

namespace {

atomic<long> execCount(0);

int fftFriendly(int n) {            // next size with prime factors <= 13 (cpupme rounds with findFFTDimension the same way)
    for (;; n++) {
        int r = n;
        for (int p : {2, 3, 5, 7, 11, 13}) while (r % p == 0) r /= p;
        if (r == 1) return n;
    }
}

class B200CalcPmeReciprocalForceKernel : public CalcPmeReciprocalForceKernel {
public:
    B200CalcPmeReciprocalForceKernel(string name, const Platform& platform) : CalcPmeReciprocalForceKernel(name, platform), ctx(nullptr), n(0), alpha(0), energy(0) {
        grid[0] = grid[1] = grid[2] = 0;
    }
    ~B200CalcPmeReciprocalForceKernel() { if (ctx) b200md_destroy(ctx); }
    void initialize(int gridx, int gridy, int gridz, int numParticles, double alpha, bool deterministic) {
        (void) deterministic;        // charge spreading is fixed point: always deterministic
        grid[0] = fftFriendly(gridx); grid[1] = fftFriendly(gridy); grid[2] = fftFriendly(gridz);
        n = numParticles; this->alpha = alpha;
        const char* dev = getenv("OPENMM_B200_DEVICE");
        if (b200md_pme_create(&ctx, dev ? atoi(dev) : 0, n, grid[0], grid[1], grid[2], alpha) != 0) {
            string msg = string("B200 PME: ") + b200md_last_error(ctx);
            if (ctx) { b200md_destroy(ctx); ctx = nullptr; }
            throw OpenMMException(msg);
        }
        force.assign((size_t) 4*n, 0.f);
    }
    void beginComputation(IO& io, const Vec3* periodicBoxVectors, bool includeEnergy) {
        double box[9];
        for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) box[3*i+k] = periodicBoxVectors[i][k];
        if (b200md_pme_exec(ctx, io.getPosq(), box, includeEnergy ? 1 : 0, force.data(), &energy) != 0)
            throw OpenMMException(string("B200 PME: ") + b200md_last_error(ctx));
        execCount++;
    }
    double finishComputation(IO& io) {
        io.setForce(force.data());
        return energy;
    }
    void getPMEParameters(double& alpha, int& nx, int& ny, int& nz) const { alpha = this->alpha; nx = grid[0]; ny = grid[1]; nz = grid[2]; }
private:
    b200md_ctx* ctx;
    int n, grid[3];
    double alpha, energy;
    vector<float> force;
};

class B200PmeKernelFactory : public KernelFactory {
public:
    KernelImpl* createKernelImpl(string name, const Platform& platform, ContextImpl& context) const {
        (void) context;
        if (name == CalcPmeReciprocalForceKernel::Name()) return new B200CalcPmeReciprocalForceKernel(name, platform);
        throw OpenMMException((string("Tried to create kernel with illegal kernel name '") + name + "'").c_str());
    }
};

} // namespace

extern "C" __attribute__((visibility("default"))) void registerPlatforms() {
}

extern "C" __attribute__((visibility("default"))) void registerKernelFactories() {
    B200PmeKernelFactory* factory = new B200PmeKernelFactory();
    for (int i = 0; i < Platform::getNumPlatforms(); i++) {
        Platform& p = Platform::getPlatform(i);
        if (p.getName() != "B200") p.registerKernelFactory(CalcPmeReciprocalForceKernel::Name(), factory);
    }
}

// number of reciprocal-space evaluations served so far (tests use it to prove the hook was taken)
extern "C" __attribute__((visibility("default"))) long b200pme_exec_count() { return execCount.load(); }
