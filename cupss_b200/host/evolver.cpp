// evolver.cpp -- system declaration, step driver and output on top of the B200 engine.
// Behavioural reference: /root/reference/src/evolver.cpp (ctor :13-74, prepareProblem :81-126,
// createField :177-196, advanceTime :199-226, createTerm :326-362, updateParameter :386-394).
#include <sys/stat.h>

#include <cmath>
#include <cstdlib>
#include <ctime>
#include <cstring>
#include <iomanip>
#include <vector>

#include <cuda_runtime.h>

#include "../../inc/cupss.h"
#include "../../include/cupss_b200.h"

void evolver::common_constructor() {
    currentTime = 0.0f;
    currentTimeStep = 0;
    dtsqrt = std::sqrt(dt);
    writePrecision = 6;
    writeParametersOnUpdate = true;
    verbose = false;
    _parser = new parser(this);
    // legacy launch geometry: public, read by user kernels (examples/07_inhomogeneous_diffusion)
    if (sy == 1 && sz == 1) {
        dimension = 1;
        blocks = dim3(1);
        threads_per_block = dim3(sx);
    } else if (sz == 1) {
        dimension = 2;
        threads_per_block = dim3(32, 32);
        blocks = dim3((sx + 31) / 32, (sy + 31) / 32);
    } else {
        dimension = 3;
        threads_per_block = dim3(16, 8, 8);
        blocks = dim3((sx + 15) / 16, (sy + 7) / 8, (sz + 7) / 8);
    }
}

evolver::evolver(bool cuda, int nx, float hx, float step, int every)
    : dt(step), sx(nx), sy(1), sz(1), dx(hx), dy(1.0f), dz(1.0f), writeEveryNSteps(every), with_cuda(cuda) {
    std::srand(time(NULL));
    common_constructor();
}
evolver::evolver(bool cuda, int nx, int ny, float hx, float hy, float step, int every)
    : dt(step), sx(nx), sy(ny), sz(1), dx(hx), dy(hy), dz(1.0f), writeEveryNSteps(every), with_cuda(cuda) {
    std::srand(time(NULL));
    common_constructor();
}
evolver::evolver(bool cuda, int nx, int ny, int nz, float hx, float hy, float hz, float step, int every)
    : dt(step), sx(nx), sy(ny), sz(nz), dx(hx), dy(hy), dz(hz), writeEveryNSteps(every), with_cuda(cuda) {
    std::srand(time(NULL));
    common_constructor();
}

evolver::~evolver() {
    if (plan) cupss_b200_destroy(plan);
    for (field *f : fields) delete f;
    delete _parser;
}

void evolver::engineCheck(int code, const char *what) {
    if (code == CUPSS_B200_OK) return;
    // fatal, like check_error() in the reference (src/cu_utils.cpp:4-10)
    std::cerr << "cuPSS B200 engine error in " << what << ": " << cupss_b200_last_error() << std::endl;
    std::exit(1);
}

field *evolver::findField(const std::string &name, const char *who) {
    auto it = fieldsMap.find(name);
    if (it == fieldsMap.end()) {
        std::cout << "ERROR in " << who << ", " << name << " not found" << std::endl;
        std::exit(1);
    }
    return it->second;
}

void evolver::setPartition(int rank, int nranks, const void *id) {
    if (plan) {
        std::cout << "ERROR: setPartition must be called before prepareProblem" << std::endl;
        std::exit(1);
    }
    partRank = rank;
    partRanks = nranks;
    if (id) memcpy(partId, id, 128);
}

int evolver::createFromFile(const std::string &path) {
    _parser->createFromFile(path);
    return 0;
}

int evolver::existsField(const std::string &name) {
    for (size_t i = 0; i < fields.size(); i++)
        if (fields[i]->name == name) return (int)i;
    return -1;
}

void evolver::addField(field *f) { fields.push_back(f); }

int evolver::createField(std::string name, bool dynamic) {
    if (existsField(name) >= 0) {
        std::cout << "Trying to create field with name that already exists" << std::endl;
        return 1;
    }
    field *f = new field(sx, sy, sz, dx, dy, dz);
    f->name = name;
    f->dynamic = dynamic;
    f->isCUDA = with_cuda;
    f->blocks = blocks;
    f->threads_per_block = threads_per_block;
    f->system_p = this;
    fields.push_back(f);
    fieldsMap[name] = f;
    fieldsReal[name] = f->real_array;
    fieldsFourier[name] = f->comp_array;
    planDirty = true;
    return 0;
}

int evolver::createTerm(const std::string &field_name, const std::vector<pres> &prefactors, const std::vector<std::string> &product) {
    const int idx = existsField(field_name);
    if (idx < 0) {
        std::cout << "Field " << field_name << " not found trying to create term" << std::endl;
        return 1;
    }
    term *t = new term(sx, sy, sz, dx, dy, dz);
    t->isCUDA = with_cuda;
    t->blocks = blocks;
    t->threads_per_block = threads_per_block;
    for (const std::string &p : product) {
        const int pi = existsField(p);
        if (pi >= 0) t->product.push_back(fields[pi]);   // unknown names are silently skipped, as in the reference
    }
    t->prefactors_h = prefactors;
    fields[idx]->terms.push_back(t);
    planDirty = true;
    return 0;
}

int evolver::addParameter(const std::string &name, float value) {
    _parser->insert_parameter(name, value);
    return 0;
}

int evolver::addEquation(const std::string &equation) {
    _parser->add_equation(equation);
    planDirty = true;
    return 0;
}

int evolver::addNoise(const std::string &field_name, const std::string &amplitude) {
    if (existsField(field_name) < 0) {
        std::cout << "Adding noise to non existing field! (" << field_name << ")" << std::endl;
        return -1;
    }
    field *f = fieldsMap[field_name];
    f->noise_amplitude = _parser->add_noise(amplitude);
    f->isNoisy = true;
    planDirty = true;
    return 0;
}

void evolver::setOutputField(const std::string &name, int on) {
    const int idx = existsField(name);
    if (idx < 0) {
        std::cout << "setOutputField EROR: " << name << " not found." << std::endl;
        return;
    }
    fields[idx]->outputToFile = on != 0;
}

// Everything the parser produced goes to the engine in one go; finalize builds the fused schedule.
void evolver::sendSystemToEngine() {
    if (!seedFixed) {
        // The reference seeds from time(NULL).  The Philox stream is keyed on (global mode, field, step, seed) and must be the
        // same on every rank of a partitioned run (conjugate partner rows of the kx = 0 and kx = sx/2 planes live on different
        // ranks): there the default seed is a hash of the 128-byte NCCL unique id, which all ranks share and which differs
        // from run to run; ranks starting in different seconds would otherwise disagree.
        if (partRanks > 1) {
            unsigned long long h = 1469598103934665603ull;   // FNV-1a
            for (int i = 0; i < 128; ++i) { h ^= (unsigned char)partId[i]; h *= 1099511628211ull; }
            noiseSeed = h;
        } else {
            noiseSeed = (unsigned long long)time(NULL);
        }
        seedFixed = true;
    }
    for (field *f : fields) {
        std::vector<cupss_b200_pres> imp;
        for (const pres &p : f->implicit) imp.push_back({p.preFactor, p.q2n, p.iqx, p.iqy, p.iqz, p.invq});
        engineCheck(cupss_b200_set_implicit(plan, f->engine_id, imp.data(), (int)imp.size()), "set_implicit");
        engineCheck(cupss_b200_clear_terms(plan, f->engine_id), "clear_terms");
        for (term *t : f->terms) {
            std::vector<cupss_b200_pres> pv;
            for (const pres &p : t->prefactors_h) pv.push_back({p.preFactor, p.q2n, p.iqx, p.iqy, p.iqz, p.invq});
            std::vector<int> prod;
            for (field *g : t->product) prod.push_back(g->engine_id);
            engineCheck(cupss_b200_add_term(plan, f->engine_id, pv.data(), (int)pv.size(), prod.data(), (int)prod.size()), "add_term");
        }
        cupss_b200_pres amp = {f->noise_amplitude.preFactor, f->noise_amplitude.q2n, 0, 0, 0, f->noise_amplitude.invq};
        engineCheck(cupss_b200_set_noise(plan, f->engine_id, f->isNoisy ? &amp : nullptr, noiseSeed), "set_noise");
    }
    engineCheck(cupss_b200_set_dealias_rule(plan, with_cuda ? CUPSS_B200_DEALIAS_GPU_RULE : CUPSS_B200_DEALIAS_CPU_RULE), "set_dealias_rule");
    engineCheck(cupss_b200_finalize(plan), "finalize");
    for (field *f : fields) {
        int needs = 0, order = 1;
        engineCheck(cupss_b200_field_alias(plan, f->engine_id, &needs, &order), "field_alias");
        f->needsaliasing = needs != 0;
        f->aliasing_order = order;
        f->comp_array_d = static_cast<float2 *>(cupss_b200_device_spectrum(plan, f->engine_id));
    }
    planDirty = false;
}

void evolver::prepareProblem() {
    if (with_cuda && verbose) check_device();
    struct stat info;
    if (stat("data", &info) != 0) {
        if (verbose) std::cout << "data directory not found, creating it.\n";
        if (mkdir("data", S_IRWXU | S_IRWXG | S_IROTH | S_IXOTH) == -1) {
            std::cout << "Error creating data directory\n";
            std::exit(1);
        }
    } else if (!(info.st_mode & S_IFDIR)) {
        std::cout << "Can't create data directory, is there a file called data?\n";
        std::exit(1);
    } else if (verbose) {
        std::cout << "data directory already found, might rewrite output data.\n";
    }
    _parser->writeParamsToFile("data/parameter_list.txt.0");

    if (verbose) std::cout << "Preparing problem." << std::endl;
    if (!plan) {
        engineCheck(cupss_b200_create(&plan, sx, sy, sz, dx, dy, dz, dt), "create");
        if (partRanks > 1) engineCheck(cupss_b200_set_partition(plan, partRank, partRanks, partId), "set_partition");
        for (field *f : fields) {
            const int id = cupss_b200_add_field(plan, f->name.c_str(), f->dynamic ? 1 : 0);
            if (id < 0) engineCheck(-id, "add_field");
            f->engine_id = id;
        }
    } else {
        for (field *f : fields) {
            if (f->engine_id >= 0) continue;   // fields declared after an earlier prepareProblem
            const int id = cupss_b200_add_field(plan, f->name.c_str(), f->dynamic ? 1 : 0);
            if (id < 0) engineCheck(-id, "add_field");
            f->engine_id = id;
        }
    }
    // initial conditions: host real arrays -> spectra (copyHostToDevice + toComp for every field)
    if (verbose) std::cout << "Copying initial states to device." << std::endl;
    for (field *f : fields) {
        for (term *t : f->terms) t->prepareDevice();
        const size_t slab = (size_t)sx * sy * (sz / partRanks) * partRank;   // this rank's z-slab of the full host array
        engineCheck(cupss_b200_upload_real(plan, f->engine_id, reinterpret_cast<const float *>(f->real_array + slab)), "upload_real");
        f->mirror_in_sync = true;
    }
    if (verbose) std::cout << "Building the fused per-equation plan." << std::endl;
    sendSystemToEngine();
    for (field *f : fields) {
        if (f->hasCBFourier && partRanks > 1) {
            std::cout << "ERROR: Fourier-space callbacks (field " << f->name << ") need a single-GPU run" << std::endl;
            std::exit(1);
        }
        f->system_p = this;
    }
}

// field::setRHS' callback hook (src/field.cpp:68-86): the user function sees the real field -- and, when products read the
// field, its dealiased copy -- as float2[N] with the value in .x: a device pointer on the RUN_GPU path, the host array on
// the RUN_CPU path.  Only flagged fields pay for the materialised view.
void evolver::applyCallback(field *f) {
    if (f->callback == NULL) {
        std::cout << "Wants to apply callback function but pointer to function is NULL" << std::endl;
        return;
    }
    // partitioned run: the callback sees this rank's z-slab, float2[sz/P][sy][sx], and is told so through its sz argument
    const int zl = sz / partRanks;
    const size_t n = (size_t)sx * sy * zl;
    const size_t slab = n * (size_t)partRank;
    for (int which = 0; which < (f->needsaliasing ? 2 : 1); ++which) {
        void *dev = nullptr;
        engineCheck(cupss_b200_real_view_begin(plan, f->engine_id, which, &dev), "real_view_begin");
        if (with_cuda) {
            f->callback(this, static_cast<float2 *>(dev), sx, sy, zl);
        } else {
            std::vector<float2> tmp;
            float2 *host = f->real_array + slab;
            if (which == 1) { tmp.resize(n); host = tmp.data(); }
            if (cudaMemcpy(host, dev, n * sizeof(float2), cudaMemcpyDeviceToHost) != cudaSuccess) engineCheck(2, "callback download");
            f->callback(this, host, sx, sy, zl);
            if (cudaMemcpy(dev, host, n * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess) engineCheck(2, "callback upload");
        }
        engineCheck(cupss_b200_real_view_commit(plan, f->engine_id, which), "real_view_commit");
    }
}

// field::setRHS' Fourier hook (src/field.cpp:48-57): the user function sees comp_array as the reference lays it out, the
// full float2[sz][sy][sx] spectrum, right after the update and before dealias / toReal.
void evolver::applyCallbackFourier(field *f) {
    if (f->callbackFourier == NULL) {
        std::cout << "Wants to apply callback function in Fourier space but pointer to function is NULL" << std::endl;
        return;
    }
    const size_t n = (size_t)sx * sy * sz;
    void *dev = nullptr;
    engineCheck(cupss_b200_comp_view_begin(plan, f->engine_id, &dev), "comp_view_begin");
    if (with_cuda) {
        f->callbackFourier(this, static_cast<float2 *>(dev), sx, sy, sz);
    } else {
        if (cudaMemcpy(f->comp_array, dev, n * sizeof(float2), cudaMemcpyDeviceToHost) != cudaSuccess) engineCheck(2, "callback download");
        f->callbackFourier(this, f->comp_array, sx, sy, sz);
        if (cudaMemcpy(dev, f->comp_array, n * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess) engineCheck(2, "callback upload");
    }
    engineCheck(cupss_b200_comp_view_commit(plan, f->engine_id), "comp_view_commit");
}

int evolver::advanceTime() {
    if (!plan) {
        std::cout << "ERROR: advanceTime called before prepareProblem" << std::endl;
        std::exit(1);
    }
    if (currentTimeStep % writeEveryNSteps == 0) writeOut();
    if (planDirty) sendSystemToEngine();
    bool anyCB = false;
    for (field *f : fields) anyCB = anyCB || f->hasCB || f->hasCBFourier;
    for (field *f : fields) f->mirror_in_sync = false;
    if (!anyCB) {
        engineCheck(cupss_b200_step(plan, 1), "step");
    } else {
        // sweeps run eagerly with the callbacks in between, in the reference's order (constraint fields, then dynamic ones)
        engineCheck(cupss_b200_step_stage(plan, 0), "step_stage");
        for (field *f : fields) if (!f->dynamic) { if (f->hasCBFourier) applyCallbackFourier(f); if (f->hasCB) applyCallback(f); }
        engineCheck(cupss_b200_step_stage(plan, 1), "step_stage");
        for (field *f : fields) if (f->dynamic) { if (f->hasCBFourier) applyCallbackFourier(f); if (f->hasCB) applyCallback(f); }
    }
    if (!with_cuda) {
        // reference-CPU semantics: host arrays are live after every step (examples/06_kpz reads them directly)
        for (field *f : fields) refreshHostMirror(f, true, true);
    }
    currentTime += dt;
    currentTimeStep += 1;
    return 0;
}

void evolver::refreshHostMirror(field *f, bool real_part, bool comp_part, bool keep_exact) {
    if (!plan || f->engine_id < 0) return;
    const size_t slab = (size_t)sx * sy * (sz / partRanks) * partRank;
    // Until the first step after an upload the host real array is the device state bit for bit (the reference's real_array_d is
    // a plain copy of it, src/field_init.cpp; writeOut at step 0 prints the initial condition exactly); the engine only keeps
    // the spectrum, and spectrum -> real would return the same values with 1e-7 of round-off on top.
    // (Only the output writer asks for this: an explicit copyDeviceToHost always returns the device state.)
    if (real_part && !(keep_exact && f->mirror_in_sync)) engineCheck(cupss_b200_download_real(plan, f->engine_id, reinterpret_cast<float *>(f->real_array + slab)), "download_real");
    // partitioned: collective (every rank calls it); each rank receives the kz planes of its own z-slab of the full spectrum
    if (comp_part) engineCheck(cupss_b200_download_comp(plan, f->engine_id, reinterpret_cast<float *>(f->comp_array + slab)), "download_comp");
}

// field::copyHostToDevice / copyRealHostToDevice (src/field.cpp:337-348): the user edited the host real array mid-run and
// pushes it to the device.  The engine keeps spectra, so the upload is the forward transform of the (slab of the) host array.
void evolver::uploadHostMirror(field *f) {
    if (!plan || f->engine_id < 0) return;   // before prepareProblem the host arrays ARE the state
    const size_t slab = (size_t)sx * sy * (sz / partRanks) * partRank;
    engineCheck(cupss_b200_upload_real(plan, f->engine_id, reinterpret_cast<const float *>(f->real_array + slab)), "upload_real");
    f->mirror_in_sync = true;
}

void evolver::writeOut() {
    for (field *f : fields) f->writeToFile(currentTimeStep, dimension, writePrecision);
}

void evolver::copyAllDataToHost() {
    for (field *f : fields) refreshHostMirror(f, true, true);
}

int evolver::updateParameter(const std::string &name, float value) {
    _parser->changeParameter(name, value);
    for (field *f : fields) f->updateParameter(name, value);
    planDirty = true;   // constants are re-baked into the plan before the next step
    if (writeParametersOnUpdate) _parser->writeParamsToFile("data/parameter_list.txt." + std::to_string(currentTimeStep));
    return 0;
}

void evolver::setVerbose() { verbose = true; }
void evolver::unsetVerbose() { verbose = false; }
int evolver::getSystemSizeX() { return sx; }
int evolver::getSystemSizeY() { return sy; }
int evolver::getSystemSizeZ() { return sz; }
float evolver::getSystemPhysicalSizeX() { return (float)sx * dx; }
float evolver::getSystemPhysicalSizeY() { return (float)sy * dy; }
float evolver::getSystemPhysicalSizeZ() { return (float)sz * dz; }
int evolver::getCurrentTimestep() { return currentTimeStep; }
float evolver::getCurrentTime() { return currentTime; }
bool evolver::getCuda() { return with_cuda; }
float evolver::getParameter(const std::string &name) { return _parser->getParameter(name); }

static std::string describe(const pres &p, bool with_iq) {
    std::string s;
    if (with_iq) {
        if (p.iqx != 0) s += "(iqx)^(" + std::to_string(p.iqx) + ")";
        if (p.iqy != 0) s += "(iqy)^(" + std::to_string(p.iqy) + ")";
        if (p.iqz != 0) s += "(iqz)^(" + std::to_string(p.iqz) + ")";
    }
    if (p.q2n != 0) s += "(q)^(" + std::to_string(2 * p.q2n) + ")";
    if (p.invq != 0) s += "(1/|q|)^(" + std::to_string(p.invq) + ")";
    return s;
}

// Same text as the reference's printInformation (src/evolver.cpp:237-324); used as a parser known-answer test.
void evolver::printInformation() {
    std::cout << std::fixed << std::setprecision(3);
    std::cout << "Information on this evolver:" << std::endl;
    std::cout << dimension << "-dimensional system of size " << sx << "x" << sy << "x" << sz << std::endl;
    std::cout << "Physical size " << (float)sx * dx << "x" << (float)sy * dy << "x" << (float)sz * dz
              << " with cells of size " << dx << "x" << dy << "x" << dz << std::endl;
    std::cout << "There are " << fields.size() << " fields." << std::endl;
    for (size_t i = 0; i < fields.size(); i++) {
        field *f = fields[i];
        std::cout << "Field " << i << ": " << f->name << (f->dynamic ? " is dynamic." : " is not dynamic");
        std::cout << " and has " << f->terms.size() << " explicit terms and " << f->implicit.size() << " implicit terms.";
        std::cout << " Runs on GPU: " << f->isCUDA;
        if (f->needsaliasing) std::cout << ". Will be dealiased for a nonlinearity of order " << f->aliasing_order;
        else std::cout << ". Will not be dealised.";
        std::cout << std::endl << "\t";
        if (f->dynamic) std::cout << "(d/dt)";
        std::cout << f->name;
        if (f->dynamic) std::cout << " = ";
        if (!f->implicit.empty()) {
            std::string line = "[";
            for (const pres &p : f->implicit) {
                if (p.preFactor > 0.0f) line += "+";
                line += std::to_string(p.preFactor) + describe(p, true);
            }
            line += "]";
            if (f->dynamic) line += f->name;
            std::cout << line;
        }
        if (!f->dynamic) std::cout << " = ";
        for (size_t j = 0; j < f->terms.size(); j++) {
            term *t = f->terms[j];
            std::string line = j ? " + [" : " [";
            for (size_t p = 0; p < t->prefactors_h.size(); p++) {
                line += " + (" + std::to_string(t->prefactors_h[p].preFactor) + ")" + describe(t->prefactors_h[p], true);
                if (p + 1 != t->prefactors_h.size()) line += " + ";
            }
            line += "] (";
            for (field *g : t->product) line += " " + g->name;
            line += " )";
            std::cout << line;
        }
        if (f->isNoisy) {
            std::cout << "+ sqrt[2*" << f->noise_amplitude.preFactor;
            if (f->noise_amplitude.q2n != 0) std::cout << "*q^" << f->noise_amplitude.q2n * 2;
            if (f->noise_amplitude.invq != 0) std::cout << "*1/|q|^" << f->noise_amplitude.q2n;
            std::cout << "] x noise";
        }
        std::cout << std::endl << std::endl;
    }
}
