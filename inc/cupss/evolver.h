// evolver.h -- the user-facing solver object.
// Same constructors, methods and public data as /root/reference/inc/cupss/evolver.h:14-86; the step itself
// (advanceTime) runs on the B200 engine behind include/cupss_b200.h.
#ifndef CUPSS_B200_EVOLVER_H
#define CUPSS_B200_EVOLVER_H

#include <map>
#include <string>
#include <vector>
#include "defines.h"

class field;
class parser;
struct cupss_b200_plan;

class evolver {
   public:
    evolver(bool with_cuda, int sx, float dx, float dt, int writeEveryNSteps);
    evolver(bool with_cuda, int sx, int sy, float dx, float dy, float dt, int writeEveryNSteps);
    evolver(bool with_cuda, int sx, int sy, int sz, float dx, float dy, float dz, float dt, int writeEveryNSteps);
    ~evolver();
    void common_constructor();

    // ---- public data used by examples and tests
    int dimension;
    dim3 threads_per_block;   // legacy launch geometry, populated as the reference does
    dim3 blocks;
    float dt;
    float dtsqrt;
    std::vector<field *> fields;
    parser *_parser;
    std::map<std::string, field *> fieldsMap;
    std::map<std::string, float2 *> fieldsReal;
    std::map<std::string, float2 *> fieldsFourier;
    int writePrecision;
    bool writeParametersOnUpdate;

    // ---- system declaration
    void addField(field *f);
    int createField(std::string name, bool dynamic);
    int createTerm(const std::string &field_name, const std::vector<pres> &prefactors, const std::vector<std::string> &product);
    int addParameter(const std::string &name, float value);
    int addEquation(const std::string &equation);
    int addNoise(const std::string &field_name, const std::string &amplitude);
    int createFromFile(const std::string &path);
    int existsField(const std::string &name);

    // ---- dynamics
    void prepareProblem();
    int advanceTime();
    void writeOut();
    void copyAllDataToHost();
    void setOutputField(const std::string &name, int on);
    int updateParameter(const std::string &name, float value);

    // ---- initial conditions (host side, before prepareProblem)
    void initializeUniform(std::string field, float value);
    void initializeUniformNoise(std::string field, float amplitude);
    void initializeNormalNoise(std::string field, float mean, float sigma);
    void initializeHalfSystem(std::string field, float value1, float value2, float interface_width, int direction);
    void initializeDroplet(std::string field, float value_out, float value_in, float radius, float interface_width, int center_x, int center_y, int center_z);
    void addDroplet(std::string field, float value, float radius, float interface_width, int center_x, int center_y, int center_z);
    void initializeFromFile(std::string field, std::string file, int skiprows, char delimiter);

    // ---- getters / misc
    int getSystemSizeX();
    int getSystemSizeY();
    int getSystemSizeZ();
    float getSystemPhysicalSizeX();
    float getSystemPhysicalSizeY();
    float getSystemPhysicalSizeZ();
    int getCurrentTimestep();
    float getCurrentTime();
    bool getCuda();
    float getParameter(const std::string &name);
    void printInformation();
    void setVerbose();
    void unsetVerbose();

    // ---- B200 engine access (additions; not in the reference)
    cupss_b200_plan *enginePlan() { return plan; }
    void setNoiseSeed(unsigned long long seed) { noiseSeed = seed; seedFixed = true; }
    unsigned long long getNoiseSeed() const { return noiseSeed; }   // valid after prepareProblem (default: time, or a hash shared by all ranks)
    void refreshHostMirror(field *f, bool real_part, bool comp_part, bool keep_exact = false);
    void uploadHostMirror(field *f);   // host real array -> device state (field::copyHostToDevice)
    void markPlanDirty() { planDirty = true; }
    // Slab partition over `nranks` processes (one GPU each); 3-D only.  Host arrays stay full-size, each
    // rank reads/writes only its own z-slab [rank*sz/nranks, (rank+1)*sz/nranks).  Call before prepareProblem.
    void setPartition(int rank, int nranks, const void *nccl_unique_id_128);

   private:
    const int sx, sy, sz;
    const float dx, dy, dz;
    const int writeEveryNSteps;
    bool with_cuda;
    float currentTime;
    int currentTimeStep;
    bool verbose;

    cupss_b200_plan *plan = nullptr;
    bool planDirty = true;
    unsigned long long noiseSeed = 0;
    bool seedFixed = false;
    int partRank = 0, partRanks = 1;
    char partId[128] = {0};
    void applyCallbackFourier(field *f);   // field::setRHS' Fourier hook (src/field.cpp:48-57)
    void applyCallback(field *f);   // field::setRHS' callback hook (src/field.cpp:68-86)
    void sendSystemToEngine();   // implicit + terms + noise of every field -> C ABI, then finalize
    void engineCheck(int code, const char *what);
    field *findField(const std::string &name, const char *who);
};

#endif
