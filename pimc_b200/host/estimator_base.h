// estimator_base.h -- stand-alone mirror of the reference's estimator plugin interface
// (class EstimatorBase, include/estimator.h:32-133; lifecycle src/estimator.cpp:163-429; wave-vector generation
// :439-570; registration macro :31-34).  Same public/protected names so that estimator_b200.cpp compiles unchanged
// against either this header or the reference's estimator.h.  Written from scratch; no Boost (boost::format is
// replaced by snprintf with the same conversion specifications).
#ifndef PIMCB_ESTIMATOR_BASE_H
#define PIMCB_ESTIMATOR_BASE_H

#include "pimc_compat.h"

class EstimatorBase {
public:
    EstimatorBase(const Path& _path, ActionBase* _actionPtr, const MTRand& _random, double _maxR, int _frequency = 1,
                  std::string _label = "");
    virtual ~EstimatorBase();

    virtual void sample();
    void reset();
    void restart(const uint32, const uint32);
    virtual void output();
    bool baseSample();
    uint32 getTotNumAccumulated() const { return totNumAccumulated; }
    uint32 getNumAccumulated() const { return numAccumulated; }
    uint32 getNumSampled() const { return numSampled; }
    virtual std::string getName() const { return "base"; }
    void prepare();
    void addEndLine() { endLine = true; }
    void appendLabel(std::string append) { label = label + append; }
    std::string dVecToString(const dVec&);
    std::string getLabel() const { return label; }
    const std::string& getHeader() const { return header; }
    const DynamicArray<double, 1>& values() const { return estimator; }

protected:
    const Path& path;
    ActionBase* actionPtr;
    MTRand random;
    double maxR;
    std::fstream* outFilePtr = nullptr;
    DynamicArray<double, 1> estimator;
    DynamicArray<double, 1> norm;
    int numEst = 0;
    int frequency;
    std::string label;
    uint32 numSampled, numAccumulated, totNumAccumulated;
    int numBeads0;
    bool diagonal, endLine, canonical;
    std::string header;

    std::map<std::string, int> estIndex;            // include/estimator.h:104
    std::vector<double> sliceFactor;                // 1.0 on every slice in PIMC mode (src/estimator.cpp:182-199)
    int startSlice = 0, endSlice = 0, endDiagSlice = 0;

    virtual void accumulate() {}
    void initialize(int);
    void initialize(std::vector<std::string> estLabel);
    void getQVectors(std::vector<dVec>&);
    std::vector<std::vector<dVec>> getQVectors2(double, double, int&, std::string);   // include/estimator.h:132
};

// src/estimator.cpp:31-34
#define REGISTER_ESTIMATOR(NAME, TYPE) \
    const std::string TYPE::name = NAME; \
    bool reg##TYPE = estimatorFactory()->Register<TYPE>(TYPE::name);

std::string pimcb_format(const char* spec, double v);   // boost::format(spec) % v for one floating conversion
std::string pimcb_format(const char* spec, int v);

#endif
