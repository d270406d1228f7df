// pimcb_measure -- stand-alone driver of the B200 measurement path through the reference's plugin API:
// builds Container / Path / constants, creates the estimators BY NAME through the estimator factory exactly as
// Setup::estimators does (src/setup.cpp:1343-1367), feeds them path configurations, and writes the reference's
// output files (OUTPUT/ce-ssfq-*.dat, ce-isf-*.dat; one row per bin, src/estimator.cpp:348-362).
//
//   pimcb_measure -N 16 -n 0.02198 -T 2.0 -P 124 --wavevector_type int --wavevector "1 0 0 0 1 0"
//                 --configs beads.bin [--bin_size 100] [--outdir OUTPUT] [--id run] [--action gsf]
//
// `--configs` is a raw little-endian file of B configurations, each double[M][N_ext][NDIM] in the reference's bead
// layout (N_ext given by --extent, default N).  --energy adds the thermodynamic energy estimator (ce-estimator-<id>.dat).  `--state f1,f2,...` measures saved text state files
// (OUTPUT/(g)ce-state-*.dat, state_file.h) instead: -N, -P and the extent are then taken from the first file, every
// state must be diagonal with that many particles and slices, and the loader's putInside is applied (pimc.cpp:1258-1268).  With --potential the total potential action, per-slice Vint and
// gradVSquared of every configuration are written to <outdir>/ce-potential-<id>.dat as well.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>

#include "action_b200.h"
#include "aziz.h"
#include "energy_estimator.h"
#include "estimator_b200.h"
#include "scattering_b200.h"
#include "virial_estimator.h"
#include "state_file.h"

static const char* arg(int argc, char** argv, const char* key, const char* def) {
    for (int i = 1; i + 1 < argc; ++i)
        if (!std::strcmp(argv[i], key)) return argv[i + 1];
    return def;
}
static bool flag(int argc, char** argv, const char* key) {
    for (int i = 1; i < argc; ++i)
        if (!std::strcmp(argv[i], key)) return true;
    return false;
}

int main(int argc, char** argv) {
    std::vector<std::string> stateFiles;
    if (const char* sf = arg(argc, argv, "--state", nullptr)) {
        std::string all(sf), item;
        for (size_t a = 0; a <= all.size(); ++a) {
            if (a == all.size() || all[a] == ',') { if (!item.empty()) stateFiles.push_back(item); item.clear(); }
            else item += all[a];
        }
    }
    PimcState first;
    if (!stateFiles.empty()) {
        std::string err;
        if (!readStateFile(stateFiles[0], first, err)) { std::cerr << stateFiles[0] << ": " << err << std::endl; return 1; }
        if (!first.isDiagonal()) { std::cerr << stateFiles[0] << ": not a diagonal configuration" << std::endl; return 1; }
    }
    const bool fromStates = !stateFiles.empty();
    const int N = fromStates ? first.numBeadsAtSlice[0] : std::atoi(arg(argc, argv, "-N", "16"));
    const double density = std::atof(arg(argc, argv, "-n", "0.02198"));
    const double T = std::atof(arg(argc, argv, "-T", "2.0"));
    int M = fromStates ? first.numTimeSlices : std::atoi(arg(argc, argv, "-P", "0"));
    double tau = std::atof(arg(argc, argv, "-t", "0.004"));
    const int extent = fromStates ? first.numWorldLines : std::atoi(arg(argc, argv, "--extent", arg(argc, argv, "-N", "16")));
    const int binSize = std::atoi(arg(argc, argv, "--bin_size", "100"));
    const std::string actionType = arg(argc, argv, "--action", "gsf");
    const char* cfgFile = arg(argc, argv, "--configs", nullptr);
    if (!cfgFile && !fromStates) {
        std::cerr << "usage: pimcb_measure -N n -n density -T temp (-P slices | -t tau) --wavevector_type T --wavevector \"...\" --configs file" << std::endl;
        return 2;
    }
    // number of time slices, forced even (src/setup.cpp:998-1012)
    if (M <= 0) {
        M = static_cast<int>(1.0 / (T * tau) + EPS);
        if (M % 2) M--;
    } else {
        if (M % 2) M--;
        tau = 1.0 / (T * M);
    }
    ConstantParameters* c = constants();
    c->initialNumParticles_ = N;
    c->numTimeSlices_ = M;
    c->tau_ = tau;
    c->T_ = T;
    c->wavevector_ = arg(argc, argv, "--wavevector", "");
    c->wavevectorType_ = arg(argc, argv, "--wavevector_type", "int");
    c->id_ = arg(argc, argv, "--id", "run");
    // ranks other than 0 of a multi-GPU run keep their (header-only) files apart: only rank 0 writes results
    const int myRank = std::atoi(arg(argc, argv, "--rank", "0"));
    communicate()->init(arg(argc, argv, "--outdir", "OUTPUT"), "ce", myRank > 0 ? c->id_ + "-rank" + std::to_string(myRank) : c->id_);

    Prism box(density, N);
    c->V_ = box.volume;
    Path path(&box, M, N, extent);
    MTRand random;
    const bool wantPotential = flag(argc, argv, "--potential");

    // --spring k: external potential V = k r^2 / 2 instead of `free` (its gradient couples to the pair forces in gradVSquared)
    FreePotential freeExternal;
    SpringPotential spring(std::atof(arg(argc, argv, "--spring", "0")));
    PotentialBase& external = std::atof(arg(argc, argv, "--spring", "0")) != 0.0 ? static_cast<PotentialBase&>(spring)
                                                                                  : static_cast<PotentialBase&>(freeExternal);
    std::unique_ptr<AzizPotential> aziz;
    LookupTable lookup;
    std::unique_ptr<LocalActionB200> action;
    if (wantPotential) {
        aziz.reset(new AzizPotential(std::atoi(arg(argc, argv, "--aziz_year", "1979")), &box));
        std::array<double, 2> VF{1.0, 1.0}, GF{0.0, 0.0};      // src/setup.cpp:1232-1253
        int period = 1;
        if (actionType == "gsf") { VF = {2.0 / 3.0, 4.0 / 3.0}; GF = {0.0, 2.0 / 9.0}; period = 2; }
        else if (actionType == "li_broughton") { GF = {1.0 / 12.0, 1.0 / 12.0}; period = 2; }
        // as Setup::action constructs LocalAction (src/setup.cpp:1258-1260)
        action.reset(new LocalActionB200(path, lookup, &external, aziz.get(), nullptr, VF, GF, true, actionType, 1.0, period));
    }

    std::vector<std::unique_ptr<EstimatorBase>> estimators;
    std::vector<const char*> names = {"static structure factor", "intermediate scattering function"};
    const bool wantEnergy = flag(argc, argv, "--energy");          // thermodynamic energy estimator on the device pair sums
    if (wantEnergy && !wantPotential) { std::cerr << "--energy needs --potential (the action object)" << std::endl; return 2; }
    if (wantEnergy) names.push_back("energy");
    // --virial [--virial_window W]: centroid-virial energy estimator (19 more scalar columns of ce-estimator-<id>.dat)
    const bool wantVirial = flag(argc, argv, "--virial");
    if (wantVirial && !wantPotential) { std::cerr << "--virial needs --potential (the action object)" << std::endl; return 2; }
    if (wantVirial) names.push_back("virial");
    c->virialWindow_ = std::atoi(arg(argc, argv, "--virial_window", "5"));
    if (c->virialWindow_ == 0) c->virialWindow_ = M;                   // src/setup.cpp:1775-1777
    // --elastic: elastic scattering (ce-es-<id>.dat); --cylinder R: cylinder S(q) of the beads within R of the z axis
    // (ce-cyl_ssf-<id>.dat).  The cylinder estimator brings its own q-set, so it gets a session of its own q-vectors:
    // combine it with the other scattering estimators only in separate runs.
    if (flag(argc, argv, "--elastic")) names.push_back("elastic scattering");
    const double maxR = std::atof(arg(argc, argv, "--cylinder", "0"));
    if (maxR > 0.0) names = {"cylinder static structure factor"};
    c->binSize_ = static_cast<uint32>(binSize);
    const char* lastScalar = wantVirial ? "virial" : "energy";
    for (const char* name : names) {
        EstimatorBase* e = estimatorFactory()->Create(name, path, action.get(), random, maxR);
        if (!e) { std::cerr << "estimator not registered: " << name << std::endl; return 1; }
        estimators.emplace_back(e);
        if (std::string(name) == lastScalar) e->addEndLine();          // the last scalar estimator closes the row (setup.cpp:1364-1366)
        e->prepare();
    }
    std::fstream* potOut = nullptr;
    if (wantPotential) {
        potOut = &communicate()->file("potential")->stream();
        (*potOut) << "# potentialAction, then Vint[0..M-1], then gradVSquared[0..M-1] per configuration" << std::endl;
    }

    // --batch B [--nranks n --rank r --idfile path]: walker batches accumulated on the device (pimcb_stage_batch +
    // pimcb_measure, no per-configuration read-back); with several ranks (one process per GPU, PIMCB_DEVICE selects it)
    // every rank takes a contiguous share of the configurations and the bins are summed onto rank 0 with ONE NCCL reduce
    // (pimcb_reduce_bins); rank 0 writes ONE row per file: the bin of all configurations.
    const int batch = std::atoi(arg(argc, argv, "--batch", "0"));
    if (batch > 0) {
        if (fromStates || wantPotential || maxR > 0.0) { std::cerr << "--batch works on --configs with the S(q)/F(q,tau) estimators" << std::endl; return 2; }
        const int nranks = std::atoi(arg(argc, argv, "--nranks", "1")), rank = std::atoi(arg(argc, argv, "--rank", "0"));
        B200Session& session = B200Session::get(path);
        if (nranks > 1) {
            const std::string idfile = arg(argc, argv, "--idfile", "pimcb_nccl_id.bin");
            char id[128];
            if (rank == 0) {
                B200Session::uniqueId(id);
                FILE* o = std::fopen((idfile + ".tmp").c_str(), "wb");
                if (!o || std::fwrite(id, 1, 128, o) != 128) { std::cerr << "cannot write " << idfile << std::endl; return 1; }
                std::fclose(o);
                std::rename((idfile + ".tmp").c_str(), idfile.c_str());
            } else {
                FILE* in = nullptr;
                for (int tries = 0; tries < 6000 && !(in = std::fopen(idfile.c_str(), "rb")); ++tries) {
                    struct timespec ts = {0, 10 * 1000 * 1000};
                    nanosleep(&ts, nullptr);
                }
                if (!in || std::fread(id, 1, 128, in) != 128) { std::cerr << "cannot read " << idfile << std::endl; return 1; }
                std::fclose(in);
            }
            session.commInit(nranks, rank, id);
        }
        FILE* fb = std::fopen(cfgFile, "rb");
        if (!fb) { std::cerr << "cannot open " << cfgFile << std::endl; return 1; }
        const size_t rec = static_cast<size_t>(M) * extent * NDIM;
        std::fseek(fb, 0, SEEK_END);
        const long total = std::ftell(fb) / static_cast<long>(rec * sizeof(double));
        const long base = total / nranks, extra = total % nranks;
        const long lo = rank * base + std::min<long>(rank, extra), hi = lo + base + (rank < extra ? 1 : 0);
        std::fseek(fb, static_cast<long>(lo * rec * sizeof(double)), SEEK_SET);
        std::vector<double> buf(rec * batch);
        long mine = 0;
        session.initBins();                      // a rank with an empty share still joins the reduce with a zero bin
        for (long k = lo; k < hi; k += batch) {
            const int nb = static_cast<int>(std::min<long>(batch, hi - k));
            if (std::fread(buf.data(), sizeof(double), rec * nb, fb) != rec * nb) { std::cerr << "short read" << std::endl; return 1; }
            session.measureBatch(buf.data(), nb, M, N, extent);
            mine += nb;
        }
        std::fclose(fb);
        long count = mine;
        if (nranks > 1) count = session.reduceBins(0);
        if (rank == 0) {
            std::vector<double> ssf, isf;
            long n = 0;
            session.readBins(ssf, isf, n);
            static_cast<StaticStructureFactorEstimatorB200*>(estimators[0].get())->addBin(ssf.data(), static_cast<uint32>(n));
            static_cast<IntermediateScatteringFunctionEstimatorB200*>(estimators[1].get())->addBin(isf.data(), static_cast<uint32>(n));
            estimators[0]->output();
            estimators[1]->output();
            count = n;
        }
        std::cout << "pimcb_measure: rank " << rank << "/" << nranks << " measured " << mine << " configurations"
                  << (rank == 0 ? ", bin of " + std::to_string(count) : std::string()) << std::endl;
        estimators.clear();
        B200Session::shutdown();
        return 0;
    }

    FILE* f = fromStates ? nullptr : std::fopen(cfgFile, "rb");
    if (!fromStates && !f) { std::cerr << "cannot open " << cfgFile << std::endl; return 1; }
    const size_t count = static_cast<size_t>(M) * extent * NDIM;
    long nconf = 0;
    size_t nextState = 0;
    // next configuration into path.beads: a raw record, or a parsed state file (left-packed, wrapped into the cell)
    auto nextConfiguration = [&]() -> bool {
        if (!fromStates) return std::fread(path.beads_data(), sizeof(double), count, f) == count;
        if (nextState >= stateFiles.size()) return false;
        PimcState st;
        std::string err;
        const std::string& name = stateFiles[nextState++];
        if (!readStateFile(name, st, err)) { std::cerr << name << ": " << err << std::endl; std::exit(EXIT_FAILURE); }
        if (!st.isDiagonal() || st.numTimeSlices != M || st.numBeadsAtSlice[0] != N || st.numWorldLines != extent) {
            std::cerr << name << ": not a diagonal configuration of " << N << " particles x " << M << " slices (extent "
                      << extent << ")" << std::endl;
            std::exit(EXIT_FAILURE);
        }
        if ((wantEnergy || wantVirial) && !st.linksClosed()) {
            std::cerr << name << ": world-line links are not closed over the active beads (the kinetic / virial estimators walk them)" << std::endl;
            std::exit(EXIT_FAILURE);
        }
        if (!st.isLeftPacked()) st.leftPack();
        st.putInside(box);
        std::memcpy(path.beads_data(), st.beads.data(), count * sizeof(double));
        path.setLinks(st.nextLink);                                   // world-line connectivity for the kinetic estimator
        return true;
    };
    while (nextConfiguration()) {
        B200Session::newConfiguration(path);         // the per-step hook
        for (auto& e : estimators) e->sample();      // src/pimc.cpp:737-738
        ++nconf;
        if (wantPotential) {
            (*potOut) << pimcb_format("%24.16E", action->potentialAction());
            for (int s = 0; s < M; ++s) (*potOut) << pimcb_format("%24.16E", action->potential(s)[1]);
            for (int s = 0; s < M; ++s) (*potOut) << pimcb_format("%24.16E", action->gradVSquared(s));
            (*potOut) << std::endl;
        }
        if (estimators[0]->getNumAccumulated() >= static_cast<uint32>(binSize))     // src/pimc.cpp:743-749
            for (auto& e : estimators) e->output();
    }
    if (f) std::fclose(f);
    if (estimators[0]->getNumAccumulated() > 0)
        for (auto& e : estimators) e->output();
    std::cout << "pimcb_measure: " << nconf << " configurations, N=" << N << " M=" << M << " tau=" << tau << std::endl;
    estimators.clear();
    action.reset();
    B200Session::shutdown();
    return 0;
}
