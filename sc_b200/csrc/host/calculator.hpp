// TotalEGpu -- the host-side mirror of the reference's energy-calculator interface
// (TotalE<pairEFce>, scOOP/mc/totalenergycalculator.h:135-297; the live subclass is TotalEMatrix<PairE>, :299-522).
// Same method names, same argument meaning, same caller protocol: the caller mutates the configuration IN PLACE
// (conf->pvec[target] or conf->geo.box), calls the ...Trial() method, then either restores the configuration itself
// (reject) or calls update(...) (accept). Every method forwards to one entry point of the C ABI (include/scgpu.h);
// there is no CPU path: construction throws when no CUDA device / library is available.
#pragma once
#include <string>
#include <vector>

#include "../../../include/scgpu.h"
#include "topology.hpp"

namespace schost {

typedef std::vector<int> Molecule;   // scOOP/structures/structures.h:15-26
struct EMResize {};                  // totalenergycalculator.h:133

class TotalEGpu {
public:
    System* conf;                    // the reference holds Conf* (+ GeoBase* through PairE::pbc)
    scgpu_ctx* ctx = nullptr;

    explicit TotalEGpu(System* conf, int device = 0);
    ~TotalEGpu();
    TotalEGpu(const TotalEGpu&) = delete;
    TotalEGpu& operator=(const TotalEGpu&) = delete;

    void initEM();                               // totalenergycalculator.h:314-316
    void update();                               // accepted volume move (:318-320)
    void update(int target);                     // accepted single-particle move (:326-328)
    void update(const Molecule& target);         // accepted chain move (:330-332)
    void update(EMResize);                       // particle count changed (:322-324)

    double allToAll();                           // :338-353
    double allToAllTrial();                      // :334-336
    double oneToAll(int target);                 // :355-381  (state BEFORE the caller's mutation = device state)
    double oneToAllTrial(int target);            // :383-415  (host state of `target` is the trial state)
    double mol2others(const Molecule& mol);      // :417-453
    double mol2othersTrial(const Molecule& mol); // :455-499
    double p2p(int part1, int part2);            // :200-203
    int overlapAll(int target, int variant = 0); // Conf::overlapAll (Conf.cpp:244-253) on the host state of target
    int checkall(int variant = 0);               // Conf::checkall (Conf.cpp:256-267)

private:
    void check(int rc, const char* what) const;
    void pushBox();
    std::vector<double> scratch;
};

}  // namespace schost
