// Host side of the drop-in: scOOP's input formats and parameter rules, kept as they are in the reference so that a
// user's options / top.init / config.init files work unchanged. This is the C++ mirror of
//   Inicializer::readTopoFile / fillTypes / fillMol / fillSystem / fillExclusions  (scOOP/mc/inicializer.cpp:453-1108)
//   Topo::genParamPairs / genTopoParams                                           (scOOP/structures/topo.cpp:5-153)
//   Inicializer::initConfig + Conf::makeMoleculeWhole + Particle::init            (inicializer.cpp:89-298, Conf.h:371-380,
//                                                                                  particle.cpp:3-79)
// and it produces the packed tables the C ABI (include/scgpu.h) takes. No energy arithmetic lives here.
#pragma once
#include <array>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../../include/scgpu.h"

namespace schost {

constexpr int MAXT = 40;      // scOOP/structures/macros.h:87
constexpr int MAXMT = 100;    // macros.h:88

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct TypeParams {           // the reference's Ia_param (structures.h:187-254), pair-path members + the angles kept for output
    int geotype[2] = {0, 0};
    double sigma = 0, epsilon = 0, A = 0, B = 0, pdis = 0, pswitch = 0, pswitchINV = 0;
    double rcut = 0, rcutSq = 0, rcutwca = 0, rcutwcaSq = 0, parallel = 0;
    double len[2] = {0, 0}, half_len[2] = {0, 0};
    double pangl[4] = {0, 0, 0, 0}, panglsw[4] = {0, 0, 0, 0}, pcangl[4] = {0, 0, 0, 0}, pcanglsw[4] = {0, 0, 0, 0};
    double pcoshalfi[4] = {0, 0, 0, 0}, psinhalfi[4] = {0, 0, 0, 0};
    double csecpatchrot[2] = {0, 0}, ssecpatchrot[2] = {0, 0}, chiral_cos[2] = {0, 0}, chiral_sin[2] = {0, 0};
    double volume = 0;
    bool exclude = false;
    std::string name;
    scgpu_iaparam pack() const;
};

struct MoleculeType {         // MoleculeParams (moleculeparams.h:10-45)
    std::string name;
    double bond1eq = -1, bond1c = -1, bond2eq = -1, bond2c = -1, bonddeq = -1, bonddc = -1, bondheq = -1, bondhc = -1;
    double angle1eq = -1, angle1c = -1, angle2eq = -1, angle2c = -1;
    std::vector<int> particleTypes;
    int molSize() const { return (int)particleTypes.size(); }
};

class Topology {
public:
    std::vector<std::vector<TypeParams>> ia;   // [MAXT][MAXT]
    std::vector<MoleculeType> mols;
    std::vector<std::pair<std::string, long>> system;   // [System] entries in file order
    std::set<std::pair<int, int>> exclusions;
    bool exterExist = false;
    double exter[3] = {0, 0, 0};
    double sqmaxcut = 0, maxcut = 0;

    Topology();
    static Topology fromText(const std::string& text);   // parse + genParamPairs + genTopoParams
    int maxTypeInUse(const std::vector<int>& types) const;

private:
    void fillType(const std::string& line);
    void fillMol(MoleculeType& mol, const std::string& line);
    void genParamPairs();
    void genTopoParams();
};

// A loaded configuration: what Conf::pvec + Conf::geo.box hold after initConfig() and partVecInit()
struct System {
    Topology topo;
    int n = 0;
    std::array<double, 3> box{{0, 0, 0}};
    std::vector<double> state;      // n * 30, the C-ABI particle record
    std::vector<int> type, moltype, switched;
    std::vector<int> first;         // ParticleVector::first[], size nmol+1
    // packed tables for scgpu_set_topology
    int ntypes = 0;
    std::vector<scgpu_iaparam> iaTable;
    std::vector<scgpu_molparam> molTable;

    static System fromText(const std::string& topText, const std::string& configText, const std::vector<long>& countsOverride = {});
    static System fromFiles(const std::string& topPath, const std::string& configPath);
    void initParticle(int i);       // Particle::init
    std::string configLast(bool testingFormat) const;   // Conf::draw + box line (main.cpp:304-311)
};

void particleInit(const TypeParams& self, double* state30);   // Particle::init (particle.cpp:3-79)

}  // namespace schost
