// oat-posisock -- `oat posisock std SOURCE`: print every position as one JSON object per line, byte for
// byte what src/positionsocket/PositionCout.cpp:53-67 + serializePosition (lib/datatypes/Position2D.h:
// 169-233) emit.  With --npy FILE the 82-byte packed records of the reference's .npy writer
// (lib/datatypes/Position2D.cpp:24-96) are appended to FILE instead.
#include <fstream>
#include <iostream>

#include "oat_cli.h"
#include "oat_host.h"

int main(int argc, char *argv[])
{
    using namespace oat;
    const std::string comp_name = "posisock";
    try {
        if (argc < 3 || std::string(argv[1]) != "std") {
            std::cout << "Usage: posisock std SOURCE [--npy FILE]\n";
            return argc < 2 ? 0 : -1;
        }
        struct Owner : Component {
            std::string name() const override { return "posicout"; }
            bool connectToNode() override { return true; }
            int process() override { return 1; }
        } sig_owner;
        const config::VariableMap vm = config::parse(argc, argv, 3, {{"npy", 0, true, ""}});
        std::ofstream npy;
        if (vm.count("npy")) npy.open(vm.values.at("npy"), std::ios::binary);
        Source<Position2D> source;
        source.touch(argv[2]);
        if (source.connect() != SourceState::CONNECTED) return 0;
        Position2D p("");
        while (!quit) {
            if (source.wait() == NodeState::END) break;
            p = *source.retrieve();
            source.post();
            if (npy.is_open()) {
                char rec[Position2D::NPY_DTYPE_BYTES];
                packPosition(p, rec);
                npy.write(rec, sizeof(rec));
            } else {
                std::cout << serializePosition(p) << "\n" << std::flush;
            }
        }
        return 0;
    } catch (const std::exception &ex) {
        std::cerr << whoError(comp_name, ex.what()) << std::endl;
    }
    return -1;
}
