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
        // .npy exactly as the reference's recorder writes it (src/recorder/Format.cpp:35-72, PositionWriter.cpp:80-82):
        // v1.0 header with a 10-digit shape placeholder that is patched at exit (emplaceNumpyShape, Format.cpp:74-93)
        static const char NPY_DTYPE[] =
            "[('tick', '<u8'),('usec', '<u8'),('unit', '<i4'),('pos_ok', '<i1'),('pos_xy', 'f8', (2)),('vel_ok', '<i1'),"
            "('vel_xy', 'f8', (2)),('head_ok', '<i1'),('head_xy', 'f8', (2)),('reg_ok', '<i1'),('reg', 'a10')]";
        std::fstream npy;
        uint64_t nrec = 0;
        if (vm.count("npy")) {
            npy.open(vm.values.at("npy"), std::ios::binary | std::ios::out | std::ios::trunc);
            std::string dict = std::string("{'shape': (0000000000, ), 'fortran_order': False, 'descr': ") + NPY_DTYPE + "}";
            const size_t rem = 16 - ((dict.size() + 10) % 16);
            dict.append(rem, ' ');
            dict.back() = '\n';
            std::string hdr = std::string("\x93NUMPY") + '\x01' + '\x00';
            hdr.push_back((char)(dict.size() & 0xff));
            hdr.push_back((char)((dict.size() >> 8) & 0xff));
            hdr += dict;
            npy.write(hdr.data(), (std::streamsize)hdr.size());
        }
        Source<Position2D> source;
        source.touch(argv[2]);
        if (source.connect() != SourceState::CONNECTED) return 0;
        Position2D p("");
        TokenClock out_clk;
        while (!quit) {
            if (source.wait() == NodeState::END) break;
            p = *source.retrieve();
            source.post();
            out_clk.tick();
            if (npy.is_open()) {
                char rec[Position2D::NPY_DTYPE_BYTES];
                packPosition(p, rec);
                npy.write(rec, sizeof(rec));
                ++nrec;
            } else {
                std::cout << serializePosition(p) << "\n" << std::flush;
            }
        }
        out_clk.report("posisock[" + std::string(argv[2]) + "]");
        if (npy.is_open()) {  // emplaceNumpyShape
            const std::string n = std::to_string(nrec);
            if (n.size() <= 10) {
                const std::string shape = "'shape': " + std::string(10 - n.size(), ' ') + "(" + n + ", ), ";
                npy.seekp(11);
                npy.write(shape.data(), (std::streamsize)shape.size());
            }
            npy.close();
        }
        return 0;
    } catch (const std::exception &ex) {
        std::cerr << whoError(comp_name, ex.what()) << std::endl;
    }
    return -1;
}
