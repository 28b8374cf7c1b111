// oat-posicom -- `oat posicom mean SOURCES SINK [--heading-anchor/-h IDX]` on the B200 through the C ABI
// (oat_posfilt_* with combine_mean).  Mirrors src/positioncombiner/{main.cpp, PositionCombiner.{h,cpp},
// MeanPosition.{h,cpp}}.
#include <iostream>
#include <memory>

#include "gpu.h"
#include "oat_cli.h"
#include "oat_host.h"

namespace oat {

static void printUsage(std::ostream &out)
{
    out << "Usage: posicom [INFO]\n"
        << "   or: posicom TYPE SOURCES SINK [CONFIGURATION]\n"
        << "Combine positional information from two or more SOURCES and Publish combined position to SINK.\n\n"
        << "TYPE\n  mean: Geometric mean of positions\n\n"
        << "SOURCES:\n  User-supplied position source names (e.g. pos1 pos2).\n\n"
        << "SINK:\n  User-supplied position sink name (e.g. pos).\n";
}

// PositionCombiner + MeanPosition (src/positioncombiner/PositionCombiner.cpp:36-130, MeanPosition.cpp:33-118)
class MeanPosition : public Component {
public:
    std::string name() const override { return name_; }
    ~MeanPosition() override
    {
        if (f_) oat_posfilt_destroy(f_);
    }
    static std::vector<config::OptionSpec> options()
    {
        return {{"heading-anchor", 'h', true,
                 "Index of the SOURCE position to use as an anchor when calculating object heading. In this case the heading "
                 "equals the mean directional vector between this anchor position and all other SOURCE positions. If "
                 "unspecified, the heading is not calculated."},
                {"gpu-index", 0, true, "Index of the GPU to use."}};
    }
    void applyConfiguration(const config::VariableMap &vm, const config::OptionTable &t)
    {
        std::vector<std::string> sources = vm.positional;  // resolvePositionSources, PositionCombiner.cpp:36-61
        if (sources.size() < 3) throw std::runtime_error("At least two SOURCES and a SINK must be specified.");
        if (sources.size() > 9) throw std::runtime_error("At most 8 SOURCES are supported.");
        sink_address_ = sources.back();
        sources.pop_back();
        name_ = "posicom[" + sources[0] + "...->" + sink_address_ + "]";
        for (auto &addr : sources) {
            addresses_.push_back(addr);
            positions_.emplace_back(addr);
            position_sources_.push_back(std::make_unique<Source<Position2D>>());
        }
        int anchor = -1;  // MeanPosition.cpp:52-54
        config::getNumericValue<int>(vm, t, "heading-anchor", anchor, 0, (int)sources.size() - 1);
        int gpu_index = 0;
        config::getNumericValue<int>(vm, t, "gpu-index", gpu_index, 0, 1 << 16);
        ctx_ = std::make_unique<gpu::Context>(gpu_index);
        gpu::ck(oat_posfilt_create(ctx_->h, (int)sources.size(), nullptr, 1, anchor, &f_));
    }

protected:
    bool connectToNode() override  // PositionCombiner.cpp:63-91
    {
        for (size_t i = 0; i < position_sources_.size(); ++i) position_sources_[i]->touch(addresses_[i]);
        for (auto &ps : position_sources_)
            if (ps->connect() != SourceState::CONNECTED) return false;
        position_sink_.bind(sink_address_, sink_address_);
        shared_position_ = position_sink_.retrieve();
        return true;
    }
    int process() override  // PositionCombiner.cpp:93-128
    {
        for (size_t i = 0; i < position_sources_.size(); ++i) {
            if (position_sources_[i]->wait() == NodeState::END) return 1;
            positions_[i] = position_sources_[i]->clone();
            position_sources_[i]->post();
        }
        combine();
        position_sink_.wait();
        *shared_position_ = internal_position_;
        position_sink_.post();
        return 0;
    }
    void combine()  // MeanPosition::combine, MeanPosition.cpp:60-118
    {
        oat_position in[8] = {}, out{};
        for (size_t i = 0; i < positions_.size(); ++i) {
            const Position2D &p = positions_[i];
            in[i].position_valid = p.position_valid;
            in[i].velocity_valid = p.velocity_valid;
            in[i].heading_valid = p.heading_valid;
            in[i].x = p.position.x;
            in[i].y = p.position.y;
            in[i].vx = p.velocity.x;
            in[i].vy = p.velocity.y;
            in[i].hx = p.heading.x;
            in[i].hy = p.heading.y;
        }
        gpu::ck(oat_posfilt_apply(f_, in, &out));
        internal_position_.position_valid = out.position_valid != 0;
        internal_position_.velocity_valid = out.velocity_valid != 0;
        internal_position_.heading_valid = out.heading_valid != 0;
        internal_position_.position = {out.x, out.y};
        internal_position_.velocity = {out.vx, out.vy};
        internal_position_.heading = {out.hx, out.hy};
    }

    std::string name_{"posicom"}, sink_address_;
    std::vector<std::string> addresses_;
    std::vector<Position2D> positions_;
    std::vector<std::unique_ptr<Source<Position2D>>> position_sources_;
    Position2D internal_position_{""};
    Sink<Position2D> position_sink_;
    Position2D *shared_position_{nullptr};
    std::unique_ptr<gpu::Context> ctx_;
    oat_posfilt *f_{nullptr};
};

}  // namespace oat

int main(int argc, char *argv[])
{
    using namespace oat;
    std::string comp_name = "posicom";
    try {
        for (int i = 1; i < argc; ++i) {
            const std::string a = argv[i];
            if (a == "--help" && argc == 2) { printUsage(std::cout); return 0; }
            if (a == "-v" || a == "--version") { std::cout << "Oat Position Combiner (B200) version 0.1\n"; return 0; }
        }
        if (argc < 2) { printUsage(std::cout); return 0; }
        if (std::string(argv[1]) != "mean") { printUsage(std::cout); std::cerr << whoError(comp_name, "Error: invalid TYPE specified.\n"); return -1; }
        auto opts = MeanPosition::options();
        opts.push_back({"config", 'c', true, "Configuration file/key pair."});
        opts.push_back({"help", 0, false, ""});
        const config::VariableMap vm = config::parse(argc, argv, 2, opts);
        if (vm.count("help")) {
            printUsage(std::cout);
            for (const auto &o : MeanPosition::options()) std::cout << "  --" << o.long_name << "  " << o.help << "\n";
            return 0;
        }
        config::OptionTable table;
        if (vm.count("config")) {
            table = config::getConfigTable(vm.values.at("config"), vm.values.at("config-key"));
            config::checkKeys(MeanPosition::options(), table);
        }
        auto combiner = std::make_shared<MeanPosition>();
        combiner->applyConfiguration(vm, table);
        comp_name = combiner->name();
        std::cout << whoMessage(comp_name, "Press CTRL+C to exit.\n");
        combiner->run();
        std::cout << whoMessage(comp_name, "Exiting.\n");
        return 0;
    } catch (const std::exception &ex) {
        std::cerr << whoError(comp_name, ex.what()) << std::endl;
    } catch (...) {
        std::cerr << whoError(comp_name, "Unknown exception.") << std::endl;
    }
    return -1;
}
