// oat-posifilt -- `oat posifilt kalman SOURCE SINK [--dt --timeout/-T --sigma-accel/-a --sigma-noise/-n]`
// on the B200 through the C ABI (oat_posfilt_*).  Mirrors src/positionfilter/{main.cpp,
// PositionFilter.{h,cpp}, KalmanFilter2D.{h,cpp}}; `homography` and `region` are not part of the
// hot path and are rejected as invalid TYPEs here.
#include <iostream>
#include <limits>
#include <memory>

#include "gpu.h"
#include "oat_cli.h"
#include "oat_host.h"

namespace oat {

static void printUsage(std::ostream &out)
{
    out << "Usage: posifilt [INFO]\n"
        << "   or: posifilt TYPE SOURCE SINK [CONFIGURATION]\n"
        << "Filter positions from SOURCE and published filtered positions to SINK.\n\n"
        << "TYPE\n  kalman: Kalman filter\n\n"
        << "SOURCE:\n  User-supplied name of the memory segment to receive positions from (e.g. rpos).\n\n"
        << "SINK:\n  User-supplied name of the memory segment to publish positions to (e.g. rpos).\n";
}

// PositionFilter (src/positionfilter/PositionFilter.{h,cpp})
class PositionFilter : public Component {
public:
    PositionFilter(const std::string &source, const std::string &sink) : source_address_(source), sink_address_(sink) {}
    std::string name() const override { return name_; }
    virtual std::vector<config::OptionSpec> options() const = 0;
    virtual void applyConfiguration(const config::VariableMap &vm, const config::OptionTable &t) = 0;

protected:
    bool connectToNode() override  // PositionFilter.cpp:36-50
    {
        position_source_.touch(source_address_);
        if (position_source_.connect() != SourceState::CONNECTED) return false;
        position_sink_.bind(sink_address_, sink_address_);
        shared_position_ = position_sink_.retrieve();
        return true;
    }
    int process() override  // PositionFilter.cpp:52-90
    {
        if (position_source_.wait() == NodeState::END) return 1;
        Position2D internal_position = position_source_.clone();
        position_source_.post();

        filter(internal_position);

        position_sink_.wait();
        *shared_position_ = internal_position;
        position_sink_.post();
        return 0;
    }
    virtual void filter(Position2D &position) = 0;

    std::string name_, source_address_, sink_address_;
    Source<Position2D> position_source_;
    Sink<Position2D> position_sink_;
    Position2D *shared_position_{nullptr};
};

// KalmanFilter2D (src/positionfilter/KalmanFilter2D.{h,cpp}); the tuning GUI (-t) is out of scope
class KalmanFilter2D : public PositionFilter {
public:
    KalmanFilter2D(const std::string &source, const std::string &sink) : PositionFilter(source, sink)
    {
        name_ = "kalman[" + source + "->" + sink + "]";
        oat_kalman_default_params(&kp_);
    }
    ~KalmanFilter2D() override
    {
        if (f_) oat_posfilt_destroy(f_);
    }
    std::vector<config::OptionSpec> options() const override
    {
        return {{"dt", 0, true, "Kalman filter time step in seconds."},
                {"timeout", 'T', true, "Seconds to perform position estimation detection with lack of position measure. Defaults to 0."},
                {"sigma-accel", 'a', true, "Standard deviation of normally distributed, random accelerations used by the internal model of object motion (position units/s2; e.g. pixels/s2)."},
                {"sigma-noise", 'n', true, "Standard deviation of randomly distributed position measurement noise (position units; e.g. pixels)."},
                {"gpu-index", 0, true, "Index of the GPU to use."}};
    }
    void applyConfiguration(const config::VariableMap &vm, const config::OptionTable &t) override
    {
        const double inf = std::numeric_limits<double>::max();
        config::getNumericValue<double>(vm, t, "dt", kp_.dt, 0.0, inf);  // KalmanFilter2D.cpp:69-85
        config::getNumericValue<double>(vm, t, "timeout", kp_.timeout, 0.0, inf);
        config::getNumericValue<double>(vm, t, "sigma-accel", kp_.sigma_accel, 0.0, inf);
        config::getNumericValue<double>(vm, t, "sigma-noise", kp_.sigma_noise, 0.0, inf);
        int gpu_index = 0;
        config::getNumericValue<int>(vm, t, "gpu-index", gpu_index, 0, 1 << 16);
        ctx_ = std::make_unique<gpu::Context>(gpu_index);
        gpu::ck(oat_posfilt_create(ctx_->h, 1, &kp_, 0, -1, &f_));
    }

private:
    void filter(Position2D &position) override  // KalmanFilter2D.cpp:95-145
    {
        oat_position in{}, out{};
        in.position_valid = position.position_valid ? 1 : 0;
        in.x = position.position.x;
        in.y = position.position.y;
        gpu::ck(oat_posfilt_apply(f_, &in, &out));
        position.position.x = out.x;
        position.velocity.x = out.vx;
        position.position.y = out.y;
        position.velocity.y = out.vy;
        position.position_valid = out.position_valid != 0;
        position.velocity_valid = out.velocity_valid != 0;
    }
    oat_kalman_params kp_;
    std::unique_ptr<gpu::Context> ctx_;
    oat_posfilt *f_{nullptr};
};

}  // namespace oat

int main(int argc, char *argv[])
{
    using namespace oat;
    std::string comp_name = "posifilt";
    try {
        for (int i = 1; i < argc; ++i) {
            const std::string a = argv[i];
            if (a == "--help" && argc == 2) { printUsage(std::cout); return 0; }
            if (a == "-v" || a == "--version") { std::cout << "Oat Position Filter (B200) version 0.1\n"; return 0; }
        }
        if (argc < 2) { printUsage(std::cout); return 0; }
        const std::string type = argv[1];
        std::vector<std::string> pos;
        for (int i = 2; i < argc && pos.size() < 2; ++i) {
            if (argv[i][0] == '-') break;
            pos.push_back(argv[i]);
        }
        if (type != "kalman") { printUsage(std::cout); std::cerr << whoError(comp_name, "Error: invalid TYPE specified.\n"); return -1; }
        if (pos.size() < 1) { printUsage(std::cout); std::cerr << whoError(comp_name, "Error: a SOURCE must be specified.\n"); return -1; }
        if (pos.size() < 2) { printUsage(std::cout); std::cerr << whoError(comp_name, "Error: a SINK name must be specified.\n"); return -1; }
        auto filter = std::make_shared<KalmanFilter2D>(pos[0], pos[1]);
        comp_name = filter->name();
        auto opts = filter->options();
        opts.push_back({"config", 'c', true, "Configuration file/key pair."});
        opts.push_back({"help", 0, false, ""});
        const config::VariableMap vm = config::parse(argc, argv, 4, opts);
        if (vm.count("help")) {
            printUsage(std::cout);
            for (const auto &o : filter->options()) std::cout << "  --" << o.long_name << "  " << o.help << "\n";
            return 0;
        }
        config::OptionTable table;
        if (vm.count("config")) {
            table = config::getConfigTable(vm.values.at("config"), vm.values.at("config-key"));
            config::checkKeys(filter->options(), table);
        }
        filter->applyConfiguration(vm, table);
        std::cout << whoMessage(comp_name, "Listening to source " + pos[0] + ".\n")
                  << whoMessage(comp_name, "Steaming to sink " + pos[1] + ".\n")
                  << whoMessage(comp_name, "Press CTRL+C to exit.\n");
        filter->run();
        std::cout << whoMessage(comp_name, "Exiting.\n");
        return 0;
    } catch (const std::exception &ex) {
        std::cerr << whoError(comp_name, ex.what()) << std::endl;
    } catch (...) {
        std::cerr << whoError(comp_name, "Unknown exception.") << std::endl;
    }
    return -1;
}
