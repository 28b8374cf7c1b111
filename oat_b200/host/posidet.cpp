// oat-posidet -- `oat posidet TYPE SOURCE SINK [CONFIGURATION]` for the hot-path TYPEs, computing on the
// B200 through the C ABI.  Mirrors src/positiondetector/{main.cpp, PositionDetector.{h,cpp},
// HSVDetector.{h,cpp}, DetectorFunc.{h,cpp}}.
//
//   hsv     inRange -> erode -> dilate -> siftContours on an HSV frame SOURCE (HSVDetector.cpp:142-173)
//   track   (extension) the fused device path: takes the RAW BGR SOURCE and does framefilt mog ->
//           framefilt col -C HSV -> posidet hsv in one pass, no shared-memory hops in between
#include <cfloat>
#include <deque>
#include <iostream>
#include <memory>

#include "gpu.h"
#include "oat_cli.h"
#include "oat_host.h"

namespace oat {

class PositionDetector : public Component {
public:
    PositionDetector(const std::string &source, const std::string &sink) : frame_source_address_(source), position_sink_address_(sink) {}
    std::string name() const override { return name_; }
    virtual std::vector<config::OptionSpec> options() const = 0;
    virtual void applyConfiguration(const config::VariableMap &vm, const config::OptionTable &t) = 0;

protected:
    // PositionDetector::connectToNode (PositionDetector.cpp:40-56)
    bool connectToNode() override
    {
        frame_source_.touch(frame_source_address_);
        ctx_.reset(new gpu::Context(gpu_index_));  // after touch(), before connect(): see FrameFilter::connectToNode
        const SourceState rc = required_color_ == PIX_ANY ? frame_source_.connect() : frame_source_.connect(required_color_);
        if (rc != SourceState::CONNECTED) return false;
        in_ = frame_source_.parameters();
        position_sink_.bind(position_sink_address_, position_sink_address_);
        shared_position_ = position_sink_.retrieve();
        require_same_device(frame_source_.header(), gpu_index_, name());
        if (frame_source_.header()->memory == FrameMemory::DEVICE)  // device -> device hand-off
            src_dev_.reset(new gpu::IpcImport(*ctx_, frame_source_.header()->ipc_handle));
        else
            src_pin_.reset(new gpu::HostRegistration(frame_source_.pixels(), in_.bytes));
        d_in_.reset(new gpu::DeviceBuffer(*ctx_, in_.bytes));
        src_memory_ = frame_source_.header()->memory;
        setup();
        return true;
    }
    // PositionDetector::process (PositionDetector.cpp:58-99)
    int process() override
    {
        Position2D internal_pos("");
        if (frame_source_.wait() == NodeState::END) return 1;
        if (frame_source_.header()->memory != src_memory_)
            throw std::runtime_error("SOURCE frame memory kind changed after connect()");
        gpu::ck(oat_memcpy(ctx_->h, d_in_->p, src_dev_ ? static_cast<const uint8_t *>(src_dev_->p) + frame_source_.header()->device_offset : static_cast<const uint8_t *>(frame_source_.pixels()), in_.bytes));
        internal_pos.set_sample(frame_source_.retrieve()->sample());  // propagate tick / usec (:80)
        frame_source_.post();

        detectPosition(d_in_->u8(), internal_pos);

        position_sink_.wait();
        *shared_position_ = internal_pos;  // everything but the label (Position2D.h:84-105)
        position_sink_.post();
        return 0;
    }
    virtual void setup() = 0;
    // device frame in -> Position2D::position / position_valid
    virtual void detectPosition(const uint8_t *d_frame, Position2D &position) = 0;

    std::string name_;
    std::string frame_source_address_, position_sink_address_;
    PixelColor required_color_{PIX_ANY};
    Source<Frame> frame_source_;
    Sink<Position2D> position_sink_;
    Position2D *shared_position_{nullptr};
    FrameParams in_;
    FrameMemory src_memory_{FrameMemory::HOST_SHM};  // what the SOURCE said when this component connected
    int gpu_index_{0};
    std::unique_ptr<gpu::Context> ctx_;
    std::unique_ptr<gpu::HostRegistration> src_pin_;
    std::unique_ptr<gpu::IpcImport> src_dev_;
    std::unique_ptr<gpu::DeviceBuffer> d_in_;
};

// options shared by `hsv` and `track` (HSVDetector.cpp:49-140)
struct HSVOptions {
    oat_hsv_params p;
    HSVOptions() { oat_hsv_default_params(&p); }  // erode off, dilate 10 (:42-43), bands [0,256], area [0, DBL_MAX)
    static std::vector<config::OptionSpec> options()
    {
        return {{"h-thresh", 'H', true, "Array of ints between 0 and 256, [min,max], specifying the hue passband."},
                {"s-thresh", 'S', true, "Array of ints between 0 and 256, [min,max], specifying the saturation passband."},
                {"v-thresh", 'V', true, "Array of ints between 0 and 256, [min,max], specifying the value passband."},
                {"erode", 'e', true, "Contour erode kernel size in pixels (normalized box filter)."},
                {"dilate", 'd', true, "Contour dilation kernel size in pixels (normalized box filter)."},
                {"area", 'a', true, "Array of floats, [min,max], specifying the minimum and maximum object contour area in pixels^2."},
                {"gpu-index", 0, true, "Index of the GPU to use."}};
    }
    void apply(const config::VariableMap &vm, const config::OptionTable &t)
    {
        std::vector<int> v;
        auto band = [&](const char *key, int &lo, int &hi) {
            if (config::getArray<int>(vm, t, key, v, 2)) {
                lo = v[0];
                hi = v[1];
                if (lo < 0 || lo > 256 || hi < 0 || hi > 256)
                    throw std::runtime_error(std::string("Values of ") + key + " should be between 0 and 256.");
            }
        };
        band("h-thresh", p.h_min, p.h_max);
        band("s-thresh", p.s_min, p.s_max);
        band("v-thresh", p.v_min, p.v_max);
        config::getNumericValue<int>(vm, t, "erode", p.erode_px, 0, 1 << 20);
        config::getNumericValue<int>(vm, t, "dilate", p.dilate_px, 0, 1 << 20);
        std::vector<double> area;
        if (config::getArray<double>(vm, t, "area", area, 2)) {
            p.min_area = area[0];
            p.max_area = area[1];
            if (p.min_area >= p.max_area) throw std::runtime_error("Max area should be larger than min area.");
        }
    }
};

static void fill(Position2D &position, const oat_detection &d)
{
    position.position_valid = d.position_valid != 0;  // DetectorFunc.cpp:46, :58-60
    if (d.position_valid) {
        position.position.x = d.x;
        position.position.y = d.y;
    }
}

// ---- posidet hsv ------------------------------------------------------------------------------------------
class HSVDetector : public PositionDetector {
public:
    HSVDetector(const std::string &source, const std::string &sink) : PositionDetector(source, sink)
    {
        name_ = "hsvdetector[" + source + "->" + sink + "]";
        required_color_ = PIX_HSV;  // HSVDetector.cpp:46
    }
    ~HSVDetector() { oat_hsvdet_destroy(det_); }
    std::vector<config::OptionSpec> options() const override { return HSVOptions::options(); }
    void applyConfiguration(const config::VariableMap &vm, const config::OptionTable &t) override
    {
        o_.apply(vm, t);
        config::getNumericValue<int>(vm, t, "gpu-index", gpu_index_, 0, 1 << 20);
    }

protected:
    void setup() override { gpu::ck(oat_hsvdet_create(ctx_->h, (int)in_.rows, (int)in_.cols, &det_)); }
    void detectPosition(const uint8_t *d_frame, Position2D &position) override
    {
        oat_detection d;
        gpu::ck(oat_hsvdet_detect(det_, d_frame, in_.cols * 3, &o_.p, &d, nullptr, 0, nullptr));
        fill(position, d);
    }

private:
    HSVOptions o_;
    oat_hsvdet *det_{nullptr};
};

// ---- posidet thresh (SimpleThreshold.{h,cpp}) ---------------------------------------------------------------
class SimpleThreshold : public PositionDetector {
public:
    SimpleThreshold(const std::string &source, const std::string &sink) : PositionDetector(source, sink)
    {
        name_ = "threshdetector[" + source + "->" + sink + "]";
        required_color_ = PIX_GREY;  // SimpleThreshold.cpp:46
    }
    ~SimpleThreshold() { oat_hsvdet_destroy(det_); }
    std::vector<config::OptionSpec> options() const override
    {
        return {{"thresh", 'T', true, "Array of ints between 0 and 256, [min,max], specifying the intensity passband."},
                {"erode", 'e', true, "Contour erode kernel size in pixels (normalized box filter)."},
                {"dilate", 'd', true, "Contour dilation kernel size in pixels (normalized box filter)."},
                {"area", 'a', true, "Array of floats, [min,max], specifying the minimum and maximum object contour area in pixels^2."},
                {"gpu-index", 0, true, "Index of the GPU to use."}};
    }
    void applyConfiguration(const config::VariableMap &vm, const config::OptionTable &t) override
    {
        std::vector<int> v;
        if (config::getArray<int>(vm, t, "thresh", v, 2)) {
            t_min_ = v[0];
            t_max_ = v[1];
            if (t_min_ < 0 || t_min_ > 256 || t_max_ < 0 || t_max_ > 256)
                throw std::runtime_error("Values of thresh should be between 0 and 256.");
        }
        o_.apply(vm, t);  // erode / dilate / area share HSVDetector's semantics (SimpleThreshold.cpp:86-110)
        config::getNumericValue<int>(vm, t, "gpu-index", gpu_index_, 0, 1 << 20);
    }

protected:
    void setup() override { gpu::ck(oat_hsvdet_create(ctx_->h, (int)in_.rows, (int)in_.cols, &det_)); }
    void detectPosition(const uint8_t *d_frame, Position2D &position) override
    {
        oat_detection d;
        gpu::ck(oat_thresh_detect(det_, d_frame, in_.cols, t_min_, t_max_, &o_.p, &d, nullptr, 0, nullptr));
        fill(position, d);
    }

private:
    HSVOptions o_;
    int t_min_{0}, t_max_{256};
    oat_hsvdet *det_{nullptr};
};

// ---- posidet diff (DifferenceDetector.{h,cpp}) ----------------------------------------------------------------
class DifferenceDetector : public PositionDetector {
public:
    DifferenceDetector(const std::string &source, const std::string &sink) : PositionDetector(source, sink)
    {
        name_ = "diffdetector[" + source + "->" + sink + "]";
        required_color_ = PIX_GREY;  // DifferenceDetector.cpp:44
    }
    ~DifferenceDetector() { oat_diffdet_destroy(det_); }
    std::vector<config::OptionSpec> options() const override
    {
        return {{"diff-threshold", 'd', true, "Intensity difference threshold to consider an object contour."},
                {"blur", 'b', true, "Blurring kernel size in pixels (normalized box filter)."},
                {"area", 'a', true, "Array of floats, [min,max], specifying the minimum and maximum object contour area in pixels^2."},
                {"gpu-index", 0, true, "Index of the GPU to use."}};
    }
    void applyConfiguration(const config::VariableMap &vm, const config::OptionTable &t) override
    {
        config::getNumericValue<int>(vm, t, "diff-threshold", diff_threshold_, 0, 1 << 20);
        config::getNumericValue<int>(vm, t, "blur", blur_, 0, 1 << 20);
        std::vector<double> area;
        if (config::getArray<double>(vm, t, "area", area, 2)) {
            min_area_ = area[0];
            max_area_ = area[1];
            if (min_area_ >= max_area_) throw std::runtime_error("Max area should be larger than min area.");
        }
        config::getNumericValue<int>(vm, t, "gpu-index", gpu_index_, 0, 1 << 20);
    }

protected:
    void setup() override { gpu::ck(oat_diffdet_create(ctx_->h, (int)in_.rows, (int)in_.cols, &det_)); }
    void detectPosition(const uint8_t *d_frame, Position2D &position) override
    {
        oat_detection d;
        gpu::ck(oat_diffdet_detect(det_, d_frame, in_.cols, diff_threshold_, blur_, min_area_, max_area_, &d, nullptr, 0));
        fill(position, d);
    }

private:
    int diff_threshold_{10}, blur_{2};  // DifferenceDetector.h:74, .cpp:41
    double min_area_{0.0}, max_area_{DBL_MAX};
    oat_diffdet *det_{nullptr};
};

// ---- posidet track: mog -> col HSV -> hsv fused on the device ----------------------------------------------
// --pipeline 1 (default): one frame at a time, as the reference's process() loop.
// --pipeline D >= 2: the component keeps up to D frames on the GPU behind its SOURCE through the STREAMING resident
// engine (oat_tracker_stream_*, chunks of D/2 frames: one fused launch + one tail-server launch per chunk).  A frame
// goes back to the SOURCE as soon as its pixels are the tracker's -- copied into its HBM (H2D from page-locked shm,
// D2D from a device frame), or at once when the SINK declared its frames persistent (a static test image, a clip
// preloaded into HBM: read in place, no copy at all) -- and its position is published later, in order, one token at
// a time, every Sample intact.  The SOURCE and the SINK both see the reference's lock-step protocol (one token in
// flight per edge); what changes is that the host does nothing per frame but move tokens.  This is the GPU-aware
// counterpart of putting an `oat buffer` (src/buffer/FrameBuffer.cpp:56-116) in front of a slow component.
class FusedTracker : public PositionDetector {
public:
    FusedTracker(const std::string &source, const std::string &sink) : PositionDetector(source, sink)
    {
        name_ = "tracker[" + source + "->" + sink + "]";
        required_color_ = PIX_BGR;
    }
    ~FusedTracker() { oat_tracker_destroy(trk_); }
    std::vector<config::OptionSpec> options() const override
    {
        auto o = HSVOptions::options();
        o.push_back({"adaptation-coeff", 'A', true, "framefilt mog's adaptation coefficient, 0 to 1.0. Default 0."});
        o.push_back({"pipeline", 'p', true, "Frames kept in flight on the GPU, 1 to 64 (1 = synchronous; positions lag by up to pipeline-1 frames)."});
        return o;
    }
    void applyConfiguration(const config::VariableMap &vm, const config::OptionTable &t) override
    {
        o_.apply(vm, t);
        config::getNumericValue<double>(vm, t, "adaptation-coeff", learning_coeff_, 0.0, 1.0);
        config::getNumericValue<int>(vm, t, "gpu-index", gpu_index_, 0, 1 << 20);
        config::getNumericValue<int>(vm, t, "pipeline", depth_, 1, 64);
    }

protected:
    void setup() override { gpu::ck(oat_tracker_create(ctx_->h, (int)in_.rows, (int)in_.cols, nullptr, depth_ < 2 ? 1 : depth_, &trk_)); }
    void detectPosition(const uint8_t *, Position2D &) override {}  // (process() below drives the tracker itself)
    void publish(const oat_detection &d)
    {
        Position2D internal_pos("");
        internal_pos.set_sample(samples_.front());  // propagate tick / usec (PositionDetector.cpp:80)
        samples_.pop_front();
        fill(internal_pos, d);
        position_sink_.wait();
        *shared_position_ = internal_pos;  // everything but the label (Position2D.h:84-105)
        position_sink_.post();
    }
    const uint8_t *source_pixels() const
    {
        // the frame goes to the GPU straight from where the SOURCE keeps it (page-locked shm or device memory)
        return src_dev_ ? static_cast<const uint8_t *>(src_dev_->p) + frame_source_.header()->device_offset
                        : static_cast<const uint8_t *>(frame_source_.pixels());
    }
    int process_one_at_a_time()
    {
        if (frame_source_.wait() == NodeState::END) return 1;
        if (frame_source_.header()->memory != src_memory_)
            throw std::runtime_error("SOURCE frame memory kind changed after connect()");
        gpu::ck(oat_tracker_submit(trk_, source_pixels(), in_.cols * 3, learning_coeff_, &o_.p, nullptr, 0));
        samples_.push_back(frame_source_.retrieve()->sample());
        gpu::ck(oat_tracker_wait_ingest(trk_));
        frame_source_.post();
        oat_detection d;
        gpu::ck(oat_tracker_collect(trk_, &d));
        publish(d);
        return 0;
    }
    void publish_finished(bool block)
    {
        oat_detection d[64];
        size_t got = 0;
        gpu::ck(oat_tracker_stream_poll(trk_, d, nullptr, 64, block ? 1 : 0, &got));
        clk_.lap(4);
        for (size_t i = 0; i < got && !quit; ++i) publish(d[i]);
        clk_.lap(5);
    }
    int process() override
    {
        if (depth_ < 2) return process_one_at_a_time();
        // 1. a frame, if the SOURCE has one (block only when nothing is outstanding: there is nothing else to do)
        NodeState state;
        bool have;
        clk_.start();
        if (samples_.empty()) {
            state = frame_source_.wait();
            have = true;
        } else {
            // frames are outstanding: look for a token for a few microseconds (a running SINK answers a post() within
            // one), but do not sleep on it -- there are chunks to launch and positions to publish meanwhile
            have = frame_source_.try_wait(&state);
            for (int spin = 0; !have && state != NodeState::END && spin < 48; ++spin) {
                detail::cpu_relax();
                have = frame_source_.try_wait(&state);
            }
        }
        clk_.lap(0);
        if (state == NodeState::END) {
            while (!samples_.empty() && !quit) publish_finished(true);  // drain: every frame that went in comes out
            clk_.report(name_);
            return 1;
        }
        if (have) {
            if (frame_source_.header()->memory != src_memory_)
                throw std::runtime_error("SOURCE frame memory kind changed after connect()");
            const bool in_place = src_dev_ && frame_source_.header()->persistent;
            gpu::ck(oat_tracker_stream_push(trk_, source_pixels(), in_.cols * 3, learning_coeff_, &o_.p, in_place ? 0u : OAT_STREAM_COPY));
            samples_.push_back(frame_source_.retrieve()->sample());
            clk_.lap(1);
            if (!in_place) {
                // the frame is on its way into the tracker's memory: launch what has been gathered in the copy's shadow
                // (if a chunk slot is free -- otherwise the GPU is the bottleneck and chunks grow), then hand it back
                gpu::ck(oat_tracker_stream_flush(trk_, 0));
                gpu::ck(oat_tracker_stream_wait_ingest(trk_));
            }
            frame_source_.post();
            clk_.lap(2);
            ++clk_.n;
        } else {
            gpu::ck(oat_tracker_stream_flush(trk_, 0));  // nothing waiting: the GPU gets what has been gathered so far
            clk_.lap(3);
        }
        // 2. the positions of whatever has finished
        publish_finished(false);
        return 0;
    }

private:
    HSVOptions o_;
    double learning_coeff_{0.0};
    int depth_{1};
    std::deque<Sample> samples_;  // Samples of the frames in flight, oldest first
    oat_tracker *trk_{nullptr};
    StageClock<6> clk_{"source wait/poll", "push", "ingest + post", "flush", "poll results", "publish"};
};

}  // namespace oat

static void printUsage(std::ostream &out)
{
    out << "Usage: posidet [INFO]\n"
           "   or: posidet TYPE SOURCE SINK [CONFIGURATION]\n"
           "Perform object detection on frames from SOURCE. Publish detected object positions to SINK.\n\n"
           "TYPE\n"
           "  hsv: Object detection using color thresholding (requires an HSV frame SOURCE)\n"
           "  thresh: Object detection using intensity thresholding (requires a GREY frame SOURCE)\n"
           "  diff: Difference detector (grey-scale, motion)\n"
           "  track: fused mog + HSV conversion + hsv detection on a BGR frame SOURCE\n\n"
           "SOURCE:\n  User-supplied name of the memory segment to receive frames from (e.g. raw).\n\n"
           "SINK:\n  User-supplied name of the memory segment to publish detected positions to (e.g. pos).\n\n"
           "INFO:\n  --help                 Produce help message.\n  -v [ --version ]       Print version information.\n\n"
           "CONFIGURATION:\n  -c [ --config ] FILE KEY   Configuration file/key pair.\n";
}

int main(int argc, char *argv[])
{
    using namespace oat;
    std::string comp_name = "posidet";
    try {
        for (int i = 1; i < argc; ++i) {
            const std::string a = argv[i];
            if (a == "--help" && argc == 2) { printUsage(std::cout); return 0; }
            if (a == "-v" || a == "--version") { std::cout << "Oat Object Position Detector (B200) version 0.1\n"; return 0; }
        }
        if (argc < 2) { printUsage(std::cout); return 0; }
        const std::string type = argv[1];
        std::vector<std::string> pos;
        for (int i = 2; i < argc && pos.size() < 2; ++i) {
            if (argv[i][0] == '-') break;
            pos.push_back(argv[i]);
        }
        if (type != "hsv" && type != "track" && type != "thresh" && type != "diff") {
            printUsage(std::cout);
            std::cerr << whoError(comp_name, "Error: invalid TYPE specified.\n");
            return -1;
        }
        if (pos.size() < 1) { printUsage(std::cout); std::cerr << whoError(comp_name, "Error: a SOURCE must be specified.\n"); return -1; }
        if (pos.size() < 2) { printUsage(std::cout); std::cerr << whoError(comp_name, "Error: a SINK name must be specified.\n"); return -1; }
        std::shared_ptr<PositionDetector> detector;
        if (type == "hsv") detector = std::make_shared<HSVDetector>(pos[0], pos[1]);
        else if (type == "thresh") detector = std::make_shared<SimpleThreshold>(pos[0], pos[1]);
        else if (type == "diff") detector = std::make_shared<DifferenceDetector>(pos[0], pos[1]);
        else detector = std::make_shared<FusedTracker>(pos[0], pos[1]);
        comp_name = detector->name();

        auto opts = detector->options();
        opts.push_back({"config", 'c', true, "Configuration file/key pair."});
        opts.push_back({"help", 0, false, ""});
        const config::VariableMap vm = config::parse(argc, argv, 4, opts);
        if (vm.count("help")) {
            printUsage(std::cout);
            for (const auto &o : detector->options()) std::cout << "  --" << o.long_name << "  " << o.help << "\n";
            return 0;
        }
        config::OptionTable table;
        if (vm.count("config")) {
            table = config::getConfigTable(vm.values.at("config"), vm.values.at("config-key"));
            config::checkKeys(detector->options(), table);
        }
        detector->applyConfiguration(vm, table);

        std::cout << whoMessage(comp_name, "Listening to source " + pos[0] + ".\n")
                  << whoMessage(comp_name, "Steaming to sink " + pos[1] + ".\n")
                  << whoMessage(comp_name, "Press CTRL+C to exit.\n");
        detector->run();
        std::cout << whoMessage(comp_name, "Exiting.\n");
        return 0;
    } catch (const std::exception &ex) {
        std::cerr << whoError(comp_name, ex.what()) << std::endl;
    } catch (...) {
        std::cerr << whoError(comp_name, "Unknown exception.") << std::endl;
    }
    return -1;
}
