// oat-framefilt -- `oat framefilt TYPE SOURCE SINK [CONFIGURATION]` for the hot-path TYPEs, computing
// on the B200 through the C ABI.  Mirrors src/framefilter/{main.cpp, FrameFilter.{h,cpp},
// BackgroundSubtractorMOG.{h,cpp}, ColorConvert.{h,cpp}, BackgroundSubtractor.{h,cpp}}.
//
//   mog    cv::BackgroundSubtractorMOG2::apply + frame.setTo(0, mask == 0)   (...MOG.cpp:114-127)
//   col    cv::cvtColor, BGR -> HSV                                           (ColorConvert.cpp:101-107)
//   bsub   frame - background, saturating                                     (BackgroundSubtractor.cpp:87-100)
//
// Both shared-memory mappings (the SOURCE's frame and this SINK's frame) are page-locked once, so the
// per-frame copies are asynchronous DMA straight from / into shared memory (HOST_PINNED variant);
// the SOURCE is released as soon as its pixels are on the device, as in FrameFilter::process().
#include <cstdio>
#include <chrono>
#include <iostream>
#include <memory>
#include <vector>

#include "gpu.h"
#include "oat_cli.h"
#include "oat_host.h"

namespace oat {

class FrameFilter : public Component {
public:
    FrameFilter(const std::string &source, const std::string &sink) : frame_source_address_(source), frame_sink_address_(sink) {}
    std::string name() const override { return name_; }
    virtual std::vector<config::OptionSpec> options() const = 0;
    virtual void applyConfiguration(const config::VariableMap &vm, const config::OptionTable &t) = 0;

protected:
    // FrameFilter::connectToNode (FrameFilter.cpp:37-57); sink_color lets `col` re-tag its output
    bool connectToNode() override
    {
        // touch() first (from now on the SOURCE's SINK waits for this component), then the CUDA context -- creating it
        // takes a good part of a second and must not happen after connect(), which returns with the first frame waiting
        frame_source_.touch(frame_source_address_);
        ctx_.reset(new gpu::Context(gpu_index_));
        if (frame_source_.connect() != SourceState::CONNECTED) return false;
        in_ = frame_source_.parameters();
        const PixelColor out_color = outputColor(in_.color);
        const size_t out_bytes = in_.rows * in_.cols * (size_t)color_bytes(out_color);
        frame_sink_.bind(frame_sink_address_, out_bytes, false);  // announced below, once the memory kind is final
        shared_frame_ = frame_sink_.retrieve(in_.rows, in_.cols, color_bytes(out_color), out_color);
        d_in_.reset(new gpu::DeviceBuffer(*ctx_, in_.bytes));
        d_out_.reset(new gpu::DeviceBuffer(*ctx_, out_bytes));
        if (!device_sink_) d_out2_.reset(new gpu::DeviceBuffer(*ctx_, out_bytes));
        out_bytes_ = out_bytes;
        // SOURCE side: a DEVICE frame is imported through its IPC handle (device -> device hand-off);
        // a host frame's mapping is page-locked once (HOST_PINNED) so the per-frame copy is plain DMA
        require_same_device(frame_source_.header(), gpu_index_, name());
        if (frame_source_.header()->memory == FrameMemory::DEVICE)
            src_dev_.reset(new gpu::IpcImport(*ctx_, frame_source_.header()->ipc_handle));
        else
            src_pin_.reset(new gpu::HostRegistration(frame_source_.pixels(), in_.bytes));
        // SINK side: --device-sink publishes the output buffer itself (no copy back to the host)
        if (device_sink_) {
            unsigned char handle[64];
            gpu::ck(oat_ipc_export(ctx_->h, d_out_->p, handle));
            frame_sink_.publish_device(handle, gpu_index_);
        } else {
            dst_pin_.reset(new gpu::HostRegistration(frame_sink_.pixels(), out_bytes));
            frame_sink_.set_memory(FrameMemory::HOST_PINNED, gpu_index_);
        }
        setup();
        src_memory_ = frame_source_.header()->memory;
        frame_sink_.announce();  // frame parameters, memory kind and IPC handle are in place: let SOURCEs connect
        return true;
    }
    // FrameFilter::process (FrameFilter.cpp:59-98), with the heap copies replaced by DMA
    // OAT_B200_TIMING=1: where a frame's time goes in this component (printed at end of stream)
    StageClock<5> clk_{"wait source", "ingest", "filter", "wait sink", "egress"};
    TokenClock out_clk_;
    // Host-memory SINK: the copy of frame t back into shared memory runs on the egress lane while the component
    // already waits for, ingests and filters frame t+1 (done one after the other a frame pays H2D + filter + D2H).
    // Frame t is published as soon as its copy has landed -- at the latest when the filter of frame t+1 is done, and
    // at once if no further frame is waiting: a slow SOURCE sees no added latency.  (A third stage -- the next frame
    // coming up while this one is filtered -- was measured and dropped: on the B200 hosts H2D and D2H in flight
    // together share ~60 GB/s, so two copies per frame bound the rate either way, profiles/README.md.)
    bool egress_pending_{false};
    Sample egress_sample_;
    int out_k_{0};
    void finish_egress()
    {
        gpu::ck(oat_memcpy_wait(ctx_->h, 1));
        shared_frame_.sample() = egress_sample_;  // filters never advance time
        frame_sink_.post();
        egress_pending_ = false;
        out_clk_.tick();
    }
    int process() override
    {
        clk_.start();
        bool have = false;
        while (egress_pending_) {  // the next frame, or the end of the copy: whichever comes first
            NodeState st;
            if (frame_source_.try_wait(&st)) {
                have = true;
                break;
            }
            int done = 0;
            gpu::ck(oat_memcpy_done(ctx_->h, 1, &done));
            if (done || st == NodeState::END || quit) {
                finish_egress();
                break;
            }
            detail::cpu_relax();
        }
        if (!have && frame_source_.wait() == NodeState::END) {
            clk_.report(name_);
            out_clk_.report(name_);
            return 1;
        }
        clk_.lap(0);
        if (frame_source_.header()->memory != src_memory_)
            throw std::runtime_error("SOURCE frame memory kind changed after connect()");
        gpu::ck(oat_memcpy(ctx_->h, d_in_->p, src_dev_ ? static_cast<const uint8_t *>(src_dev_->p) + frame_source_.header()->device_offset : static_cast<const uint8_t *>(frame_source_.pixels()), in_.bytes));
        const Sample sample = frame_source_.retrieve()->sample();
        frame_source_.post();
        clk_.lap(1);

        if (device_sink_) {
            // the published buffer IS d_out_: it may only change while no SOURCE is reading it
            frame_sink_.wait();
            clk_.lap(3);
            filter(d_in_->u8(), d_out_->u8());
            clk_.lap(2);
            shared_frame_.sample() = sample;  // filters never advance time
            frame_sink_.post();
            out_clk_.tick();
        } else {
            gpu::DeviceBuffer &out = out_k_ ? *d_out2_ : *d_out_;  // (the previous frame may still be leaving the other one)
            filter(d_in_->u8(), out.u8());
            clk_.lap(2);
            if (egress_pending_) finish_egress();
            frame_sink_.wait();
            clk_.lap(3);
            gpu::ck(oat_memcpy_async(ctx_->h, 1, frame_sink_.pixels(), out.p, out_bytes_));
            egress_pending_ = true;
            egress_sample_ = sample;
            out_k_ ^= 1;
        }
        clk_.lap(4);
        ++clk_.n;
        return 0;
    }
    virtual PixelColor outputColor(PixelColor in) const { return in; }
    virtual void setup() {}
    // device in -> device out
    virtual void filter(const uint8_t *d_in, uint8_t *d_out) = 0;

    std::string name_;
    std::string frame_source_address_, frame_sink_address_;
    Source<Frame> frame_source_;
    Sink<Frame> frame_sink_;
    Frame shared_frame_;
    FrameParams in_;
    size_t out_bytes_{0};
    int gpu_index_{0};
    FrameMemory src_memory_{FrameMemory::HOST_SHM};  // what the SOURCE said when this component connected
    bool device_sink_{false};  // --device-sink: publish frames in device memory (SharedFrameHeader memory kind DEVICE)
    std::unique_ptr<gpu::Context> ctx_;
    std::unique_ptr<gpu::HostRegistration> src_pin_, dst_pin_;
    std::unique_ptr<gpu::IpcImport> src_dev_;
    std::unique_ptr<gpu::DeviceBuffer> d_in_, d_out_, d_out2_;

public:
    void set_device_sink(bool v) { device_sink_ = v; }
};

// ---- framefilt mog (BackgroundSubtractorMOG.{h,cpp}) -----------------------------------------------------
class BackgroundSubtractorMOG : public FrameFilter {
public:
    BackgroundSubtractorMOG(const std::string &source, const std::string &sink) : FrameFilter(source, sink)
    {
        name_ = "mogfilt[" + source + "->" + sink + "]";
    }
    ~BackgroundSubtractorMOG() { oat_mog_destroy(mog_); }
    std::vector<config::OptionSpec> options() const override
    {
        return {{"adaptation-coeff", 'a', true,
                 "Value, 0 to 1.0, specifying how quickly the statistical model of the background image should be updated. Default is 0, specifying no adaptation."},
                {"gpu-index", 0, true, "Index of the GPU to use for performing background subtraction."}};
    }
    void applyConfiguration(const config::VariableMap &vm, const config::OptionTable &t) override
    {
        // ...MOG.cpp:87-88: range-checked [0, 1], default 0.0 (frozen model after the first frame)
        config::getNumericValue<double>(vm, t, "adaptation-coeff", learning_coeff_, 0.0, 1.0);
        // ...MOG.cpp:92-111: "Selected GPU index is invalid." comes back from oat_ctx_create
        config::getNumericValue<int>(vm, t, "gpu-index", gpu_index_, 0, 1 << 20);
    }

protected:
    void setup() override
    {
        if (in_.color != PIX_BGR) throw std::runtime_error("framefilt mog requires a BGR frame source.");
        gpu::ck(oat_mog_create(ctx_->h, (int)in_.rows, (int)in_.cols, nullptr /* MOG2 defaults, ...MOG.cpp:83 */, &mog_));
    }
    void filter(const uint8_t *d_in, uint8_t *d_out) override
    {
        const size_t pitch = in_.cols * 3;
        gpu::ck(oat_mog_apply(mog_, d_in, pitch, d_out, pitch, nullptr, 0, learning_coeff_));
    }

private:
    double learning_coeff_{0.0};
    oat_mog *mog_{nullptr};
};

// ---- framefilt col (ColorConvert.{h,cpp}) -----------------------------------------------------------------
class ColorConvert : public FrameFilter {
public:
    ColorConvert(const std::string &source, const std::string &sink) : FrameFilter(source, sink)
    {
        name_ = "colorconvert[" + source + "->" + sink + "]";
    }
    std::vector<config::OptionSpec> options() const override
    {
        return {{"color", 'C', true, "Pixel color of the output frames. Values: GREY, BGR, HSV (only BGR -> HSV runs on the GPU path)."},
                {"gpu-index", 0, true, "Index of the GPU to use."}};
    }
    void applyConfiguration(const config::VariableMap &vm, const config::OptionTable &t) override
    {
        std::string c;
        if (!config::getString(vm, t, "color", c)) throw std::runtime_error("A pixel color must be specified (-C).");
        color_ = str_color(c);
        config::getNumericValue<int>(vm, t, "gpu-index", gpu_index_, 0, 1 << 20);
    }

protected:
    PixelColor outputColor(PixelColor in) const override
    {
        // ColorConvert.cpp:78-86: a no-op conversion is an error
        if (in == color_) throw std::runtime_error("Color conversion is not required: the source is already " + color_str(in) + ".");
        if (!(in == PIX_BGR && color_ == PIX_HSV))
            throw std::runtime_error("Only the BGR -> HSV conversion of the tracking path is implemented on the GPU.");
        return color_;
    }
    void filter(const uint8_t *d_in, uint8_t *d_out) override
    {
        const size_t pitch = in_.cols * 3;
        gpu::ck(oat_bgr2hsv(ctx_->h, d_in, pitch, d_out, pitch, (int)in_.rows, (int)in_.cols));
    }

private:
    PixelColor color_{PIX_HSV};
};

// ---- framefilt bsub (BackgroundSubtractor.{h,cpp}) -----------------------------------------------------------
class BackgroundSubtractor : public FrameFilter {
public:
    BackgroundSubtractor(const std::string &source, const std::string &sink) : FrameFilter(source, sink)
    {
        name_ = "bsub[" + source + "->" + sink + "]";
    }
    ~BackgroundSubtractor() { oat_bsub_destroy(bsub_); }
    std::vector<config::OptionSpec> options() const override
    {
        return {{"adaptation-coeff", 'a', true, "Scalar value, 0 to 1.0, specifying how quickly the new frames are used to update the background image. Default is 0."},
                {"gpu-index", 0, true, "Index of the GPU to use."}};
    }
    void applyConfiguration(const config::VariableMap &vm, const config::OptionTable &t) override
    {
        config::getNumericValue<double>(vm, t, "adaptation-coeff", alpha_, 0.0, 1.0);
        config::getNumericValue<int>(vm, t, "gpu-index", gpu_index_, 0, 1 << 20);
    }

protected:
    void setup() override
    {
        gpu::ck(oat_bsub_create(ctx_->h, (int)in_.rows, (int)in_.cols, color_bytes(in_.color), alpha_, &bsub_));
    }
    void filter(const uint8_t *d_in, uint8_t *d_out) override
    {
        const size_t pitch = in_.cols * (size_t)color_bytes(in_.color);
        gpu::ck(oat_bsub_apply(bsub_, d_in, pitch, d_out, pitch));
    }

private:
    double alpha_{0.0};
    oat_bsub *bsub_{nullptr};
};

// ---- framefilt thresh (Threshold.{h,cpp}) ----------------------------------------------------------------------
class Threshold : public FrameFilter {
public:
    Threshold(const std::string &source, const std::string &sink) : FrameFilter(source, sink) { name_ = "thresh[" + source + "->" + sink + "]"; }
    std::vector<config::OptionSpec> options() const override
    {
        return {{"intensity", 'I', true, "Array of ints between 0 and 256, [min,max], specifying the intensity passband."},
                {"gpu-index", 0, true, "Index of the GPU to use."}};
    }
    void applyConfiguration(const config::VariableMap &vm, const config::OptionTable &t) override
    {
        std::vector<int> v;
        if (config::getArray<int>(vm, t, "intensity", v, 2)) {
            i_min_ = v[0];
            i_max_ = v[1];
            if (i_min_ < 0 || i_min_ > 256 || i_max_ < 0 || i_max_ > 256)
                throw std::runtime_error("Values of intensity should be between 0 and 256.");  // Threshold.cpp:58-62
        }
        config::getNumericValue<int>(vm, t, "gpu-index", gpu_index_, 0, 1 << 20);
    }

protected:
    void filter(const uint8_t *d_in, uint8_t *d_out) override
    {
        const int ch = color_bytes(in_.color);
        gpu::ck(oat_keep_where(ctx_->h, d_in, in_.cols * ch, d_out, in_.cols * ch, (int)in_.rows, (int)in_.cols, ch, nullptr, 0, i_min_, i_max_));
    }

private:
    int i_min_{0}, i_max_{256};
};

// ---- framefilt mask (FrameMasker.{h,cpp}); the mask image is a binary PGM (P5) file --------------------------------
class FrameMasker : public FrameFilter {
public:
    FrameMasker(const std::string &source, const std::string &sink) : FrameFilter(source, sink) { name_ = "framemask[" + source + "->" + sink + "]"; }
    std::vector<config::OptionSpec> options() const override
    {
        return {{"mask", 'm', true, "Path to a binary (P5) PGM image used to mask frames from SOURCE: pixels where the mask is 0 are set to 0."},
                {"gpu-index", 0, true, "Index of the GPU to use."}};
    }
    void applyConfiguration(const config::VariableMap &vm, const config::OptionTable &t) override
    {
        std::string path;
        if (config::getString(vm, t, "mask", path)) {
            FILE *f = std::fopen(path.c_str(), "rb");
            int w = 0, h = 0, maxv = 0;
            if (!f || std::fscanf(f, "P5 %d %d %d", &w, &h, &maxv) != 3 || maxv != 255 || std::fgetc(f) == EOF) {
                if (f) std::fclose(f);
                throw std::runtime_error("File \"" + path + "\" could not be read.");  // FrameMasker.cpp:64-65
            }
            roi_.resize((size_t)w * h);
            const size_t got = std::fread(roi_.data(), 1, roi_.size(), f);
            std::fclose(f);
            if (got != roi_.size()) throw std::runtime_error("File \"" + path + "\" could not be read.");
            roi_w_ = w;
            roi_h_ = h;
        }
        config::getNumericValue<int>(vm, t, "gpu-index", gpu_index_, 0, 1 << 20);
    }

protected:
    void setup() override
    {
        if (!roi_.empty()) {
            if ((size_t)roi_w_ != in_.cols || (size_t)roi_h_ != in_.rows) throw std::runtime_error("Mask image and frames must have the same size.");
            d_roi_.reset(new gpu::DeviceBuffer(*ctx_, roi_.size()));
            gpu::ck(oat_memcpy(ctx_->h, d_roi_->p, roi_.data(), roi_.size()));
        }
    }
    void filter(const uint8_t *d_in, uint8_t *d_out) override
    {
        const int ch = color_bytes(in_.color);
        if (d_roi_)
            gpu::ck(oat_keep_where(ctx_->h, d_in, in_.cols * ch, d_out, in_.cols * ch, (int)in_.rows, (int)in_.cols, ch, d_roi_->u8(), in_.cols, 0, 0));
        else  // no mask configured: frames pass through (FrameMasker.cpp:73)
            gpu::ck(oat_memcpy(ctx_->h, d_out, d_in, in_.bytes));
    }

private:
    std::vector<uint8_t> roi_;
    int roi_w_{0}, roi_h_{0};
    std::unique_ptr<gpu::DeviceBuffer> d_roi_;
};

}  // namespace oat

static void printUsage(std::ostream &out)
{
    out << "Usage: framefilt [INFO]\n"
           "   or: framefilt TYPE SOURCE SINK [CONFIGURATION]\n"
           "Filter frames from SOURCE and published filtered frames to SINK.\n\n"
           "TYPE\n"
           "  bsub: Background subtraction\n"
           "  col: Color conversion (BGR to HSV)\n"
           "  mask: Binary frame masking\n"
           "  thresh: Simple intensity threshold\n"
           "  mog: Mixture of Gaussians background segmentation\n\n"
           "SOURCE:\n  User-supplied name of the memory segment to receive frames from (e.g. raw).\n\n"
           "SINK:\n  User-supplied name of the memory segment to publish frames to (e.g. filt).\n\n"
           "INFO:\n  --help                 Produce help message.\n  -v [ --version ]       Print version information.\n\n"
           "CONFIGURATION:\n  -c [ --config ] FILE KEY   Configuration file/key pair.\n";
}

int main(int argc, char *argv[])
{
    using namespace oat;
    std::string comp_name = "framefilt";
    try {
        for (int i = 1; i < argc; ++i) {
            const std::string a = argv[i];
            if (a == "--help" && argc == 2) { printUsage(std::cout); return 0; }
            if (a == "-v" || a == "--version") { std::cout << "Oat Frame Filter (B200) version 0.1\n"; return 0; }
        }
        if (argc < 2) { printUsage(std::cout); return 0; }
        const std::string type = argv[1];
        // first pass: positional only, everything else left for the TYPE's own options (main.cpp:117-156)
        std::vector<std::string> pos;
        for (int i = 2; i < argc && pos.size() < 2; ++i) {
            if (argv[i][0] == '-') break;
            pos.push_back(argv[i]);
        }
        if (type != "mog" && type != "col" && type != "bsub" && type != "thresh" && type != "mask") {
            printUsage(std::cout);
            std::cerr << whoError(comp_name, "Error: invalid TYPE specified.\n");
            return -1;
        }
        if (pos.size() < 1) { printUsage(std::cout); std::cerr << whoError(comp_name, "Error: a SOURCE must be specified.\n"); return -1; }
        if (pos.size() < 2) { printUsage(std::cout); std::cerr << whoError(comp_name, "Error: a SINK must be specified.\n"); return -1; }
        std::shared_ptr<FrameFilter> filter;
        if (type == "mog") filter = std::make_shared<BackgroundSubtractorMOG>(pos[0], pos[1]);
        else if (type == "col") filter = std::make_shared<ColorConvert>(pos[0], pos[1]);
        else if (type == "thresh") filter = std::make_shared<Threshold>(pos[0], pos[1]);
        else if (type == "mask") filter = std::make_shared<FrameMasker>(pos[0], pos[1]);
        else filter = std::make_shared<BackgroundSubtractor>(pos[0], pos[1]);
        comp_name = filter->name();

        auto opts = filter->options();
        opts.push_back({"device-sink", 0, false, "Publish filtered frames in GPU memory (zero-copy hand-off to GPU components on the same device)."});
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
        filter->set_device_sink(vm.count("device-sink"));

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
