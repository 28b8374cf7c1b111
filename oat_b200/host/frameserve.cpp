// oat-frameserve -- the pure frame SINKs the hot path is measured with:
//
//   test   `oat frameserve test SINK -f IMAGE [-n N] [-r FPS] [-C COLOR]` -- the reference's TestFrame
//          (src/frameserver/TestFrame.cpp:37-128): ONE static image, published N times with zero copies (only the
//          Sample clock advances).  It is the driver of the reference's own performance protocol
//          (test/perf/framefilt-mog.sh:1-3, test/perf/results.md:12-18: 1000 x 1 MP frames, free-running).
//          The image is a binary PPM/PGM or a NumPy .npy (image_io.h; cv::imread is not available here).
//   file   `oat frameserve file SINK -f CLIP [-r FPS] [--roi "[x0,y0,w,h]"]` -- the reference's FileReader
//          (src/frameserver/FileReader.cpp:36-131): the frames of a clip in order, then end of stream.  The clip is a
//          lossless NumPy .npy (frames x rows x cols [x 3], image_io.h; cv::VideoCapture is not available here).  With
//          --device the whole clip is uploaded to HBM once and every frame is published IN PLACE (the shared frame
//          header's device_offset moves; no per-frame copy at all).
//   synth  the synthetic tracking stream of SURVEY.md 8(d) (deterministic, the same arithmetic as the oracle and
//          the CUDA generator), N frames.
//
// Both take the GPU Frame variant: --device publishes the frames in device memory (SharedFrameHeader memory kind
// DEVICE, CUDA IPC handle in the header) on --gpu-index; a test image is uploaded once, synthetic frames are
// generated on the device.  Cameras, video files and codecs are out of scope.
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <memory>
#include <thread>

#include "gpu.h"
#include "image_io.h"
#include "oat_cli.h"
#include "oat_host.h"
#include "synth.h"

static void printUsage(std::ostream &out)
{
    out << "Usage: frameserve [INFO]\n"
           "   or: frameserve TYPE SINK [CONFIGURATION]\n"
           "Serve frames to SINK.\n\n"
           "TYPE\n"
           "  test: Serve a static test image (binary PPM/PGM or uint8 .npy).\n"
           "  file: Serve the frames of a clip (uint8 .npy, frames x rows x cols [x 3]).\n"
           "  synth: Serve the synthetic single-blob tracking stream.\n\n"
           "SINK:\n  User-supplied name of the memory segment to publish frames to (e.g. raw).\n\n"
           "INFO:\n  --help                 Produce help message.\n  -v [ --version ]       Print version information.\n\n"
           "CONFIGURATION:\n  -c [ --config ] FILE KEY   Configuration file/key pair.\n";
}

int main(int argc, char *argv[])
{
    using namespace oat;
    std::string comp_name = "frameserve";
    try {
        for (int i = 1; i < argc; ++i) {
            const std::string a = argv[i];
            if (a == "--help" && argc == 2) { printUsage(std::cout); return 0; }
            if (a == "-v" || a == "--version") { std::cout << "Oat Frame Server (B200) version 0.1\n"; return 0; }
        }
        if (argc < 2) { printUsage(std::cout); return 0; }
        const std::string type = argv[1];
        if (type != "test" && type != "synth" && type != "file") {
            printUsage(std::cout);
            std::cerr << whoError(comp_name, "Error: invalid TYPE specified.\n");
            return -1;
        }
        if (argc < 3 || argv[2][0] == '-') {
            printUsage(std::cout);
            std::cerr << whoError(comp_name, "Error: a SINK must be specified.\n");
            return -1;
        }
        struct Server : Component {
            std::string name() const override { return "frameserve"; }
            bool connectToNode() override { return true; }
            int process() override { return 1; }
        } sig_owner;  // installs the SIGINT handler
        const std::string sink_addr = argv[2];
        comp_name = (type == "test" ? "testframe[*->" : type == "file" ? "filereader[*->" : "synthserve[*->") + sink_addr + "]";

        std::vector<config::OptionSpec> opts;
        if (type == "test")  // TestFrame::options (TestFrame.cpp:37-55)
            opts = {{"test-image", 'f', true, "Path to test image used as frame source."},
                    {"color", 'C', true, "Pixel color format. Defaults to BGR. Values: GREY, BGR."},
                    {"fps", 'r', true, "Frames to serve per second."},
                    {"num-frames", 'n', true, "Number of frames to serve before exiting."}};
        else if (type == "file")  // FileReader::options (FileReader.cpp:36-54)
            opts = {{"video-file", 'f', true, "Path to the clip to serve frames from (uint8 .npy)."},
                    {"fps", 'r', true, "Frames to serve per second."},
                    {"roi", 0, true, "Four element array of unsigned ints, [x0,y0,width,height], defining a rectangular region of interest."}};
        else
            opts = {{"rows", 0, true, "Frame height."}, {"cols", 0, true, "Frame width."},
                    {"num-samples", 'n', true, "Number of frames to serve before exiting."},
                    {"fps", 'r', true, "Frames to serve per second."}, {"seed", 0, true, "Stream seed."}};
        opts.push_back({"device", 0, false, "Publish frames in device memory (CUDA IPC)."});
        opts.push_back({"gpu-index", 0, true, "Index of the GPU to use."});
        const std::vector<config::OptionSpec> own = opts;
        opts.push_back({"config", 'c', true, "Configuration file/key pair."});
        opts.push_back({"help", 0, false, ""});
        const config::VariableMap vm = config::parse(argc, argv, 3, opts);
        if (vm.count("help")) {
            printUsage(std::cout);
            for (const auto &o : own) std::cout << "  --" << o.long_name << "  " << o.help << "\n";
            return 0;
        }
        config::OptionTable table;
        if (vm.count("config")) {
            table = config::getConfigTable(vm.values.at("config"), vm.values.at("config-key"));
            config::checkKeys(own, table);
        }

        int rows = 480, cols = 640, seed = 1000, channels = 3;
        uint64_t n = type == "test" ? std::numeric_limits<uint64_t>::max() : 100;  // TestFrame.h: serves until interrupted by default
        double fps = 0.0;
        PixelColor color = PIX_BGR;
        Image image;
        Clip clip;
        std::vector<size_t> roi;
        if (type == "file") {
            std::string file;
            if (!config::getString(vm, table, "video-file", file))
                throw std::runtime_error("Required configuration key 'video-file' was not specified.");
            clip.open(file);
            rows = clip.rows;
            cols = clip.cols;
            channels = clip.channels;
            color = channels == 3 ? PIX_BGR : PIX_GREY;
            n = clip.frames;
            if (config::getArray<size_t>(vm, table, "roi", roi, 4)) {
                if (roi[2] == 0 || roi[3] == 0 || roi[0] + roi[2] > (size_t)cols || roi[1] + roi[3] > (size_t)rows)
                    throw std::runtime_error("ROI must fit within the frame size.");
                rows = (int)roi[3];
                cols = (int)roi[2];
            }
        } else if (type == "test") {
            std::string file, col;
            if (!config::getString(vm, table, "test-image", file))
                throw std::runtime_error("Required configuration key 'test-image' was not specified.");  // getValue(..., required = true)
            if (config::getString(vm, table, "color", col)) {
                if (col == "GREY") color = PIX_GREY;
                else if (col == "BGR") color = PIX_BGR;
                else throw std::runtime_error("Invalid color format: " + col);
            }
            config::getNumericValue<uint64_t>(vm, table, "num-frames", n, 1, std::numeric_limits<uint64_t>::max());
            image = to_channels(read_image(file), color == PIX_GREY ? 1 : 3);
            rows = image.rows;
            cols = image.cols;
            channels = image.channels;
        } else {
            config::getNumericValue<int>(vm, table, "rows", rows, 20, 32768);
            config::getNumericValue<int>(vm, table, "cols", cols, 4, 32768);
            config::getNumericValue<uint64_t>(vm, table, "num-samples", n, 0, (uint64_t)1 << 62);
            config::getNumericValue<int>(vm, table, "seed", seed, 0, 1 << 30);
        }
        config::getNumericValue<double>(vm, table, "fps", fps, 0.0, 1e6);
        int gpu_index = 0;
        config::getNumericValue<int>(vm, table, "gpu-index", gpu_index, 0, 1 << 20);
        const bool device = vm.count("device");  // frames live in device memory and are published through a CUDA IPC handle
        const size_t bytes = (size_t)rows * cols * channels;

        Sink<Frame> frame_sink;
        frame_sink.bind(sink_addr, bytes, false);  // announced once parameters and memory kind are final
        Frame shared_frame = frame_sink.retrieve(rows, cols, channels, color);
        if (fps > 0.0) shared_frame.set_rate_hz(fps);
        std::unique_ptr<gpu::Context> ctx;
        std::unique_ptr<gpu::DeviceBuffer> d_frame;
        // file --device: the whole clip lives in HBM and frames are published in place (device_offset)
        const bool clip_in_hbm = device && type == "file" && roi.empty();
        if (device) {
            ctx.reset(new gpu::Context(gpu_index));
            d_frame.reset(new gpu::DeviceBuffer(*ctx, clip_in_hbm ? clip.frames * bytes : bytes));
            if (clip_in_hbm) gpu::ck(oat_memcpy(ctx->h, d_frame->p, clip.data, clip.frames * bytes));
            unsigned char handle[64];
            gpu::ck(oat_ipc_export(ctx->h, d_frame->p, handle));
            frame_sink.publish_device(handle, gpu_index);
            // a static image / a clip held in HBM: every published frame stays where it is, a SOURCE may read it in place
            // after it has posted (SharedFrameHeader::persistent)
            frame_sink.set_persistent(type == "test" || clip_in_hbm);
        }
        if (type == "test") {  // static image, never changes (TestFrame.cpp:93-94)
            if (device)
                gpu::ck(oat_memcpy(ctx->h, d_frame->p, image.data.data(), bytes));
            else
                std::memcpy(shared_frame.data(), image.data.data(), bytes);
        }
        frame_sink.announce();

        std::vector<uint8_t> next((type == "synth" && !device) || !roi.empty() ? bytes : 0);
        auto tick = std::chrono::steady_clock::now();
        const auto serve_t0 = tick;
        uint64_t served = 0;
        for (uint64_t t = 0; t < n && !quit; ++t, ++served) {
            if (type == "synth" && !device) synth::frame(next.data(), rows, cols, (uint32_t)seed, (uint32_t)t);  // outside the critical section
            const uint8_t *src = nullptr;
            if (type == "file") {
                src = clip.data + t * clip.frame_bytes();
                if (!roi.empty()) {  // frame(region_of_interest_) (FileReader.cpp:110-111), packed
                    const size_t rb = (size_t)cols * channels;
                    for (int y = 0; y < rows; ++y)
                        std::memcpy(next.data() + (size_t)y * rb, src + ((roi[1] + y) * (size_t)clip.cols + roi[0]) * channels, rb);
                    src = next.data();
                }
            }
            frame_sink.wait();
            if (type == "file") {
                if (clip_in_hbm)
                    frame_sink.set_device_offset(t * bytes);
                else if (device)
                    gpu::ck(oat_memcpy(ctx->h, d_frame->p, src, bytes));
                else
                    std::memcpy(shared_frame.data(), src, bytes);
            } else if (type == "synth") {
                if (device)
                    gpu::ck(oat_synth_frame(ctx->h, d_frame->u8(), (size_t)cols * 3, rows, cols, (uint32_t)seed, (uint32_t)t));
                else
                    std::memcpy(shared_frame.data(), next.data(), next.size());
            }
            shared_frame.incrementSampleCount();  // only pure SINKs advance time (TestFrame.cpp:114)
            frame_sink.post();
            if (fps > 0.0) {  // without --fps the server free-runs (TestFrame.cpp:122; the perf protocol relies on it)
                tick += std::chrono::duration_cast<std::chrono::steady_clock::duration>(std::chrono::duration<double>(1.0 / fps));
                std::this_thread::sleep_until(tick);
            }
        }
        // let downstream read the last frame before the sink leaves (its destructor flags END)
        frame_sink.wait();
        if (getenv("OAT_B200_TIMING")) {
            // the serving loop alone (process start-up -- for --device the CUDA context, ~0.5 s -- is not in it)
            const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - serve_t0).count();
            const double t0_epoch = std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count() - sec;
            char line[256];
            snprintf(line, sizeof line, "%s: served %llu frames in %.6f s, serving started at %.6f\n", comp_name.c_str(), (unsigned long long)served, sec, t0_epoch);
            std::cerr << line;
        }
        // persistent device frames may still be read in place: the allocation outlives its readers
        if (device && (type == "test" || clip_in_hbm)) frame_sink.end_and_linger(20000);
        return 0;
    } catch (const std::exception &ex) {
        std::cerr << whoError(comp_name, ex.what()) << std::endl;
    }
    return -1;
}
