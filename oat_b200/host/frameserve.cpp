// oat-frameserve -- stand-in pure SINK for the reference's `oat frameserve test` (src/frameserver/
// TestFrame.cpp:81-128): serves N frames of the synthetic stream into a frame SINK,
// advancing the Sample clock the way frame servers do.  TYPE `synth` only; camera / codec sources are
// out of scope.
#include <iostream>
#include <thread>

#include <memory>

#include "gpu.h"
#include "oat_cli.h"
#include "oat_host.h"
#include "synth.h"

int main(int argc, char *argv[])
{
    using namespace oat;
    const std::string comp_name = "frameserve";
    try {
        if (argc < 3 || std::string(argv[1]) != "synth") {
            std::cout << "Usage: frameserve synth SINK [--rows R --cols C --num-samples N --fps F --seed S --device --gpu-index I]\n";
            return argc < 2 ? 0 : -1;
        }
        struct Server : Component {
            std::string name() const override { return "synthserve"; }
            bool connectToNode() override { return true; }
            int process() override { return 1; }
        } sig_owner;  // installs the SIGINT handler
        const std::string sink_addr = argv[2];
        const std::vector<config::OptionSpec> opts = {{"rows", 0, true, ""}, {"cols", 0, true, ""}, {"num-samples", 'n', true, ""},
                                                      {"fps", 'r', true, ""}, {"seed", 0, true, ""}, {"device", 0, false, ""},
                                                      {"gpu-index", 0, true, ""}};
        const config::VariableMap vm = config::parse(argc, argv, 3, opts);
        config::OptionTable none;
        int rows = 480, cols = 640, seed = 1000;
        uint64_t n = 100;
        double fps = 0.0;
        config::getNumericValue<int>(vm, none, "rows", rows, 20, 32768);
        config::getNumericValue<int>(vm, none, "cols", cols, 4, 32768);
        config::getNumericValue<uint64_t>(vm, none, "num-samples", n, 0, (uint64_t)1 << 62);
        config::getNumericValue<double>(vm, none, "fps", fps, 0.0, 1e6);
        config::getNumericValue<int>(vm, none, "seed", seed, 0, 1 << 30);

        int gpu_index = 0;
        config::getNumericValue<int>(vm, none, "gpu-index", gpu_index, 0, 1 << 20);
        const bool device = vm.count("device");  // frames are generated on the GPU and published in device memory
        Sink<Frame> frame_sink;
        frame_sink.bind(sink_addr, (size_t)rows * cols * 3);
        Frame shared_frame = frame_sink.retrieve(rows, cols, 3, PIX_BGR);
        if (fps > 0.0) shared_frame.set_rate_hz(fps);
        std::unique_ptr<gpu::Context> ctx;
        std::unique_ptr<gpu::DeviceBuffer> d_frame;
        if (device) {
            ctx.reset(new gpu::Context(gpu_index));
            d_frame.reset(new gpu::DeviceBuffer(*ctx, (size_t)rows * cols * 3));
            unsigned char handle[64];
            gpu::ck(oat_ipc_export(ctx->h, d_frame->p, handle));
            frame_sink.publish_device(handle, gpu_index);
        }
        std::vector<uint8_t> next((size_t)rows * cols * 3);
        auto tick = std::chrono::steady_clock::now();
        for (uint64_t t = 0; t < n && !quit; ++t) {
            if (!device) synth::frame(next.data(), rows, cols, (uint32_t)seed, (uint32_t)t);  // outside the critical section
            frame_sink.wait();
            if (device)
                gpu::ck(oat_synth_frame(ctx->h, d_frame->u8(), (size_t)cols * 3, rows, cols, (uint32_t)seed, (uint32_t)t));
            else
                std::memcpy(shared_frame.data(), next.data(), next.size());
            shared_frame.incrementSampleCount();  // only pure SINKs advance time (TestFrame.cpp:114)
            frame_sink.post();
            if (fps > 0.0) {
                tick += std::chrono::duration_cast<std::chrono::steady_clock::duration>(std::chrono::duration<double>(1.0 / fps));
                std::this_thread::sleep_until(tick);
            }
        }
        // let downstream read the last frame before the sink leaves (its destructor flags END)
        frame_sink.wait();
        return 0;
    } catch (const std::exception &ex) {
        std::cerr << whoError(comp_name, ex.what()) << std::endl;
    }
    return -1;
}
