// oat-buffer -- `oat buffer frame SOURCE SINK`: a FIFO between a frame SOURCE and a frame SINK that absorbs a
// consumer's hiccups so that the producer never waits for it (src/buffer/FrameBuffer.cpp:33-116, Buffer.h:60-80:
// the component's main thread pushes every frame it receives, a second thread pops them to the SINK as its readers
// come back for more; a full FIFO drops the frame and says "Buffer overrun.").
//
// GPU-aware: the FIFO is ONE allocation in HBM (--capacity frames, default 1000 like the reference's BUFFSIZE: 6 GB
// of 1080p frames out of 180 GB).  A frame enters it with a single copy -- DMA from the SOURCE's page-locked shared
// memory, or device-to-device from a DEVICE frame -- and with --device-sink it leaves it with NO copy: the SINK exports
// the FIFO's allocation once (CUDA IPC handle in the shared frame header) and every post() only moves the header's
// device_offset to the next slot.  Without --device-sink the frame is copied back into the SINK's page-locked shm.
// Both edges keep the reference's lock-step protocol and every frame keeps its Sample.
#include <atomic>
#include <condition_variable>
#include <iostream>
#include <memory>
#include <mutex>
#include <thread>

#include "gpu.h"
#include "oat_cli.h"
#include "oat_host.h"

namespace oat {

class FrameBuffer : public Component {
public:
    FrameBuffer(const std::string &source, const std::string &sink)
    : name_("buffer[" + source + "->" + sink + "]"), source_address_(source), sink_address_(sink)
    {
    }
    ~FrameBuffer() override
    {
        sink_running_ = false;
        cv_.notify_all();
        if (sink_thread_.joinable()) sink_thread_.join();
    }
    std::string name() const override { return name_; }
    static std::vector<config::OptionSpec> options()
    {
        return {{"capacity", 'n', true, "Frames the FIFO holds (device memory). Default 1000."},
                {"device-sink", 0, false, "Publish frames in device memory (CUDA IPC), in place: no copy out of the FIFO."},
                {"gpu-index", 0, true, "Index of the GPU to use."}};
    }
    void applyConfiguration(const config::VariableMap &vm, const config::OptionTable &t)
    {
        config::getNumericValue<size_t>(vm, t, "capacity", capacity_, 2, (size_t)1 << 24);
        config::getNumericValue<int>(vm, t, "gpu-index", gpu_index_, 0, 1 << 20);
        device_sink_ = vm.count("device-sink");
    }
    uint64_t overruns() const { return overruns_; }

protected:
    bool connectToNode() override  // FrameBuffer.cpp:33-55
    {
        source_.touch(source_address_);
        push_ctx_.reset(new gpu::Context(gpu_index_));  // one context (stream set) per thread
        pop_ctx_.reset(new gpu::Context(gpu_index_));
        if (source_.connect() != SourceState::CONNECTED) return false;
        in_ = source_.parameters();
        src_memory_ = source_.header()->memory;
        require_same_device(source_.header(), gpu_index_, name());
        if (src_memory_ == FrameMemory::DEVICE)
            src_dev_.reset(new gpu::IpcImport(*push_ctx_, source_.header()->ipc_handle));
        else
            src_pin_.reset(new gpu::HostRegistration(source_.pixels(), in_.bytes));
        fifo_.reset(new gpu::DeviceBuffer(*push_ctx_, capacity_ * in_.bytes));
        samples_.resize(capacity_);
        sink_.bind(sink_address_, in_.bytes, false);
        shared_frame_ = sink_.retrieve(in_.rows, in_.cols, color_bytes(in_.color), in_.color);
        if (device_sink_) {
            unsigned char handle[64];
            gpu::ck(oat_ipc_export(push_ctx_->h, fifo_->p, handle));
            sink_.publish_device(handle, gpu_index_);
        } else {
            dst_pin_.reset(new gpu::HostRegistration(sink_.pixels(), in_.bytes));
            sink_.set_memory(FrameMemory::HOST_PINNED, gpu_index_);
        }
        sink_.announce();
        sink_thread_ = std::thread(&FrameBuffer::pop, this);
        return true;
    }
    int process() override  // FrameBuffer.cpp:57-83
    {
        if (source_.wait() == NodeState::END) {
            // every frame that went in comes out before the SINK leaves
            while (published_.load() != head_.load() && !quit) {
                cv_.notify_one();
                std::this_thread::sleep_for(std::chrono::milliseconds(1));
            }
            return 1;
        }
        if (source_.header()->memory != src_memory_) throw std::runtime_error("SOURCE frame memory kind changed after connect()");
        const uint64_t h = head_.load(std::memory_order_relaxed);
        if (h - tail_.load(std::memory_order_acquire) >= capacity_) {
            std::cerr << "Buffer overrun.\n";  // FrameBuffer.cpp:67-68: the frame is dropped
            ++overruns_;
        } else {
            const size_t slot = (size_t)(h % capacity_);
            const uint8_t *src = src_dev_ ? static_cast<const uint8_t *>(src_dev_->p) + source_.header()->device_offset
                                          : static_cast<const uint8_t *>(source_.pixels());
            gpu::ck(oat_memcpy(push_ctx_->h, fifo_->u8() + slot * in_.bytes, src, in_.bytes));  // the one copy into the FIFO
            samples_[slot] = source_.retrieve()->sample();
            head_.store(h + 1, std::memory_order_release);
        }
        source_.post();
        cv_.notify_one();
        return 0;
    }
    void pop()  // FrameBuffer.cpp:85-116
    {
        uint64_t next = 0;  // next frame to publish; frames < tail_ may be overwritten
        while (sink_running_) {
            {
                std::unique_lock<std::mutex> lk(cv_m_);
                if (next == head_.load(std::memory_order_acquire) && cv_.wait_for(lk, std::chrono::milliseconds(10)) == std::cv_status::timeout)
                    continue;
            }
            while (next != head_.load(std::memory_order_acquire) && sink_running_) {
                const size_t slot = (size_t)(next % capacity_);
                sink_.wait();  // every reader has finished with the frame published before ...
                tail_.store(next, std::memory_order_release);  // ... so its slot (and all older ones) may be refilled
                if (device_sink_)
                    sink_.set_device_offset((uint64_t)slot * in_.bytes);  // published in place
                else
                    gpu::ck(oat_memcpy(pop_ctx_->h, sink_.pixels(), fifo_->u8() + slot * in_.bytes, in_.bytes));
                shared_frame_.sample() = samples_[slot];
                sink_.post();
                ++next;
                published_.store(next, std::memory_order_release);
            }
        }
        // the readers of the last frame are done with it before the FIFO's memory goes away
        if (next > 0) {
            sink_.wait();
            tail_.store(next, std::memory_order_release);
        }
    }

private:
    const std::string name_, source_address_, sink_address_;
    Source<Frame> source_;
    Sink<Frame> sink_;
    Frame shared_frame_;
    FrameParams in_;
    FrameMemory src_memory_{FrameMemory::HOST_SHM};
    size_t capacity_{1000};  // Buffer.h:74 BUFFSIZE
    int gpu_index_{0};
    bool device_sink_{false};
    std::unique_ptr<gpu::Context> push_ctx_, pop_ctx_;
    std::unique_ptr<gpu::HostRegistration> src_pin_, dst_pin_;
    std::unique_ptr<gpu::IpcImport> src_dev_;
    std::unique_ptr<gpu::DeviceBuffer> fifo_;
    std::vector<Sample> samples_;
    std::atomic<uint64_t> head_{0}, tail_{0}, published_{0};  // frames pushed / frames whose slots may be refilled / frames posted
    uint64_t overruns_{0};
    std::atomic<bool> sink_running_{true};
    std::thread sink_thread_;
    std::mutex cv_m_;
    std::condition_variable cv_;
};

}  // namespace oat

static void printUsage(std::ostream &out)
{
    out << "Usage: buffer [INFO]\n"
           "   or: buffer TYPE SOURCE SINK [CONFIGURATION]\n"
           "Place tokens from SOURCE into a FIFO. Publish tokens in FIFO to SINK.\n\n"
           "TYPE\n"
           "  frame: Frame buffer (FIFO in device memory)\n\n"
           "SOURCE:\n  User-supplied name of the memory segment to receive tokens from (e.g. input).\n\n"
           "SINK:\n  User-supplied name of the memory segment to publish tokens to (e.g. output).\n\n"
           "INFO:\n  --help                 Produce help message.\n  -v [ --version ]       Print version information.\n\n"
           "CONFIGURATION:\n  -c [ --config ] FILE KEY   Configuration file/key pair.\n";
}

int main(int argc, char *argv[])
{
    using namespace oat;
    std::string comp_name = "buffer";
    try {
        for (int i = 1; i < argc; ++i) {
            const std::string a = argv[i];
            if (a == "--help" && argc == 2) { printUsage(std::cout); return 0; }
            if (a == "-v" || a == "--version") { std::cout << "Oat Buffer (B200) version 0.1\n"; return 0; }
        }
        if (argc < 2) { printUsage(std::cout); return 0; }
        const std::string type = argv[1];
        std::vector<std::string> pos;
        for (int i = 2; i < argc && pos.size() < 2; ++i) {
            if (argv[i][0] == '-') break;
            pos.push_back(argv[i]);
        }
        if (type != "frame") {  // (pos2D buffers carry ~300-byte tokens: nothing for a GPU to do)
            printUsage(std::cout);
            std::cerr << whoError(comp_name, "Error: invalid TYPE specified.\n");
            return -1;
        }
        if (pos.size() < 1) { printUsage(std::cout); std::cerr << whoError(comp_name, "Error: a SOURCE must be specified.\n"); return -1; }
        if (pos.size() < 2) { printUsage(std::cout); std::cerr << whoError(comp_name, "Error: a SINK must be specified.\n"); return -1; }
        FrameBuffer buffer(pos[0], pos[1]);
        comp_name = buffer.name();
        auto opts = FrameBuffer::options();
        opts.push_back({"config", 'c', true, "Configuration file/key pair."});
        opts.push_back({"help", 0, false, ""});
        const config::VariableMap vm = config::parse(argc, argv, 4, opts);
        if (vm.count("help")) {
            printUsage(std::cout);
            for (const auto &o : FrameBuffer::options()) std::cout << "  --" << o.long_name << "  " << o.help << "\n";
            return 0;
        }
        config::OptionTable table;
        if (vm.count("config")) {
            table = config::getConfigTable(vm.values.at("config"), vm.values.at("config-key"));
            config::checkKeys(FrameBuffer::options(), table);
        }
        buffer.applyConfiguration(vm, table);
        std::cout << whoMessage(comp_name, "Listening to source " + pos[0] + ".\n")
                  << whoMessage(comp_name, "Steaming to sink " + pos[1] + ".\n")
                  << whoMessage(comp_name, "Press CTRL+C to exit.\n");
        buffer.run();
        std::cout << whoMessage(comp_name, "Exiting.\n");
        return 0;
    } catch (const std::exception &ex) {
        std::cerr << whoError(comp_name, ex.what()) << std::endl;
    } catch (...) {
        std::cerr << whoError(comp_name, "Unknown exception.") << std::endl;
    }
    return -1;
}
