// oat_host.h -- dependency-free C++17 re-statement of the part of Oat's host side that the tracking
// hot path lives behind: the tokens (lib/datatypes: Sample, PixelColor, Frame, Position2D), the
// shared-memory SOURCE -> SINK token transport (lib/shmemdf: Node, Sink<T>, Source<T>,
// SharedFrameHeader), and the Component run loop (lib/base/Component.{h,cpp}).
//
// The reference builds these on Boost.Interprocess + OpenCV (absent from this image: SURVEY.md
// 8(c)); here they sit directly on POSIX shm_open/mmap + process-shared sem_t, with the same names,
// call sequences, blocking behaviour and error messages, so the reference's own transport tests
// (test/shmemdf/*_test.cpp) re-express one to one (oat_b200/host/shmemdf_test.cpp).  The segment
// layout is NOT binary compatible with Boost's managed_shared_memory: every process of a dataflow
// must link this layer (frameserve/posisock stand-ins are provided for that reason).
//
// New relative to the reference (BASELINE.json north star): SharedFrameHeader carries a memory-kind
// tag so that a frame can be plain shared memory (HOST_SHM), shared memory page-locked by each
// process for asynchronous DMA (HOST_PINNED), or a device allocation exported through a CUDA IPC
// handle (DEVICE).
#pragma once
#include <fcntl.h>
#include <sched.h>
#include <semaphore.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <initializer_list>
#include <stdexcept>
#include <string>
#include <typeinfo>

namespace oat {

// ---- lib/base/Globals.h ------------------------------------------------------------------------
extern volatile sig_atomic_t quit;

// ---- lib/datatypes/Sample.h:35-122 ---------------------------------------------------------------
class Sample {
public:
    Sample() = default;
    explicit Sample(double period_sec) { set_rate_hz(1.0 / period_sec); }
    // Only pure SINKs advance time (Sample.h:78-81)
    uint64_t incrementCount()
    {
        microseconds_ += period_microseconds_;
        return ++count_;
    }
    uint64_t incrementCount(uint64_t usec)
    {
        microseconds_ = usec;
        return ++count_;
    }
    void set_rate_hz(double value)
    {
        rate_hz_ = value;
        period_sec_ = 1.0 / value;
        period_microseconds_ = (uint64_t)(period_sec_ * 1e6);  // duration_cast truncates
    }
    uint64_t count() const { return count_; }
    uint64_t microseconds() const { return microseconds_; }
    double period_sec() const { return period_sec_; }
    uint64_t period_microseconds() const { return period_microseconds_; }
    double rate_hz() const { return rate_hz_; }

private:
    uint64_t count_{0};
    uint64_t microseconds_{0};
    double period_sec_{0.0};
    uint64_t period_microseconds_{0};
    double rate_hz_{0.0};
};

// ---- lib/datatypes/Color.h:29-51 -------------------------------------------------------------------
enum PixelColor { PIX_BINARY = 0, PIX_GREY, PIX_BGR, PIX_HSV, PIX_ANY };
inline int color_bytes(PixelColor c) { return (c == PIX_BGR || c == PIX_HSV) ? 3 : 1; }
inline std::string color_str(PixelColor c)
{
    switch (c) {
    case PIX_BINARY: return "BINARY";
    case PIX_GREY: return "GREY";
    case PIX_BGR: return "BGR";
    case PIX_HSV: return "HSV";
    default: return "ANY";
    }
}
inline PixelColor str_color(const std::string &s)
{
    if (s == "BINARY") return PIX_BINARY;
    if (s == "GREY") return PIX_GREY;
    if (s == "BGR") return PIX_BGR;
    if (s == "HSV") return PIX_HSV;
    throw std::runtime_error("Unknown pixel color '" + s + "'.");
}

// ---- lib/datatypes/Frame.h:41-146 (a cv::Mat header + Sample* + PixelColor; here a plain view) ----
// Frames are always 8 bit, 1 or 3 interleaved channels, contiguous.
class Frame {
public:
    Frame() = default;
    // view over caller-owned memory (shared memory in Sink/Source)
    Frame(int rows, int cols, PixelColor color, void *data, void *sample)
        : rows_(rows), cols_(cols), color_(color), data_((uint8_t *)data), sample_ptr_((Sample *)sample)
    {
    }
    ~Frame() { release(); }
    Frame(const Frame &) = delete;
    Frame &operator=(const Frame &) = delete;
    Frame(Frame &&o) noexcept { *this = std::move(o); }
    Frame &operator=(Frame &&o) noexcept
    {
        if (this != &o) {
            release();
            rows_ = o.rows_; cols_ = o.cols_; color_ = o.color_; data_ = o.data_; sample_ptr_ = o.sample_ptr_;
            owned_ = o.owned_; own_sample_ = o.own_sample_;
            if (sample_ptr_ == &o.own_sample_) sample_ptr_ = &own_sample_;
            o.data_ = nullptr; o.owned_ = false; o.sample_ptr_ = nullptr;
        }
        return *this;
    }
    // private heap frame (what Component::process() works on)
    void create(int rows, int cols, PixelColor color)
    {
        if (owned_ && rows == rows_ && cols == cols_ && color_bytes(color) == color_bytes(color_)) {
            color_ = color;
            return;
        }
        release();
        rows_ = rows; cols_ = cols; color_ = color;
        data_ = new uint8_t[bytes()];
        owned_ = true;
        sample_ptr_ = &own_sample_;
    }
    // Frame::copyTo copies pixels AND the sample, by value (Frame.h:113-118)
    void copyTo(Frame &dst) const
    {
        if (!dst.data_ || dst.rows_ != rows_ || dst.cols_ != cols_ || color_bytes(dst.color_) != color_bytes(color_))
            dst.create(rows_, cols_, color_);
        dst.color_ = color_;
        std::memcpy(dst.data_, data_, bytes());
        *dst.sample_ptr_ = *sample_ptr_;
    }
    int rows() const { return rows_; }
    int cols() const { return cols_; }
    int channels() const { return color_bytes(color_); }
    size_t bytes() const { return (size_t)rows_ * cols_ * color_bytes(color_); }
    size_t pitch() const { return (size_t)cols_ * color_bytes(color_); }
    uint8_t *data() { return data_; }
    const uint8_t *data() const { return data_; }
    PixelColor color() const { return color_; }
    void set_color(PixelColor c) { color_ = c; }
    Sample &sample() { return *sample_ptr_; }
    const Sample &sample() const { return *sample_ptr_; }
    void set_rate_hz(double v) { sample_ptr_->set_rate_hz(v); }
    uint64_t sample_count() const { return sample_ptr_->count(); }
    void incrementSampleCount() { sample_ptr_->incrementCount(); }

private:
    void release()
    {
        if (owned_) delete[] data_;
        data_ = nullptr;
        owned_ = false;
    }
    int rows_{0}, cols_{0};
    PixelColor color_{PIX_BGR};
    uint8_t *data_{nullptr};
    Sample *sample_ptr_{nullptr};
    bool owned_{false};
    Sample own_sample_;
};

// ---- lib/datatypes/Position2D.h:60-155 --------------------------------------------------------------
struct Point2D { double x{0.0}, y{0.0}; };
enum class DistanceUnit { PIXELS = 0, WORLD = 1 };

class Position2D {
public:
    explicit Position2D(const std::string &label)
    {
        std::strncpy(label_, label.c_str(), sizeof(label_));
        label_[sizeof(label_) - 1] = '\0';
    }
    // Copy all but label, which is specific to the component (Position2D.h:84-105)
    Position2D &operator=(const Position2D &p)
    {
        if (this == &p) return *this;
        unit_of_length_ = p.unit_of_length_;
        sample_ = p.sample_;
        position_valid = p.position_valid;
        velocity_valid = p.velocity_valid;
        heading_valid = p.heading_valid;
        position = p.position;
        velocity = p.velocity;
        heading = p.heading;
        region_valid = p.region_valid;
        std::strncpy(region, p.region, sizeof(region));
        region[sizeof(region) - 1] = '\0';
        return *this;
    }
    Position2D(const Position2D &) = default;
    char *label() { return label_; }
    DistanceUnit unit_of_length() const { return unit_of_length_; }
    static constexpr size_t REGION_LEN{10};
    bool region_valid{false};
    char region[REGION_LEN]{0};
    bool position_valid{false};
    bool velocity_valid{false};
    bool heading_valid{false};
    Point2D position, velocity, heading;
    void set_sample(const Sample &v) { sample_ = v; }
    uint64_t sample_count() const { return sample_.count(); }
    uint64_t sample_usec() const { return sample_.microseconds(); }
    double sample_period_sec() const { return sample_.period_sec(); }
    static constexpr size_t NPY_DTYPE_BYTES{82};

private:
    char label_[100]{0};
    DistanceUnit unit_of_length_{DistanceUnit::PIXELS};
    Sample sample_;
    double homography_[9]{1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0};
};

// rapidjson Writer::Double with SetMaxDecimalPlaces(5): shortest form, at most 5 decimals, at least one
inline std::string json_double(double v)
{
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%.5f", v);
    std::string s(buf);
    while (s.size() > 1 && s.back() == '0' && s[s.size() - 2] != '.') s.pop_back();
    return s;
}
// serializePosition (Position2D.h:169-233): same keys, order and conditional fields
inline std::string serializePosition(const Position2D &p, bool verbose = false)
{
    std::string s = "{\"tick\":" + std::to_string(p.sample_count()) + ",\"usec\":" + std::to_string(p.sample_usec()) +
                    ",\"unit\":" + std::to_string((int)p.unit_of_length());
    auto xy = [](const Point2D &q) { return "[" + json_double(q.x) + "," + json_double(q.y) + "]"; };
    s += std::string(",\"pos_ok\":") + ((p.position_valid || verbose) ? "true" : "false");
    if (p.position_valid || verbose) s += ",\"pos_xy\":" + xy(p.position);
    s += std::string(",\"vel_ok\":") + ((p.velocity_valid || verbose) ? "true" : "false");
    if (p.velocity_valid || verbose) s += ",\"vel_xy\":" + xy(p.velocity);
    s += std::string(",\"head_ok\":") + (p.heading_valid ? "true" : "false");
    if (p.heading_valid || verbose) s += ",\"head_xy\":" + xy(p.heading);
    s += std::string(",\"reg_ok\":") + (p.region_valid ? "true" : "false");
    if (p.region_valid || verbose) s += std::string(",\"reg\":\"") + p.region + "\"";
    return s + "}";
}
// packPosition (Position2D.cpp:37-96): the 82-byte record of the .npy writer
inline void packPosition(const Position2D &p, char out[Position2D::NPY_DTYPE_BYTES])
{
    char *o = out;
    auto put = [&o](const void *v, size_t n) { std::memcpy(o, v, n); o += n; };
    uint64_t sc = p.sample_count(), su = p.sample_usec();
    int u = (int)p.unit_of_length();
    put(&sc, 8); put(&su, 8); put(&u, 4);
    char ok = p.position_valid ? 1 : 0; put(&ok, 1); put(&p.position.x, 8); put(&p.position.y, 8);
    ok = p.velocity_valid ? 1 : 0; put(&ok, 1); put(&p.velocity.x, 8); put(&p.velocity.y, 8);
    ok = p.heading_valid ? 1 : 0; put(&ok, 1); put(&p.heading.x, 8); put(&p.heading.y, 8);
    ok = p.region_valid ? 1 : 0; put(&ok, 1); put(p.region, Position2D::REGION_LEN);
}

// ---- lib/shmemdf/Node.h:34-183 -------------------------------------------------------------------------
enum class NodeState : int { END = -1, UNDEFINED = 0, SINK_BOUND = 1, ERROR = 2 };

namespace detail {
// A token hand-off between two running components takes a few hundred nanoseconds when the waiter is still looking;
// once it has gone to sleep it costs a futex wake-up plus the core's way out of its idle state -- measured on the B200
// hosts at up to a millisecond, i.e. a graph whose frame period is just above the spin budget runs five times slower
// than one just below it.  So every barrier wait spins first (OAT_B200_SPIN_US, default 2000 us: a core stays busy
// while tokens flow at more than ~500 per second and sleeps between the frames of a camera; 0 = sleep at once, as
// the reference's interprocess semaphores do) and only then blocks.
inline int spin_budget_us()
{
    static const int v = [] {
        const char *e = getenv("OAT_B200_SPIN_US");
        return e ? atoi(e) : 2000;
    }();
    return v;
}
inline void cpu_relax()
{
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    __asm__ __volatile__("" ::: "memory");
#endif
}
inline bool sem_spin(sem_t *s)
{
    if (sem_trywait(s) == 0) return true;
    const int budget = spin_budget_us();
    if (budget <= 0) return false;
    const auto t0 = std::chrono::steady_clock::now();
    const auto until = t0 + std::chrono::microseconds(budget), polite = t0 + std::chrono::microseconds(20);
    for (;;) {
        for (int i = 0; i < 32; ++i) {
            if (sem_trywait(s) == 0) return true;
            cpu_relax();
        }
        const auto now = std::chrono::steady_clock::now();
        if (now >= until) return false;
        // the process we are waiting for may be runnable on THIS core: after 20 us give it the chance every round
        // (returns at once when nothing else is runnable here)
        if (now >= polite) sched_yield();
    }
}
inline bool sem_timedwait_ms(sem_t *s, int ms)
{
    if (sem_spin(s)) return true;
    timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    ts.tv_nsec += (long)ms * 1000000L;
    if (ts.tv_nsec >= 1000000000L) { ts.tv_sec += 1; ts.tv_nsec -= 1000000000L; }
    for (;;) {
        if (sem_timedwait(s, &ts) == 0) return true;
        if (errno == EINTR) { if (quit) return false; continue; }
        return false;  // ETIMEDOUT
    }
}
inline void sem_wait_forever(sem_t *s) { while (sem_wait(s) != 0 && errno == EINTR) {} }
}  // namespace detail

class Node {
public:
    static constexpr size_t NUM_SLOTS{10};
    static constexpr uint32_t MAGIC = 0x4f41544eu;  // "OATN"

    void init()
    {
        sem_init(&write_barrier, 1, 1);  // write always occurs before read (Node.h:142)
        sem_init(&mutex_, 1, 1);
        for (auto &rb : rb_) sem_init(&rb, 1, 0);
        source_slots_ = 0; source_read_required_ = 0; source_ref_count_ = 0; write_number_ = 0;
        sink_state_.store((int)NodeState::UNDEFINED);
        magic_.store(MAGIC, std::memory_order_release);
    }
    bool ready() const { return magic_.load(std::memory_order_acquire) == MAGIC; }

    void set_sink_state(NodeState v) { sink_state_.store((int)v); }
    NodeState sink_state() const { return (NodeState)sink_state_.load(); }
    uint64_t write_number() const { return write_number_; }

    void notifySinkWriteComplete()  // Node.h:69-84
    {
        detail::sem_wait_forever(&mutex_);
        source_read_required_ = source_slots_;
        for (size_t i = 0; i < NUM_SLOTS; i++)
            if (source_slots_ & (1u << i)) sem_post(&rb_[i]);
        ++write_number_;
        sem_post(&mutex_);
    }
    bool notifySourceReadComplete(size_t index)  // Node.h:87-97
    {
        detail::sem_wait_forever(&mutex_);
        source_read_required_ &= ~(1u << index);
        const bool reads_finished = (source_read_required_ == 0);
        sem_post(&mutex_);
        return reads_finished;
    }
    int acquireSlot(size_t &index)  // Node.h:102-121
    {
        detail::sem_wait_forever(&mutex_);
        if (source_slots_ == (1u << NUM_SLOTS) - 1) { sem_post(&mutex_); return -1; }
        index = 0;
        while (source_slots_ & (1u << index)) ++index;
        source_slots_ |= (1u << index);
        source_ref_count_ = (size_t)__builtin_popcount(source_slots_);
        sem_post(&mutex_);
        return 0;
    }
    int releaseSlot(size_t index)
    {
        if (index >= NUM_SLOTS) return -1;
        detail::sem_wait_forever(&mutex_);
        source_slots_ &= ~(1u << index);
        source_ref_count_ = (size_t)__builtin_popcount(source_slots_);
        sem_post(&mutex_);
        return 0;
    }
    size_t source_ref_count() const { return source_ref_count_; }

    sem_t write_barrier;
    // All ten slots are usable here (the reference's switch has no `case 5`, Node.h:154-167).
    sem_t &read_barrier(size_t index)
    {
        if (index >= NUM_SLOTS) throw std::runtime_error("Source index out of range.");
        if (!(source_slots_ & (1u << index)))
            throw std::runtime_error("Requested index refers to a SOURCE that is not bound to this node.");
        return rb_[index];
    }

private:
    std::atomic<uint32_t> magic_;
    std::atomic<int> sink_state_;
    uint32_t source_slots_, source_read_required_;
    size_t source_ref_count_;
    uint64_t write_number_;
    sem_t mutex_;
    sem_t rb_[NUM_SLOTS];
};

// ---- a named POSIX shared-memory segment ("<addr>_node" / "<addr>_obj", Sink.h:244-267) ------------------
class Shmem {
public:
    Shmem() = default;
    ~Shmem() { close(); }
    Shmem(const Shmem &) = delete;
    Shmem &operator=(const Shmem &) = delete;
    enum Mode { OPEN_OR_CREATE, CREATE_ONLY, OPEN_ONLY };
    // returns true if this call created the segment
    bool open(const std::string &name, size_t bytes, Mode mode)
    {
        name_ = "/" + name;
        bool created = false;
        int fd = -1;
        if (mode != OPEN_ONLY) {
            fd = shm_open(name_.c_str(), O_CREAT | O_EXCL | O_RDWR, 0666);
            if (fd >= 0) {
                created = true;
                if (ftruncate(fd, (off_t)bytes) != 0) {
                    ::close(fd);
                    shm_unlink(name_.c_str());
                    throw std::runtime_error("ftruncate failed for shared memory '" + name + "'");
                }
            } else if (mode == CREATE_ONLY || errno != EEXIST) {
                throw std::runtime_error("Could not create shared memory '" + name + "': " + std::strerror(errno));
            }
        }
        if (fd < 0) {
            fd = shm_open(name_.c_str(), O_RDWR, 0666);
            if (fd < 0) throw std::runtime_error("Could not open shared memory '" + name + "': " + std::strerror(errno));
            struct stat st;
            // the creator may not have sized the segment yet
            for (int i = 0; i < 2000; ++i) {
                if (fstat(fd, &st) == 0 && st.st_size > 0) break;
                usleep(1000);
            }
            bytes = (size_t)st.st_size;
        }
        void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        ::close(fd);
        if (p == MAP_FAILED) throw std::runtime_error("mmap failed for shared memory '" + name + "'");
        base_ = (uint8_t *)p;
        bytes_ = bytes;
        return created;
    }
    void close()
    {
        if (base_) munmap(base_, bytes_);
        base_ = nullptr;
    }
    static bool remove(const std::string &name) { return shm_unlink(("/" + name).c_str()) == 0; }
    uint8_t *base() const { return base_; }
    size_t bytes() const { return bytes_; }

private:
    std::string name_;
    uint8_t *base_{nullptr};
    size_t bytes_{0};
};

inline Node *attach_node(Shmem &shm, const std::string &node_address)
{
    const bool created = shm.open(node_address, sizeof(Node), Shmem::OPEN_OR_CREATE);
    Node *n = reinterpret_cast<Node *>(shm.base());
    if (created) {
        n->init();
    } else {
        for (int i = 0; i < 2000 && !n->ready(); ++i) usleep(1000);
        if (!n->ready()) throw std::runtime_error("Node '" + node_address + "' was never initialised (stale segment? run oat-clean).");
    }
    return n;
}

// header of every "<addr>_obj" segment: what bip::managed_shared_memory's name index did
struct ObjectHeader {
    static constexpr uint32_t MAGIC = 0x4f41544fu;  // "OATO"
    uint32_t magic;
    uint32_t type_hash;  // of typeid(T).name(): Source<T> may only connect to Sink<T> (Source.h:142-161)
    uint64_t bytes;
};
inline uint32_t type_hash_of(const char *name)
{
    uint32_t h = 2166136261u;
    for (; *name; ++name) h = (h ^ (uint8_t)*name) * 16777619u;
    return h;
}
constexpr size_t OBJ_OFFSET = 64;

// ---- lib/shmemdf/SharedFrameHeader.h:31-97 (+ the memory-kind tag) -------------------------------------
enum class FrameMemory : int { HOST_SHM = 0, HOST_PINNED = 1, DEVICE = 2 };
struct FrameParams {
    size_t cols{0}, rows{0};
    int channels{0};
    PixelColor color{PIX_BGR};
    size_t bytes{0};
};
struct SharedFrameHeader {
    FrameParams params;
    uint64_t data_offset;    // pixels, relative to the segment base
    uint64_t sample_offset;  // oat::Sample
    FrameMemory memory;
    int device_index;
    unsigned char ipc_handle[64];  // cudaIpcMemHandle_t of the pixels when memory == DEVICE
    uint64_t device_offset{0};     // memory == DEVICE: where THIS frame's pixels start inside the exported allocation
                                   // (a SINK that keeps many frames in HBM -- a preloaded clip, a device FIFO -- exports ONE
                                   // allocation and moves this offset between post()s; 0 for single-buffer SINKs)
    uint32_t persistent{0};        // memory == DEVICE: the pixels of EVERY published frame stay valid and unchanged until the
                                   // SINK leaves (a static test image, a clip preloaded into HBM): a SOURCE may post() as soon
                                   // as it has read offset and Sample, and read the pixels in place later
};

// a DEVICE frame can only be taken over by a component running on the GPU that holds it (CUDA IPC maps the
// exporter's allocation on the same device); say so instead of failing inside cudaIpcOpenMemHandle
inline void require_same_device(const SharedFrameHeader *h, int gpu_index, const std::string &who)
{
    if (h->memory == FrameMemory::DEVICE && h->device_index != gpu_index)
        throw std::runtime_error(who + ": the SOURCE publishes frames in the memory of GPU " + std::to_string(h->device_index) +
                                 ", this component runs on GPU " + std::to_string(gpu_index) + " (use --gpu-index " +
                                 std::to_string(h->device_index) + ").");
}

// ---- lib/shmemdf/Sink.h -------------------------------------------------------------------------------
template <typename T>
class SinkBase {
public:
    SinkBase() = default;
    virtual ~SinkBase()
    {
        if (bound_) {  // Sink.h:73-91
            node_->set_sink_state(NodeState::END);
            if (node_->source_ref_count() == 0) {
                Shmem::remove(node_address_);
                Shmem::remove(obj_address_);
            }
        }
    }
    SinkBase(const SinkBase &) = delete;
    SinkBase &operator=(const SinkBase &) = delete;

    void wait()  // Sink.h:93-116
    {
        if (!bound_) throw std::runtime_error("Sink must be bound before calling wait()");
        if (did_wait_need_post_) throw std::runtime_error("wait() called when post() was required.");
        // Only wait if there is a SOURCE attached to the node; 10 ms polls so quit is observed
        while (node_->source_ref_count() > 0 && !detail::sem_timedwait_ms(&node_->write_barrier, 10) && !quit) {}
        did_wait_need_post_ = true;
    }
    void post()  // Sink.h:118-138
    {
        if (!bound_) throw std::runtime_error("Source must be bound before calling post()");
        if (!did_wait_need_post_) throw std::runtime_error("post() called when wait() was required.");
        node_->notifySinkWriteComplete();
        did_wait_need_post_ = false;
    }
    // Flag END now and stay until every SOURCE has detached (or `ms` have passed): for SINKs whose published frames
    // are read in place after post() (SharedFrameHeader::persistent) -- the memory must outlive the readers.
    void end_and_linger(int ms)
    {
        if (!bound_) return;
        node_->set_sink_state(NodeState::END);
        for (int i = 0; i < ms && node_->source_ref_count() > 0; ++i) usleep(1000);
    }

protected:
    void bind_segments(const std::string &address, size_t obj_bytes, uint32_t type_hash)
    {
        if (bound_) throw std::runtime_error("A sink can only bind a single time to a single node.");
        address_ = address;
        node_address_ = address + "_node";
        obj_address_ = address + "_obj";
        node_ = attach_node(node_shmem_, node_address_);
        if (node_->sink_state() != NodeState::UNDEFINED)
            throw std::runtime_error("Requested SINK address, '" + address + "', is not available.");
        obj_shmem_.open(obj_address_, OBJ_OFFSET + obj_bytes, Shmem::CREATE_ONLY);
        ObjectHeader *h = reinterpret_cast<ObjectHeader *>(obj_shmem_.base());
        h->type_hash = type_hash;
        h->bytes = obj_bytes;
        h->magic = ObjectHeader::MAGIC;
    }
    std::string address_, node_address_, obj_address_;
    Shmem node_shmem_, obj_shmem_;
    Node *node_{nullptr};
    T *sh_object_{nullptr};
    bool bound_{false};

private:
    bool did_wait_need_post_{false};
};

template <typename T>
class Sink : public SinkBase<T> {
public:
    template <typename... Targs>
    void bind(const std::string &address, Targs... args)  // Sink.h:164-207
    {
        this->bind_segments(address, sizeof(T), type_hash_of(typeid(T).name()));
        this->sh_object_ = new (this->obj_shmem_.base() + OBJ_OFFSET) T(args...);
        this->node_->set_sink_state(NodeState::SINK_BOUND);
        this->bound_ = true;
    }
    T *retrieve()
    {
        if (!this->bound_) throw std::runtime_error("SINK must be bound before shared object is retrieved.");
        return this->sh_object_;
    }
};

template <>
class Sink<Frame> : public SinkBase<SharedFrameHeader> {
public:
    // announce_now = false: the node stays un-announced (SOURCEs keep waiting in connect()) until announce() -- for
    // components that must first fill in what a SOURCE reads ONCE when it connects: the frame parameters
    // (retrieve()) and, new with the GPU Frame variants, the memory kind / CUDA IPC handle / device index, which
    // only exist after the CUDA context and the device buffers do (hundreds of ms).
    void bind(const std::string &address, size_t bytes, bool announce_now = true)  // Sink.h:232-272
    {
        bind_segments(address, sizeof(SharedFrameHeader) + bytes + sizeof(Sample) + 64, type_hash_of(typeid(SharedFrameHeader).name()));
        sh_object_ = new (obj_shmem_.base() + OBJ_OFFSET) SharedFrameHeader();
        sh_object_->memory = FrameMemory::HOST_SHM;
        sh_object_->device_index = -1;
        pixel_bytes_ = bytes;
        bound_ = true;
        if (announce_now) announce();
    }
    void announce()
    {
        if (!bound_) throw std::runtime_error("SINK must be bound before it is announced.");
        node_->set_sink_state(NodeState::SINK_BOUND);
    }
    Frame retrieve(size_t rows, size_t cols, int channels, PixelColor color)  // Sink.h:274-298
    {
        if (!bound_) throw std::runtime_error("SINK must be bound before shared frame is retrieved.");
        if (rows * cols * (size_t)channels > pixel_bytes_) throw std::runtime_error("Shared frame does not fit the bound segment.");
        uint8_t *base = obj_shmem_.base();
        const uint64_t sample_off = OBJ_OFFSET + sizeof(SharedFrameHeader);
        const uint64_t data_off = (sample_off + sizeof(Sample) + 63) & ~(uint64_t)63;
        new (base + sample_off) Sample();
        sh_object_->params.rows = rows;
        sh_object_->params.cols = cols;
        sh_object_->params.channels = channels;
        sh_object_->params.color = color;
        sh_object_->params.bytes = rows * cols * (size_t)channels;
        sh_object_->sample_offset = sample_off;
        sh_object_->data_offset = data_off;
        return Frame((int)rows, (int)cols, color, base + data_off, base + sample_off);
    }
    // the pixel region of the segment (for page-locking: HOST_PINNED variant)
    void *pixels() const { return obj_shmem_.base() + sh_object_->data_offset; }
    void set_memory(FrameMemory m, int device = -1) { sh_object_->memory = m; sh_object_->device_index = device; }
    // DEVICE variant: the pixels live in a device allocation exported through a CUDA IPC handle; the shm
    // pixel area stays unused.  Call before the first post().
    void publish_device(const unsigned char handle[64], int device)
    {
        std::memcpy(sh_object_->ipc_handle, handle, 64);
        sh_object_->device_index = device;
        sh_object_->memory = FrameMemory::DEVICE;
    }
    // the frame published by the NEXT post() lives at this offset of the exported allocation (call between wait() and post())
    void set_device_offset(uint64_t off) { sh_object_->device_offset = off; }
    // every frame this SINK publishes stays where it is, unchanged, until the SINK leaves (call before announce())
    void set_persistent(bool v) { sh_object_->persistent = v ? 1u : 0u; }
    SharedFrameHeader *header() { return sh_object_; }

private:
    size_t pixel_bytes_{0};
};

// ---- lib/shmemdf/Source.h ---------------------------------------------------------------------------------
enum class SourceState : int { ERR_CONNECT = -3, ERR_NODEFULL = -2, ERR_TYPEMIS = -1, VIRGIN = 0, TOUCHED = 1, CONNECTED = 2 };

template <typename T>
class SourceBase {
public:
    SourceBase() = default;
    virtual ~SourceBase()  // Source.h:91-112
    {
        if (node_ && (state_ >= SourceState::TOUCHED || state_ == SourceState::ERR_TYPEMIS)) node_->releaseSlot(slot_index_);
        if (node_ != nullptr && node_->source_ref_count() == 0 && node_->sink_state() != NodeState::SINK_BOUND) {
            Shmem::remove(node_address_);
            Shmem::remove(obj_address_);
        }
    }
    SourceBase(const SourceBase &) = delete;
    SourceBase &operator=(const SourceBase &) = delete;

    void touch(const std::string &address)  // Source.h:114-147
    {
        if (state_ != SourceState::VIRGIN) throw std::runtime_error("A source can only connect a single time to a single node.");
        address_ = address;
        node_address_ = address + "_node";
        obj_address_ = address + "_obj";
        node_ = attach_node(node_shmem_, node_address_);
        if (node_->acquireSlot(slot_index_) < 0) {
            state_ = SourceState::ERR_NODEFULL;  // no throw here: connect() does (Source.h:139-143, :153-156)
            return;
        }
        state_ = SourceState::TOUCHED;
    }
    NodeState wait()  // Source.h:187-215
    {
        if (state_ < SourceState::TOUCHED) throw std::runtime_error("Source must have touched node before calling wait()");
        if (did_wait_need_post_) throw std::runtime_error("wait() called when post() was required.");
        while (!detail::sem_timedwait_ms(&node_->read_barrier(slot_index_), 10) && !quit) {
            if (node_->sink_state() == NodeState::END) break;  // if the sink has left the room, we should too
        }
        did_wait_need_post_ = true;
        return node_->sink_state();
    }
    // wait() that does not block: true = a new token was taken (post() is required), false = none is there yet.
    // *sink_state is what wait() would have returned (END once the SINK has left and nothing is left to read).
    // For components that keep work in flight behind the SOURCE and have other things to do meanwhile.
    bool try_wait(NodeState *sink_state)
    {
        if (state_ < SourceState::CONNECTED) throw std::runtime_error("Source must be connected before calling try_wait()");
        if (did_wait_need_post_) throw std::runtime_error("wait() called when post() was required.");
        const bool got = sem_trywait(&node_->read_barrier(slot_index_)) == 0;
        *sink_state = node_->sink_state();
        if (got) did_wait_need_post_ = true;
        return got;
    }
    void post()  // Source.h:217-232
    {
        if (state_ < SourceState::CONNECTED) throw std::runtime_error("source must be connected before calling post()");
        if (!did_wait_need_post_) throw std::runtime_error("post() called when wait() was required.");
        if (node_->notifySourceReadComplete(slot_index_)) sem_post(&node_->write_barrier);
        did_wait_need_post_ = false;
    }
    uint64_t write_number() const { return node_ == nullptr ? 0 : node_->write_number(); }
    SourceState state() const { return state_; }

protected:
    // waits for the SINK to bind, then maps the object segment and checks its type
    SourceState connect_object(uint32_t type_hash)
    {
        if (state_ != SourceState::TOUCHED) throw std::runtime_error("A source can only connect() after it has touch()ed a node.");
        if (node_->sink_state() != NodeState::SINK_BOUND) {
            if (wait() != NodeState::SINK_BOUND) return SourceState::ERR_CONNECT;  // can occur at quit
            // self post: makes the first call to wait() a 'freebie' (Source.h:166-170)
            sem_post(&node_->read_barrier(slot_index_));
            did_wait_need_post_ = false;
        }
        obj_shmem_.open(obj_address_, 0, Shmem::OPEN_ONLY);
        const ObjectHeader *h = reinterpret_cast<const ObjectHeader *>(obj_shmem_.base());
        for (int i = 0; i < 2000 && h->magic != ObjectHeader::MAGIC; ++i) usleep(1000);
        if (h->magic != ObjectHeader::MAGIC || h->type_hash != type_hash) {
            state_ = SourceState::ERR_TYPEMIS;
            throw std::runtime_error("Type mismatch: Source<T> can only connect to Node<T>.");
        }
        sh_object_ = reinterpret_cast<T *>(obj_shmem_.base() + OBJ_OFFSET);
        state_ = SourceState::CONNECTED;
        return SourceState::CONNECTED;
    }
    Shmem node_shmem_, obj_shmem_;
    T *sh_object_{nullptr};
    Node *node_{nullptr};
    std::string address_, node_address_, obj_address_;
    size_t slot_index_{0};
    SourceState state_{SourceState::VIRGIN};
    bool did_wait_need_post_{false};
};

template <typename T>
class Source : public SourceBase<T> {
public:
    SourceState connect() { return this->connect_object(type_hash_of(typeid(T).name())); }
    T *retrieve() const
    {
        if (this->state_ < SourceState::CONNECTED) throw std::runtime_error("Source must be connected before shared object is retrieved.");
        return this->sh_object_;
    }
    T clone() const
    {
        if (this->state_ < SourceState::CONNECTED) throw std::runtime_error("Source must be connected before shared object is cloned.");
        return *this->sh_object_;
    }
};

template <>
class Source<Frame> : public SourceBase<SharedFrameHeader> {
public:
    SourceState connect()  // Source.h:315-368
    {
        const SourceState rc = connect_object(type_hash_of(typeid(SharedFrameHeader).name()));
        if (rc != SourceState::CONNECTED) return rc;
        const FrameParams &p = sh_object_->params;
        uint8_t *base = obj_shmem_.base();
        frame_ = Frame((int)p.rows, (int)p.cols, p.color, base + sh_object_->data_offset, base + sh_object_->sample_offset);
        parameters_ = p;
        return rc;
    }
    SourceState connect(PixelColor color)  // Source.h:300-313
    {
        const SourceState rc = connect();
        if (rc == SourceState::CONNECTED && frame_.color() != color)
            throw std::runtime_error("Component requires frame source with pixels of type " + color_str(color) +
                                     ". Maybe use oat-framefilt col?");
        return rc;
    }
    const Frame *retrieve() const { return &frame_; }
    void copyTo(Frame &frame) const { frame_.copyTo(frame); }
    FrameParams parameters() const { return parameters_; }
    void *pixels() const { return obj_shmem_.base() + sh_object_->data_offset; }
    const SharedFrameHeader *header() const { return sh_object_; }

private:
    Frame frame_;
    FrameParams parameters_;
};

// OAT_B200_TIMING=1: where a token's time goes inside a component -- named laps, averaged per token, printed by the
// component at end of stream (tools/graph_bench.py reads them).  Costs two clock reads per lap when on, a branch when off.
template <int N>
struct StageClock {
    const char *names[N] = {};
    bool on = getenv("OAT_B200_TIMING") != nullptr;
    double t[N] = {};
    uint64_t n = 0;
    std::chrono::steady_clock::time_point last{};
    explicit StageClock(std::initializer_list<const char *> l) { int i = 0; for (const char *x : l) if (i < N) names[i++] = x; }
    void start() { if (on) last = std::chrono::steady_clock::now(); }
    void lap(int i)
    {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        t[i] += std::chrono::duration<double, std::micro>(now - last).count();
        last = now;
    }
    void report(const std::string &who) const
    {
        if (!on || !n) return;
        std::string line = who + ": per frame (us):";
        char buf[96];
        for (int i = 0; i < N; ++i) {
            snprintf(buf, sizeof buf, "%s %s %.3f", i ? "," : "", names[i], t[i] / (double)n);
            line += buf;
        }
        snprintf(buf, sizeof buf, " over %llu frames\n", (unsigned long long)n);
        fputs((line + buf).c_str(), stderr);
    }
};

// OAT_B200_TIMING=1: wall-clock time at which the 1st, 10th, 100th, ... and the last token left a component, so that a
// harness can take a rate that excludes start-up (the first frame pays for model allocation, lazy module loading and
// the GPU's clock ramp: anything from 10 ms to a second).
struct TokenClock {
    bool on = getenv("OAT_B200_TIMING") != nullptr;
    uint64_t n = 0, next = 1;
    std::string marks;
    double last = 0.0;
    void tick()
    {
        if (!on) return;
        ++n;
        last = std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count();
        if (n == next) {
            char buf[64];
            snprintf(buf, sizeof buf, " %llu:%.6f", (unsigned long long)n, last);
            marks += buf;
            next *= 10;
        }
    }
    void report(const std::string &who) const
    {
        if (on && n) fprintf(stderr, "%s: tokens out at:%s last(%llu):%.6f\n", who.c_str(), marks.c_str(), (unsigned long long)n, last);
    }
};

// ---- lib/base/Component.{h,cpp} ---------------------------------------------------------------------------
class Component {
public:
    Component();
    virtual ~Component() = default;
    // run until end of stream or SIGINT (Component.cpp:50-76)
    void run()
    {
        if (!connectToNode()) return;
        bool end_of_stream = false;
        while (!end_of_stream && !quit) end_of_stream = process() != 0;
        if (getenv("OAT_B200_TIMING")) {  // (measurement harness: when did the last token leave this component)
            const double now = std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count();
            fprintf(stderr, "%s: end of stream at %.6f\n", name().c_str(), now);
        }
    }
    virtual std::string name() const = 0;

protected:
    virtual bool connectToNode() = 0;
    virtual int process() = 0;  // 0 = more, 1 = end of stream
};

}  // namespace oat
