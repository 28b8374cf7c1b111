// shmemdf_test -- the reference's transport tests (test/shmemdf/{Node,Sink,Source,Helpers,concurrency}
// _test.cpp, Catch BDD) re-expressed against the POSIX re-statement in oat_host.h, plus the Frame
// specialisations the reference leaves untested ("TODO: specialization tests", Source_test.cpp:218)
// and an end-to-end frameserve -> Source<Frame> check of the synthetic stream.  Like the reference's
// tests it uses REAL shared memory (names prefixed oatb200test_), no mocks.  Exit code = failures.
#include <sys/wait.h>

#include <future>
#include <iostream>
#include <thread>
#include <vector>

#include "oat_host.h"
#include "synth.h"

namespace oat { volatile sig_atomic_t quit = 0; Component::Component() {} }
using namespace oat;

static int failures = 0, checks = 0;
#define CHECK(cond) do { ++checks; if (!(cond)) { ++failures; std::cerr << "FAIL " << __FILE__ << ":" << __LINE__ << ": " #cond "\n"; } } while (0)
#define CHECK_THROWS(expr) do { ++checks; bool t_ = false; try { expr; } catch (const std::exception &) { t_ = true; } if (!t_) { ++failures; std::cerr << "FAIL " << __FILE__ << ":" << __LINE__ << ": no throw: " #expr "\n"; } } while (0)
#define CHECK_NOTHROW(expr) do { ++checks; try { expr; } catch (const std::exception &e_) { ++failures; std::cerr << "FAIL " << __FILE__ << ":" << __LINE__ << ": threw " << e_.what() << ": " #expr "\n"; } } while (0)

static const std::string node_addr = "oatb200test_node";
static void scrub(const std::string &a) { Shmem::remove(a + "_node"); Shmem::remove(a + "_obj"); }

// Node_test.cpp:28-75
static void test_node()
{
    Node n;
    n.init();
    size_t idx = 0;
    for (size_t i = 0; i < Node::NUM_SLOTS; ++i) { CHECK(n.acquireSlot(idx) == 0); CHECK(idx == i); }
    CHECK(n.source_ref_count() == Node::NUM_SLOTS);
    CHECK(n.acquireSlot(idx) == -1);  // 11th slot refused
    CHECK(n.releaseSlot(3) == 0);
    CHECK(n.source_ref_count() == Node::NUM_SLOTS - 1);
    CHECK(n.acquireSlot(idx) == 0);
    CHECK(idx == 3);
    CHECK(n.releaseSlot(Node::NUM_SLOTS) == -1);
    CHECK_THROWS(n.read_barrier(Node::NUM_SLOTS));  // index out of range
    n.releaseSlot(7);
    CHECK_THROWS(n.read_barrier(7));  // not bound to this node
    CHECK_NOTHROW(n.read_barrier(5)); // every slot is usable here
}

// Sink_test.cpp:34-116
static void test_sink()
{
    scrub(node_addr);
    {
        Sink<int> sink;
        CHECK_THROWS(sink.wait());  // must be bound first
        CHECK_NOTHROW(sink.bind(node_addr));
        CHECK_THROWS(sink.bind(node_addr));  // a sink binds a single time
        Sink<int> sink2;
        CHECK_THROWS(sink2.bind(node_addr));  // one SINK per address
        CHECK_THROWS(sink.post());            // post before wait
        CHECK_NOTHROW(sink.wait());
        CHECK_THROWS(sink.wait());            // wait twice
        CHECK_NOTHROW(sink.post());
        CHECK(sink.retrieve() != nullptr);
    }
    // sink left with no sources: segments are unlinked, the address is free again
    Sink<int> again;
    CHECK_NOTHROW(again.bind(node_addr));
}

// Source_test.cpp:34-215
static void test_source()
{
    scrub(node_addr);
    {
        Sink<int> sink;
        sink.bind(node_addr);
        std::vector<std::unique_ptr<Source<int>>> sources;
        for (size_t i = 0; i < Node::NUM_SLOTS; ++i) {
            sources.emplace_back(new Source<int>());
            CHECK_NOTHROW(sources.back()->touch(node_addr));
            CHECK(sources.back()->connect() == SourceState::CONNECTED);
        }
        Source<int> eleventh;
        CHECK_NOTHROW(eleventh.touch(node_addr));
        CHECK(eleventh.state() == SourceState::ERR_NODEFULL);
        CHECK_THROWS(eleventh.connect());
        Source<int> virgin;
        CHECK_THROWS(virgin.wait());
        CHECK_THROWS(virgin.connect());
        CHECK_THROWS(sources[0]->touch(node_addr));  // single touch
        CHECK_THROWS(sources[0]->post());            // post before wait
        // shared object mutation is visible on both sides (Source_test.cpp:163-215)
        sink.wait();
        *sink.retrieve() = 42;
        sink.post();
        for (auto &s : sources) {
            CHECK(s->wait() == NodeState::SINK_BOUND);
            CHECK(*s->retrieve() == 42);
            CHECK(s->clone() == 42);
            CHECK_THROWS(s->wait());  // wait twice
            s->post();
        }
        CHECK(sources[0]->write_number() == 1);
    }
    scrub(node_addr);
    {   // Source<float> cannot connect to Sink<int> (Source_test.cpp:142-161)
        Sink<int> sink;
        sink.bind(node_addr);
        Source<float> wrong;
        wrong.touch(node_addr);
        CHECK_THROWS(wrong.connect());
        CHECK(wrong.state() == SourceState::ERR_TYPEMIS);
    }
}

// Helpers_test.cpp:28-60 + token semantics
static void test_tokens()
{
    Sample s;
    s.set_rate_hz(100.0);
    CHECK(std::fabs(s.period_sec() - 0.01) < 1e-12);
    CHECK(s.period_microseconds() == 10000);
    for (int i = 0; i < 5; ++i) s.incrementCount();
    CHECK(s.count() == 5 && s.microseconds() == 50000);
    Position2D a("label_a"), b("label_b");
    a.position_valid = true;
    a.position.x = 100.5;
    a.position.y = 7.25;
    a.set_sample(s);
    b = a;
    CHECK(std::string(b.label()) == "label_b");  // operator= skips the label (Position2D.h:84-105)
    CHECK(b.position_valid && b.position.x == 100.5 && b.sample_count() == 5);
    CHECK(serializePosition(b) ==
          "{\"tick\":5,\"usec\":50000,\"unit\":0,\"pos_ok\":true,\"pos_xy\":[100.5,7.25],\"vel_ok\":false,\"head_ok\":false,\"reg_ok\":false}");
    Position2D c("");
    CHECK(serializePosition(c) == "{\"tick\":0,\"usec\":0,\"unit\":0,\"pos_ok\":false,\"vel_ok\":false,\"head_ok\":false,\"reg_ok\":false}");
    char rec[Position2D::NPY_DTYPE_BYTES];
    packPosition(b, rec);
    uint64_t tick; double px;
    std::memcpy(&tick, rec, 8); std::memcpy(&px, rec + 21, 8);
    CHECK(tick == 5 && px == 100.5 && rec[20] == 1);
    CHECK(json_double(3.0) == "3.0" && json_double(0.123456789) == "0.12346" && json_double(960.5) == "960.5");
}

// concurrency_test.cpp:87-527
static void test_concurrency()
{
    using namespace std::chrono_literals;
    scrub(node_addr);
    {   // source blocks until sink posts; sink blocks until every source has posted (:87-192)
        Sink<int> sink;
        sink.bind(node_addr);
        Source<int> s0, s1;
        s0.touch(node_addr); s0.connect();
        s1.touch(node_addr); s1.connect();
        auto f0 = std::async(std::launch::async, [&] { return s0.wait(); });
        std::this_thread::sleep_for(30ms);
        CHECK(f0.wait_for(0ms) != std::future_status::ready);
        sink.wait();
        *sink.retrieve() = 1;
        sink.post();
        CHECK(f0.wait_for(500ms) == std::future_status::ready);
        auto fs = std::async(std::launch::async, [&] { sink.wait(); return 0; });
        std::this_thread::sleep_for(30ms);
        CHECK(fs.wait_for(0ms) != std::future_status::ready);  // s0 and s1 have not posted
        s0.post();
        std::this_thread::sleep_for(30ms);
        CHECK(fs.wait_for(0ms) != std::future_status::ready);  // s1 still has not read
        s1.wait();
        s1.post();
        CHECK(fs.wait_for(500ms) == std::future_status::ready);
        sink.post();
        s0.wait(); s0.post(); s1.wait(); s1.post();
    }
    scrub(node_addr);
    {   // bind / connect order is irrelevant: the source blocks in connect() until the sink binds (:239-421)
        Source<int> src;
        src.touch(node_addr);
        auto fc = std::async(std::launch::async, [&] { return src.connect(); });
        std::this_thread::sleep_for(30ms);
        CHECK(fc.wait_for(0ms) != std::future_status::ready);
        Sink<int> sink;
        sink.bind(node_addr, 7);
        sink.wait();
        sink.post();
        CHECK(fc.wait_for(1000ms) == std::future_status::ready);
        CHECK(fc.get() == SourceState::CONNECTED);
        CHECK(src.wait() == NodeState::SINK_BOUND);  // the freebie
        CHECK(*src.retrieve() == 7);
        src.post();
    }
    scrub(node_addr);
    {   // END propagates: a source waiting on a departed sink returns END (Source.h:202-210)
        auto sink = std::make_unique<Sink<int>>();
        sink->bind(node_addr);
        Source<int> src;
        src.touch(node_addr);
        src.connect();
        auto fw = std::async(std::launch::async, [&] { return src.wait(); });
        std::this_thread::sleep_for(20ms);
        sink.reset();
        CHECK(fw.wait_for(1000ms) == std::future_status::ready);
        CHECK(fw.get() == NodeState::END);
    }
    scrub(node_addr);
    {   // late-joining source (:475-527) and source destruction during the sink's wait (:423-472)
        Sink<int> sink;
        sink.bind(node_addr);
        for (int i = 0; i < 3; ++i) { sink.wait(); *sink.retrieve() = i; sink.post(); }
        auto late = std::make_unique<Source<int>>();
        late->touch(node_addr);
        CHECK(late->connect() == SourceState::CONNECTED);
        sink.wait();
        *sink.retrieve() = 99;
        sink.post();
        CHECK(late->wait() == NodeState::SINK_BOUND);
        CHECK(*late->retrieve() == 99);
        auto fs = std::async(std::launch::async, [&] { sink.wait(); return 0; });
        std::this_thread::sleep_for(20ms);
        late.reset();  // leaves without posting: the sink must not deadlock
        CHECK(fs.wait_for(1000ms) == std::future_status::ready);
        sink.post();
    }
    scrub(node_addr);
}

// Frame specialisations: pixels + Sample travel together; colour is checked at connect
static void test_frames()
{
    scrub(node_addr);
    const int rows = 48, cols = 64;
    Sink<Frame> sink;
    sink.bind(node_addr, (size_t)rows * cols * 3);
    Frame shared = sink.retrieve(rows, cols, 3, PIX_BGR);
    shared.set_rate_hz(30.0);
    Source<Frame> hsv_only;
    hsv_only.touch(node_addr);
    CHECK_THROWS(hsv_only.connect(PIX_HSV));  // "Maybe use oat-framefilt col?"
    Source<Frame> src;
    src.touch(node_addr);
    CHECK(src.connect(PIX_BGR) == SourceState::CONNECTED);
    CHECK(src.parameters().rows == (size_t)rows && src.parameters().cols == (size_t)cols && src.parameters().bytes == (size_t)rows * cols * 3);
    Frame internal;
    for (uint32_t t = 0; t < 3; ++t) {
        sink.wait();
        synth::frame(shared.data(), rows, cols, 1000, t);
        shared.incrementSampleCount();
        sink.post();
        CHECK(src.wait() == NodeState::SINK_BOUND);
        src.copyTo(internal);
        src.post();
        // hsv_only holds a slot (ERR state after the throw releases nothing until destruction): drain it
        std::vector<uint8_t> want((size_t)rows * cols * 3);
        synth::frame(want.data(), rows, cols, 1000, t);
        CHECK(std::memcmp(internal.data(), want.data(), want.size()) == 0);
        CHECK(internal.sample().count() == t + 1);
        CHECK(internal.color() == PIX_BGR);
        if (hsv_only.state() == SourceState::CONNECTED) { hsv_only.wait(); hsv_only.post(); }
    }
}

// frameserve (separate process) -> Source<Frame>: every frame of the synthetic stream arrives bit-exact, in order
// the additions the GPU components rely on: Source::try_wait, the persistent tag, Sink::end_and_linger
static void test_nonblocking_and_persistent()
{
    const std::string addr = "oatb200test_nb";
    scrub(addr);
    {
        Sink<Frame> sink;
        sink.bind(addr, 4 * 4 * 3, false);
        Frame shared = sink.retrieve(4, 4, 3, PIX_BGR);
        sink.set_persistent(true);
        sink.announce();
        Source<Frame> src;
        src.touch(addr);
        NodeState st = NodeState::UNDEFINED;
        CHECK_THROWS(src.try_wait(&st));  // not connected yet
        CHECK(src.connect(PIX_BGR) == SourceState::CONNECTED);
        CHECK(src.header()->persistent == 1u);
        CHECK(!src.try_wait(&st) && st == NodeState::SINK_BOUND);  // nothing published yet: no token, no blocking
        sink.wait();
        shared.incrementSampleCount();
        sink.post();
        CHECK(src.try_wait(&st) && st == NodeState::SINK_BOUND);  // the token is there
        CHECK_THROWS(src.try_wait(&st));                          // post() is required first
        CHECK(src.retrieve()->sample().count() == 1);
        src.post();
        CHECK(!src.try_wait(&st));
        // end_and_linger: END is flagged at once, the SINK stays until the SOURCE has detached
        auto gone = std::async(std::launch::async, [&] {
            const auto t0 = std::chrono::steady_clock::now();
            sink.end_and_linger(2000);
            return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        });
        std::this_thread::sleep_for(std::chrono::milliseconds(50));
        CHECK(!src.try_wait(&st) && st == NodeState::END);
        CHECK(gone.wait_for(std::chrono::milliseconds(0)) != std::future_status::ready);  // still lingering: a SOURCE is attached
    }
    scrub(addr);
    {
        Sink<Frame> sink;
        sink.bind(addr, 4 * 4 * 3);
        sink.retrieve(4, 4, 3, PIX_BGR);
        auto src = std::make_unique<Source<Frame>>();
        src->touch(addr);
        CHECK(src->connect() == SourceState::CONNECTED);
        CHECK(src->header()->persistent == 0u);  // the default: frames are the SINK's again after post()
        auto gone = std::async(std::launch::async, [&] { sink.end_and_linger(5000); return true; });
        std::this_thread::sleep_for(std::chrono::milliseconds(20));
        src.reset();  // the SOURCE detaches: the SINK may leave
        CHECK(gone.wait_for(std::chrono::milliseconds(1000)) == std::future_status::ready);
    }
    scrub(addr);
}

static void test_pipeline(const std::string &frameserve)
{
    const std::string addr = "oatb200test_pipe";
    scrub(addr);
    const int rows = 60, cols = 80, n = 12;
    Source<Frame> src;
    src.touch(addr);  // consumer first, like the reference's example scripts
    const pid_t pid = fork();
    if (pid == 0) {
        execl(frameserve.c_str(), "oat-frameserve", "synth", addr.c_str(), "--rows", "60", "--cols", "80", "--num-samples", "12", "--fps", "200", (char *)nullptr);
        _exit(127);
    }
    CHECK(src.connect(PIX_BGR) == SourceState::CONNECTED);
    Frame internal;
    std::vector<uint8_t> want((size_t)rows * cols * 3);
    int got = 0;
    for (;;) {
        if (src.wait() == NodeState::END) break;
        src.copyTo(internal);
        src.post();
        synth::frame(want.data(), rows, cols, 1000, (uint32_t)got);
        CHECK(std::memcmp(internal.data(), want.data(), want.size()) == 0);
        CHECK(internal.sample().count() == (uint64_t)got + 1);
        CHECK(internal.sample().period_microseconds() == 5000);
        ++got;
        if (got > n) break;
    }
    CHECK(got == n);
    int status = 0;
    waitpid(pid, &status, 0);
    CHECK(WIFEXITED(status) && WEXITSTATUS(status) == 0);
    scrub(addr);
}

// helper mode for the position egress test: publish N positions on ADDR (what a posidet SINK does)
static int emit_positions(const std::string &addr, int n)
{
    Sink<Position2D> sink;  // the consumer may already have created the node: do not scrub it away
    sink.bind(addr, addr);
    Position2D *shared = sink.retrieve();
    Sample clock;
    clock.set_rate_hz(50.0);
    // wait for a reader so that nothing is published into the void
    for (int i = 0; i < 3000; ++i) {
        bool attached = false;
        {
            Shmem probe;
            probe.open(addr + "_node", sizeof(Node), Shmem::OPEN_OR_CREATE);
            attached = reinterpret_cast<Node *>(probe.base())->source_ref_count() > 0;
        }
        if (attached) break;
        usleep(1000);
    }
    for (int t = 0; t < n; ++t) {
        Position2D p("");
        clock.incrementCount();
        p.set_sample(clock);
        p.position_valid = (t % 3) != 0;
        p.position.x = 10.5 + t;
        p.position.y = 0.125 * t;
        sink.wait();
        *shared = p;
        sink.post();
    }
    sink.wait();
    return 0;
}

// helper mode for the frame-server tests: read frames from ADDR until the SINK leaves; per frame one line
// "tick usec rows cols channels color fnv1a(pixels)" on stdout (a plain Source<Frame>, like any reference component)
static int dump_frames(const std::string &addr)
{
    Source<Frame> src;
    src.touch(addr);
    if (src.connect() != SourceState::CONNECTED) return 2;
    Frame internal;
    for (;;) {
        if (src.wait() == NodeState::END) break;
        src.copyTo(internal);
        src.post();
        uint64_t hsh = 1469598103934665603ull;
        const uint8_t *p = internal.data();
        for (size_t i = 0; i < internal.bytes(); ++i) hsh = (hsh ^ p[i]) * 1099511628211ull;
        std::cout << internal.sample().count() << " " << internal.sample().microseconds() << " " << internal.rows() << " "
                  << internal.cols() << " " << internal.channels() << " " << (int)internal.color() << " " << hsh << std::endl;
    }
    return 0;
}

// helper mode for the process-graph tests: block until every ADDR has at least N SOURCEs attached (a SINK only waits
// for SOURCEs that have touched its node, Sink.h:93-116, so a server started too early serves into the void);
// returns 0 when they have, 3 after TIMEOUT seconds
static int wait_sources(int n, int timeout_s, const std::vector<std::string> &addrs)
{
    const auto until = std::chrono::steady_clock::now() + std::chrono::seconds(timeout_s);
    for (const auto &addr : addrs) {
        for (;;) {
            bool ok = false;
            try {
                Shmem shm;
                shm.open(addr + "_node", sizeof(Node), Shmem::OPEN_ONLY);
                const Node *node = reinterpret_cast<const Node *>(shm.base());
                ok = shm.bytes() >= sizeof(Node) && node->ready() && (int)node->source_ref_count() >= n;
            } catch (const std::exception &) {}  // not there yet
            if (ok) break;
            if (std::chrono::steady_clock::now() >= until) return 3;
            usleep(2000);
        }
    }
    return 0;
}

// helper mode for transport-rate measurements: take tokens from ADDR without touching the pixels until the SINK leaves
static int count_frames(const std::string &addr)
{
    Source<Frame> src;
    src.touch(addr);
    if (src.connect() != SourceState::CONNECTED) return 2;
    uint64_t n = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        if (src.wait() == NodeState::END) break;
        if (n == 0) t0 = std::chrono::steady_clock::now();
        src.post();
        ++n;
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout << n << " tokens, " << (n > 1 ? 1e6 * sec / (double)(n - 1) : 0.0) << " us per token" << std::endl;
    return 0;
}

int main(int argc, char **argv)
{
    if (argc > 2 && std::string(argv[1]) == "count-frames") return count_frames(argv[2]);
    if (argc > 3 && std::string(argv[1]) == "emit-positions") return emit_positions(argv[2], std::atoi(argv[3]));
    if (argc > 2 && std::string(argv[1]) == "dump-frames") return dump_frames(argv[2]);
    if (argc > 4 && std::string(argv[1]) == "wait-sources")
        return wait_sources(std::atoi(argv[2]), std::atoi(argv[3]), std::vector<std::string>(argv + 4, argv + argc));
    test_node();
    test_sink();
    test_source();
    test_tokens();
    test_concurrency();
    test_frames();
    test_nonblocking_and_persistent();
    if (argc > 1) test_pipeline(argv[1]);
    scrub(node_addr);
    std::cout << "shmemdf_test: " << checks << " checks, " << failures << " failures\n";
    return failures;
}
