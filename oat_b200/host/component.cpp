// lib/base/Component.cpp:33-48: the quit flag and its SIGINT handler
#include "oat_host.h"

namespace oat {
volatile sig_atomic_t quit = 0;
static void sigHandler(int) { quit = 1; }
Component::Component()
{
    struct sigaction sa;
    std::memset(&sa, 0, sizeof(sa));
    sa.sa_handler = sigHandler;  // no SA_RESTART: blocking waits return with EINTR and observe quit
    sigaction(SIGINT, &sa, nullptr);
}
}  // namespace oat
