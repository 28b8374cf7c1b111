// gpu.h -- the C++ host side of the C ABI (include/oatgpu.h): RAII handles that rethrow oat_status
// failures as std::runtime_error so that main() keeps the reference's "name: message, exit -1"
// contract (src/framefilter/main.cpp:278-295).  This is the binding a maintainer of the reference
// adds to its FrameFilter / PositionDetector subclasses.
#pragma once
#include <stdexcept>
#include <string>

#include "../../include/oatgpu.h"

namespace oat {
namespace gpu {

inline void ck(int status)
{
    if (status != OAT_OK) throw std::runtime_error(std::string("liboatgpu: ") + oat_last_error());
}

struct Context {
    oat_ctx *h{nullptr};
    explicit Context(int device_index) { ck(oat_ctx_create(device_index, &h)); }
    ~Context() { oat_ctx_destroy(h); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
};

// device staging buffer
struct DeviceBuffer {
    oat_ctx *ctx;
    void *p{nullptr};
    DeviceBuffer(Context &c, size_t bytes) : ctx(c.h) { ck(oat_alloc_device(ctx, bytes, &p)); }
    ~DeviceBuffer() { oat_free_device(ctx, p); }
    DeviceBuffer(const DeviceBuffer &) = delete;
    DeviceBuffer &operator=(const DeviceBuffer &) = delete;
    uint8_t *u8() const { return static_cast<uint8_t *>(p); }
};

// page-lock a shared-memory mapping for the life of the object (HOST_PINNED frame variant)
struct HostRegistration {
    void *p{nullptr};
    HostRegistration(void *ptr, size_t bytes) : p(ptr) { ck(oat_register_host(ptr, bytes)); }
    ~HostRegistration() { oat_unregister_host(p); }
    HostRegistration(const HostRegistration &) = delete;
    HostRegistration &operator=(const HostRegistration &) = delete;
};

// a device frame published by another process (SharedFrameHeader memory kind DEVICE)
struct IpcImport {
    oat_ctx *ctx;
    void *p{nullptr};
    IpcImport(Context &c, const unsigned char handle[64]) : ctx(c.h) { ck(oat_ipc_open(ctx, handle, &p)); }
    ~IpcImport() { oat_ipc_close(ctx, p); }
    IpcImport(const IpcImport &) = delete;
    IpcImport &operator=(const IpcImport &) = delete;
};

}  // namespace gpu
}  // namespace oat
