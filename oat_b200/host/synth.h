// synth.h -- the deterministic synthetic stream of SURVEY.md 8(d) on the host (bit-identical to the
// device generator oat_synth_frame and to the test suite's generators).
#pragma once
#include <cstdint>

namespace oat {
namespace synth {
inline uint32_t fmix32(uint32_t h)
{
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}
inline void disc_centre(int rows, int cols, uint32_t t, int &cx, int &cy)
{
    cx = cols / 4 + (int)((7u * t) % (uint32_t)(cols / 2));
    cy = rows / 3 + (int)((4u * t) % (uint32_t)(rows / 3));
}
inline void frame(uint8_t *dst, int rows, int cols, uint32_t seed, uint32_t t)
{
    const uint32_t kbg = fmix32(seed ^ 0x9e3779b9u), knz = fmix32(seed + 0x7f4a7c15u * (t + 1u));
    const int r = rows / 20;
    int cx, cy;
    disc_centre(rows, cols, t, cx, cy);
    static const uint8_t disc[3] = {40, 220, 60};
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            const int dx = x - cx, dy = y - cy;
            const bool in = (t != 0) && (dx * dx + dy * dy <= r * r);
            for (int c = 0; c < 3; ++c) {
                const uint32_t idx = ((uint32_t)y * (uint32_t)cols + (uint32_t)x) * 3u + (uint32_t)c;
                const int bg = 40 + (int)(fmix32(idx ^ kbg) % 81u);
                const int nz = (int)(fmix32(idx ^ knz) % 7u) - 3;
                dst[((size_t)y * cols + x) * 3 + c] = in ? disc[c] : (uint8_t)(bg + nz);
            }
        }
}
}  // namespace synth
}  // namespace oat
