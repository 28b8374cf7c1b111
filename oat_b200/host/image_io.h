// image_io.h -- still images for `oat-frameserve test` (the reference reads them with cv::imread,
// src/frameserver/TestFrame.cpp:83; OpenCV's C++ library is not available to this build).  Dependency-free
// readers for the lossless 8-bit formats a test image can be handed over in: binary PPM (P6, BGR order is NOT
// assumed: PPM is RGB and is swapped to Oat's BGR), binary PGM (P5) and NumPy .npy (uint8, C order, shape
// (rows, cols, 3) taken as BGR -- what cv2.imread returns -- or (rows, cols)).
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace oat {

struct Image {
    int rows = 0, cols = 0, channels = 0;  // channels: 1 (GREY) or 3 (BGR)
    std::vector<uint8_t> data;             // rows * cols * channels, row-major, interleaved
};

namespace detail {
inline int pnm_int(std::istream &f)
{
    for (;;) {  // skip whitespace and '#' comments
        int c = f.peek();
        if (c == '#') {
            std::string skip;
            std::getline(f, skip);
        } else if (c == ' ' || c == '\t' || c == '\r' || c == '\n') {
            f.get();
        } else {
            break;
        }
    }
    int v = -1;
    f >> v;
    return v;
}
}  // namespace detail

inline Image read_image(const std::string &path)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("File \"" + path + "\" could not be read.");  // TestFrame.cpp:85-86
    char magic[6] = {0};
    f.read(magic, 2);
    Image img;
    if (magic[0] == 'P' && (magic[1] == '6' || magic[1] == '5')) {
        const int w = detail::pnm_int(f), h = detail::pnm_int(f), maxv = detail::pnm_int(f);
        if (w <= 0 || h <= 0 || maxv != 255) throw std::runtime_error("File \"" + path + "\": only 8-bit binary PPM/PGM images are supported.");
        f.get();  // the single whitespace byte after maxval
        img.rows = h;
        img.cols = w;
        img.channels = magic[1] == '6' ? 3 : 1;
        img.data.resize((size_t)w * h * img.channels);
        f.read(reinterpret_cast<char *>(img.data.data()), (std::streamsize)img.data.size());
        if ((size_t)f.gcount() != img.data.size()) throw std::runtime_error("File \"" + path + "\" is truncated.");
        if (img.channels == 3)
            for (size_t i = 0; i < img.data.size(); i += 3) std::swap(img.data[i], img.data[i + 2]);  // RGB -> BGR
        return img;
    }
    f.read(magic + 2, 4);
    if (std::memcmp(magic, "\x93NUMPY", 6) == 0) {
        unsigned char ver[2];
        f.read(reinterpret_cast<char *>(ver), 2);
        uint32_t hlen = 0;
        if (ver[0] == 1) {
            unsigned char b[2];
            f.read(reinterpret_cast<char *>(b), 2);
            hlen = b[0] | (b[1] << 8);
        } else {
            unsigned char b[4];
            f.read(reinterpret_cast<char *>(b), 4);
            hlen = b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24);
        }
        std::string hdr(hlen, '\0');
        f.read(&hdr[0], hlen);
        if (hdr.find("'|u1'") == std::string::npos && hdr.find("'u1'") == std::string::npos && hdr.find("'<u1'") == std::string::npos)
            throw std::runtime_error("File \"" + path + "\": .npy images must be uint8.");
        if (hdr.find("'fortran_order': False") == std::string::npos)
            throw std::runtime_error("File \"" + path + "\": .npy images must be C-ordered.");
        const size_t a = hdr.find("'shape': ("), b = hdr.find(')', a);
        if (a == std::string::npos || b == std::string::npos) throw std::runtime_error("File \"" + path + "\": malformed .npy header.");
        std::vector<long> dims;
        std::string num;
        for (size_t i = a + 10; i <= b; ++i) {
            const char c = hdr[i];
            if (c >= '0' && c <= '9') {
                num += c;
            } else if (!num.empty()) {
                dims.push_back(std::stol(num));
                num.clear();
            }
        }
        if (!(dims.size() == 2 || (dims.size() == 3 && (dims[2] == 3 || dims[2] == 1))))
            throw std::runtime_error("File \"" + path + "\": .npy images must have shape (rows, cols) or (rows, cols, 3).");
        img.rows = (int)dims[0];
        img.cols = (int)dims[1];
        img.channels = dims.size() == 3 ? (int)dims[2] : 1;
        img.data.resize((size_t)img.rows * img.cols * img.channels);
        f.read(reinterpret_cast<char *>(img.data.data()), (std::streamsize)img.data.size());
        if ((size_t)f.gcount() != img.data.size()) throw std::runtime_error("File \"" + path + "\" is truncated.");
        return img;
    }
    throw std::runtime_error("File \"" + path + "\" could not be read.");
}

// A clip for `oat-frameserve file` (the reference decodes a video file with cv::VideoCapture,
// src/frameserver/FileReader.cpp:59, :103-131; no codec is available to this build): a NumPy .npy of uint8 frames,
// C order, shape (frames, rows, cols, 3) taken as BGR or (frames, rows, cols) taken as GREY -- lossless, and mapped,
// not read, so that a long clip costs no host memory.
struct Clip {
    size_t frames = 0;
    int rows = 0, cols = 0, channels = 0;
    const uint8_t *data = nullptr;  // frames * rows * cols * channels
    void *map = nullptr;
    size_t map_bytes = 0;
    size_t frame_bytes() const { return (size_t)rows * cols * channels; }
    Clip() = default;
    Clip(const Clip &) = delete;
    Clip &operator=(const Clip &) = delete;
    ~Clip();
    void open(const std::string &path);
};

// cv::imread(file, IMREAD_GRAYSCALE) for a colour file / IMREAD_COLOR for a grey one (lib/datatypes/Color.h imread_code)
inline Image to_channels(const Image &in, int channels)
{
    if (in.channels == channels) return in;
    Image out;
    out.rows = in.rows;
    out.cols = in.cols;
    out.channels = channels;
    out.data.resize((size_t)in.rows * in.cols * channels);
    const size_t n = (size_t)in.rows * in.cols;
    if (channels == 3) {
        for (size_t i = 0; i < n; ++i) out.data[3 * i] = out.data[3 * i + 1] = out.data[3 * i + 2] = in.data[i];
    } else {  // BGR -> grey, OpenCV's fixed-point weights (R 0.299, G 0.587, B 0.114; 14-bit: 4899, 9617, 1868)
        for (size_t i = 0; i < n; ++i) {
            const int b = in.data[3 * i], g = in.data[3 * i + 1], r = in.data[3 * i + 2];
            out.data[i] = (uint8_t)((b * 1868 + g * 9617 + r * 4899 + (1 << 13)) >> 14);
        }
    }
    return out;
}

inline Clip::~Clip()
{
    if (map) munmap(map, map_bytes);
}
inline void Clip::open(const std::string &path)
{
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) throw std::runtime_error("File \"" + path + "\" could not be read.");
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 16) {
        ::close(fd);
        throw std::runtime_error("File \"" + path + "\" could not be read.");
    }
    map_bytes = (size_t)st.st_size;
    map = mmap(nullptr, map_bytes, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (map == MAP_FAILED) {
        map = nullptr;
        throw std::runtime_error("File \"" + path + "\" could not be mapped.");
    }
    const unsigned char *b = static_cast<const unsigned char *>(map);
    if (std::memcmp(b, "\x93NUMPY", 6) != 0) throw std::runtime_error("File \"" + path + "\": clips must be NumPy .npy files (uint8, frames x rows x cols [x 3]).");
    size_t hoff, hlen;
    if (b[6] == 1) {
        hlen = b[8] | (b[9] << 8);
        hoff = 10;
    } else {
        hlen = b[8] | (b[9] << 8) | (b[10] << 16) | ((size_t)b[11] << 24);
        hoff = 12;
    }
    if (hoff + hlen > map_bytes) throw std::runtime_error("File \"" + path + "\" is truncated.");
    const std::string hdr(reinterpret_cast<const char *>(b + hoff), hlen);
    if (hdr.find("u1'") == std::string::npos) throw std::runtime_error("File \"" + path + "\": .npy clips must be uint8.");
    if (hdr.find("'fortran_order': False") == std::string::npos) throw std::runtime_error("File \"" + path + "\": .npy clips must be C-ordered.");
    const size_t a = hdr.find("'shape': ("), e = hdr.find(')', a);
    if (a == std::string::npos || e == std::string::npos) throw std::runtime_error("File \"" + path + "\": malformed .npy header.");
    std::vector<long> dims;
    std::string num;
    for (size_t i = a + 10; i <= e; ++i) {
        const char c = hdr[i];
        if (c >= '0' && c <= '9') {
            num += c;
        } else if (!num.empty()) {
            dims.push_back(std::stol(num));
            num.clear();
        }
    }
    if (!(dims.size() == 3 || (dims.size() == 4 && dims[3] == 3)))
        throw std::runtime_error("File \"" + path + "\": .npy clips must have shape (frames, rows, cols) or (frames, rows, cols, 3).");
    frames = (size_t)dims[0];
    rows = (int)dims[1];
    cols = (int)dims[2];
    channels = dims.size() == 4 ? 3 : 1;
    if (hoff + hlen + frames * frame_bytes() > map_bytes) throw std::runtime_error("File \"" + path + "\" is truncated.");
    data = b + hoff + hlen;
}

}  // namespace oat
