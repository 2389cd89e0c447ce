// wrh_writer.cpp — see wrh_writer.hpp.  ASCII "P3" PPM through ftruncate + mmap(MAP_SHARED); pixel lines of
// 1024-pixel chunks are formatted on the thread pool after a size pre-pass fixes every chunk's file offset.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "wrh_writer.hpp"

namespace wrh {

// ---- thread pool ---------------------------------------------------------------------------------------------------
ThreadPool::ThreadPool(size_t n_jobs) {
    if (n_jobs == 0) n_jobs = 1;
    workers_.reserve(n_jobs);
    for (size_t i = 0; i < n_jobs; ++i) {
        workers_.emplace_back([this] {
            std::unique_lock<std::mutex> lk(mu_);
            for (;;) {
                cv_job_.wait(lk, [this] { return stop_ || !jobs_.empty(); });
                if (stop_ && jobs_.empty()) return;
                runOne(lk);
            }
        });
    }
}

ThreadPool::~ThreadPool() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_job_.notify_all();
    for (auto& t : workers_) t.join();
}

bool ThreadPool::runOne(std::unique_lock<std::mutex>& lk) {
    if (jobs_.empty()) return false;
    auto job = std::move(jobs_.front());
    jobs_.pop();
    lk.unlock();
    job();
    lk.lock();
    if (--pending_ == 0) cv_done_.notify_all();
    return true;
}

void ThreadPool::spawnWg(std::function<void()> job) {
    {
        std::lock_guard<std::mutex> lk(mu_);
        jobs_.push(std::move(job));
        ++pending_;
    }
    cv_job_.notify_one();
}

void ThreadPool::waitAndWork() {
    std::unique_lock<std::mutex> lk(mu_);
    while (runOne(lk)) {}
    cv_done_.wait(lk, [this] { return pending_ == 0; });
}

// ---- quantisation --------------------------------------------------------------------------------------------------
std::array<uint8_t, 3> encodeColor(const Real rgb[3]) {  // writer.zig:68-94
    const Real rgb_max = 256.0;
    const Interval intensity{0.0, 0.999};
    std::array<uint8_t, 3> out{};
    for (int k = 0; k < 3; ++k) {
        Real c = rgb[k];
        if (std::isnan(c)) c = 0;  // clampNaN
        c = std::sqrt(c);          // gammaCorrection (gamma 2)
        out[static_cast<size_t>(k)] = static_cast<uint8_t>(rgb_max * intensity.clamp(c));
    }
    return out;
}

size_t sizeOfDigit(uint8_t digit) {  // writer.zig:107-114
    size_t result = 0x1;
    result <<= (digit > 9) ? 1 : 0;
    result |= (digit > 99) ? 1 : 0;
    return result;
}

size_t sizeOfLine(const std::array<uint8_t, 3>& pixel) {  // writer.zig:96-100
    return pixel.size() + sizeOfDigit(pixel[0]) + sizeOfDigit(pixel[1]) + sizeOfDigit(pixel[2]);
}

// ---- mmap handle (mmap.zig:4-35) -------------------------------------------------------------------------------------
namespace {

struct MmapHandlePosix {
    int fd = -1;
    uint8_t* ptr = nullptr;
    size_t size = 0;
    MmapHandlePosix(const std::string& path, size_t bytes) : size(bytes) {
        fd = ::open(path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
        if (fd < 0) throw std::runtime_error("cannot create " + path + ": " + std::strerror(errno));
        if (::ftruncate(fd, static_cast<off_t>(bytes)) != 0) {
            ::close(fd);
            throw std::runtime_error("ftruncate failed on " + path);
        }
        void* p = ::mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        if (p == MAP_FAILED) {
            ::close(fd);
            throw std::runtime_error("mmap failed on " + path);
        }
        ptr = static_cast<uint8_t*>(p);
    }
    ~MmapHandlePosix() {
        if (ptr) ::munmap(ptr, size);
        if (fd >= 0) ::close(fd);
    }
    void shrink(size_t bytes) {
        if (::ftruncate(fd, static_cast<off_t>(bytes)) != 0) throw std::runtime_error("ftruncate (shrink) failed");
    }
};

constexpr size_t kPixelNumBytes = sizeof("255 255 255\n") - 1;  // PPM_PIXEL_NUM_BYTES, writer.zig:11
constexpr size_t kChunkSize = 1024;                             // writer.zig:29

size_t formatPixel(uint8_t* out, const std::array<uint8_t, 3>& px) {  // "{d} {d} {d}\n", writer.zig:62
    size_t n = 0;
    for (int k = 0; k < 3; ++k) {
        const unsigned v = px[static_cast<size_t>(k)];
        if (v > 99) out[n++] = static_cast<uint8_t>('0' + v / 100);
        if (v > 9) out[n++] = static_cast<uint8_t>('0' + (v / 10) % 10);
        out[n++] = static_cast<uint8_t>('0' + v % 10);
        out[n++] = (k == 2) ? '\n' : ' ';
    }
    return n;
}

template <typename PixelAt>
size_t writeImpl(const WriterPPM& w, const std::string& out_path, size_t num_pixels, size_t num_cols, size_t num_rows, PixelAt pixel_at) {
    if (!w.thread_pool) throw std::runtime_error("WriterPPM: thread_pool is null");
    char header[64];
    const int header_len = std::snprintf(header, sizeof header, "P3\n%zu %zu\n255\n", num_cols, num_rows);  // writer.zig:9,18
    const size_t content_size = num_pixels * kPixelNumBytes + static_cast<size_t>(header_len);             // writer.zig:20
    MmapHandlePosix handle(out_path, content_size);
    std::memcpy(handle.ptr, header, static_cast<size_t>(header_len));

    size_t file_index = static_cast<size_t>(header_len);
    for (size_t data_index = 0; data_index < num_pixels; data_index += kChunkSize) {
        const size_t end = std::min(num_pixels, data_index + kChunkSize);
        size_t chunk_content_size = 0;  // serial size pre-pass, writer.zig:36-39
        for (size_t i = data_index; i < end; ++i) chunk_content_size += sizeOfLine(pixel_at(i));
        uint8_t* dst = handle.ptr + file_index;
        w.thread_pool->spawnWg([dst, data_index, end, pixel_at] {  // writeChunk, writer.zig:58-66
            size_t out_idx = 0;
            for (size_t i = data_index; i < end; ++i) out_idx += formatPixel(dst + out_idx, pixel_at(i));
        });
        file_index += chunk_content_size;
    }
    w.thread_pool->waitAndWork();
    if (w.truncate_to_content) handle.shrink(file_index);
    return file_index;
}

}  // namespace

size_t WriterPPM::write(const std::string& out_path, const Real* data, size_t lanes, size_t num_cols, size_t num_rows) const {
    const size_t n = num_cols * num_rows;
    return writeImpl(*this, out_path, n, num_cols, num_rows, [data, lanes](size_t i) { return encodeColor(data + i * lanes); });
}

size_t WriterPPM::writeQuantised(const std::string& out_path, const uint8_t* rgb, size_t num_cols, size_t num_rows) const {
    const size_t n = num_cols * num_rows;
    return writeImpl(*this, out_path, n, num_cols, num_rows, [rgb](size_t i) {
        return std::array<uint8_t, 3>{rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]};
    });
}

size_t WriterPPM::writeOnDevice(wrt_ctx* ctx, const std::string& out_path, const uint8_t* rgb, size_t num_cols, size_t num_rows) const {
    char header[64];
    const int header_len = std::snprintf(header, sizeof header, "P3\n%zu %zu\n255\n", num_cols, num_rows);
    const size_t file_size = num_cols * num_rows * kPixelNumBytes + static_cast<size_t>(header_len);  // writer.zig:20
    MmapHandlePosix handle(out_path, file_size);
    uint64_t content = 0;
    if (wrt_format_ppm(ctx, rgb, static_cast<uint32_t>(num_cols), static_cast<uint32_t>(num_rows), handle.ptr, file_size, &content) != WRT_OK)
        throw std::runtime_error(std::string("wrt_format_ppm: ") + wrt_last_error(ctx));
    if (truncate_to_content) handle.shrink(static_cast<size_t>(content));
    return static_cast<size_t>(content);
}

}  // namespace wrh
