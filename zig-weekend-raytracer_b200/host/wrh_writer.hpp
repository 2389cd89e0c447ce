// wrh_writer.hpp — host-side mirror of the reference's PPM writer (src/writer/writer.zig, src/writer/mmap.zig) and of
// the std.Thread.Pool use around it (spawnWg / waitAndWork).  Stays on the host by design (BASELINE.json north_star:
// "multithreaded mmap PPM write"); the quantisation itself (encodeColor) can be taken from the device's fused final
// pass instead (wrt_encode_rgb8), which produces identical bytes.
#pragma once

#include <array>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#include "../../include/wrt.h"
#include "wrh_math.hpp"

namespace wrh {

class ThreadPool {  // std.Thread.Pool with a WaitGroup folded in (main.zig:68-70, writer.zig:32-50)
   public:
    explicit ThreadPool(size_t n_jobs);
    ~ThreadPool();
    void spawnWg(std::function<void()> job);
    void waitAndWork();  // the caller helps drain the queue, then waits for running jobs
    size_t size() const { return workers_.size(); }

   private:
    bool runOne(std::unique_lock<std::mutex>& lk);
    std::vector<std::thread> workers_;
    std::queue<std::function<void()>> jobs_;
    std::mutex mu_;
    std::condition_variable cv_job_, cv_done_;
    size_t pending_ = 0;
    bool stop_ = false;
};

std::array<uint8_t, 3> encodeColor(const Real rgb[3]);   // writer.zig:68-94
size_t sizeOfLine(const std::array<uint8_t, 3>& pixel);  // writer.zig:96-100
size_t sizeOfDigit(uint8_t digit);                       // writer.zig:107-114

struct WriterPPM {  // writer.zig:6-52
    ThreadPool* thread_pool = nullptr;
    // The reference sizes the file for 12 bytes per pixel and never shrinks it, so real outputs end in NUL bytes
    // (writer.zig:20-23, mmap.zig:15-16).  Kept by default; set to cut the file at the last written byte.
    bool truncate_to_content = false;

    // `data`: num_rows * num_cols pixels, `lanes` doubles apiece (linear radiance in lanes 0..2).
    // Returns the number of content bytes (header + pixel lines).  Throws std::runtime_error on I/O failure.
    size_t write(const std::string& out_path, const Real* data, size_t lanes, size_t num_cols, size_t num_rows) const;
    // Same file from an already quantised frame (3 bytes per pixel, e.g. the device's fused final pass).
    size_t writeQuantised(const std::string& out_path, const uint8_t* rgb, size_t num_cols, size_t num_rows) const;
    // Same file, formatted by the back end (wrt_format_ppm) straight into the mapping: the frame of the last render
    // (rgb == nullptr) or the given quantised frame.  No host threads, no serial size pre-pass.
    size_t writeOnDevice(wrt_ctx* ctx, const std::string& out_path, const uint8_t* rgb, size_t num_cols, size_t num_rows) const;
};

}  // namespace wrh
