// Host side of the packed upload of host-resident chunks (dsp_mesh_host_chunks / dsp_mesh_host_chunks_ex): a few worker threads
// turn each chunk's 4,096 u16 voxels (Chunk::blocks, /root/reference/src/content/world/chunks/chunk.rs:17) into the smallest of
// five lossless forms before anything crosses PCIe, and unpack_host_chunks_kernel expands them into the voxel store:
//   air      every voxel 0 (or Chunk::palette.len() == 1: not even read)        0 bytes
//   uniform  every voxel the same id                                            0 bytes (the id rides in the descriptor)
//   2 bit    all ids < 4     voxel i = (byte[i >> 2] >> 2 (i & 3)) & 3          1 KiB
//   4 bit    all ids < 16    voxel i = (byte[i >> 1] >> 4 (i & 1)) & 15         2 KiB
//   8 bit    all ids < 256   voxel i = byte[i]                                  4 KiB
//   raw      anything else: the 8 KiB as they are
// A terrain region is mostly air, solid rock and chunks of two or three block ids: 33.5 MB of voxels become about 1.3 MB on the
// bus, and the call is bound by how fast the host cores can read the voxels (they are in host memory anyway) instead of by PCIe.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace dsp_host {

constexpr uint32_t kPackAir = 0, kPackUniform = 1, kPack2 = 2, kPack4 = 3, kPack8 = 4, kPackRaw = 5;
constexpr uint32_t kPackInPlace = 6;  // not touched by the host at all: the device reads the chunk's 8 KiB where the caller holds them
constexpr uint32_t kPackSlotBytes = 8192;  // staging stride per chunk (the raw form fills it)

// descriptor: bits 0-2 form, bits 16-31 the block id of a uniform chunk
uint32_t pack_chunk(const uint16_t* voxels, uint8_t* dst);          // dispatches to AVX2 where the CPU has it
uint32_t pack_chunk_portable(const uint16_t* voxels, uint8_t* dst);  // plain C++ (also what the AVX2 path is tested against)

// One job at a time: chunks [0, n) of `blocks` -> descriptors + staging slots, handed out to the workers in blocks of 16 chunks in
// ascending order.  The caller watches `packed_upto()` to launch the pieces of its pipeline as their chunks become ready.
class PackPool {
public:
    explicit PackPool(int workers);
    ~PackPool();
    PackPool(const PackPool&) = delete;
    PackPool& operator=(const PackPool&) = delete;
    int workers() const { return static_cast<int>(threads_.size()); }
    // palette_lens may be NULL; desc[n], staging[n * kPackSlotBytes].  keep_of_16 < 16: of every 16 consecutive chunks only the first
    // keep_of_16 are packed, the others are left where they are (kPackInPlace) for the device to fetch -- the share of the work PCIe
    // takes off the host cores when there are few of them (the caller's array must be device-addressable for that).
    void start(const uint16_t* blocks, const uint32_t* palette_lens, uint64_t n, uint32_t* desc, uint8_t* staging, uint32_t keep_of_16 = 16);
    // chunks [0, returned) are packed (monotonic; == n once the job is complete)
    uint64_t packed_upto();
    void help();   // the calling thread packs blocks too until none is left to hand out
    void finish(); // returns when every chunk of the job is packed (call before touching / freeing the job's buffers)

private:
    static constexpr uint64_t kBlock = 16;
    struct Job {
        const uint16_t* blocks = nullptr;
        const uint32_t* lens = nullptr;
        uint32_t* desc = nullptr;
        uint8_t* staging = nullptr;
        uint64_t n = 0, n_blocks = 0;
        uint32_t generation = 0, keep_of_16 = 16;
    };
    void run();
    void work(const Job& job);
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_;
    std::atomic<uint32_t> generation_{0};
    std::atomic<bool> quit_{false};
    Job job_;                                  // written by start() under m_, copied by a worker under m_
    // (generation << 32) | next block to hand out: a worker that woke up late for an older job can never take a block of the current one
    std::atomic<uint64_t> next_{0};
    std::atomic<uint64_t> done_blocks_{0};     // blocks of the current job completed
    std::unique_ptr<std::atomic<uint8_t>[]> block_done_;
    uint64_t block_cap_ = 0;
    uint64_t upto_blocks_ = 0;                 // (caller's thread only) blocks [0, upto_blocks_) are known complete
};

}  // namespace dsp_host
