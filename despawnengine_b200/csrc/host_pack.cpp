// Host-side chunk packer + its worker pool (see host_pack.hpp).  Plain C++: compiled by the host compiler, no CUDA in here.
#include "host_pack.hpp"

#include <cstring>
#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define DSP_HOST_X86 1
#endif

namespace dsp_host {

namespace {

inline uint32_t classify(uint32_t any_bits, uint32_t differs, uint32_t first) {
    if (any_bits == 0u) return kPackAir;
    if (differs == 0u) return kPackUniform | (first << 16);
    if (any_bits < 4u) return kPack2;
    if (any_bits < 16u) return kPack4;
    if (any_bits < 256u) return kPack8;
    return kPackRaw;
}

#ifdef DSP_HOST_X86
// 32 x u16 (two registers, values < 256) -> 32 bytes in voxel order
__attribute__((target("avx2"))) inline __m256i narrow16(__m256i a, __m256i b) {
    return _mm256_permute4x64_epi64(_mm256_packus_epi16(a, b), 0xD8);
}
// 64 bytes (values < 16, voxel order) -> 32 bytes, byte j = v[2j] | v[2j+1] << 4
__attribute__((target("avx2"))) inline __m256i nibbles(__m256i a, __m256i b) {
    const __m256i k = _mm256_set1_epi16(0x1001);  // per byte pair: lo * 1 + hi * 16
    return narrow16(_mm256_maddubs_epi16(a, k), _mm256_maddubs_epi16(b, k));
}
// 64 nibble-bytes (v0 | v1 << 4, values < 4) -> 32 bytes, byte j = v0 | v1 << 2 | v2 << 4 | v3 << 6
__attribute__((target("avx2"))) inline __m256i crumbs(__m256i a, __m256i b) {
    const __m256i lo = _mm256_set1_epi8(0x03), hi = _mm256_set1_epi8(0x0C);
    const __m256i a4 = _mm256_or_si256(_mm256_and_si256(a, lo), _mm256_and_si256(_mm256_srli_epi16(a, 2), hi));
    const __m256i b4 = _mm256_or_si256(_mm256_and_si256(b, lo), _mm256_and_si256(_mm256_srli_epi16(b, 2), hi));
    return nibbles(a4, b4);
}

// OR of the sixteen u16 lanes
__attribute__((target("avx2"))) inline uint32_t fold_or16(__m256i x) {
    __m128i r = _mm_or_si128(_mm256_castsi256_si128(x), _mm256_extracti128_si256(x, 1));
    r = _mm_or_si128(r, _mm_srli_si128(r, 8));
    r = _mm_or_si128(r, _mm_srli_si128(r, 4));
    r = _mm_or_si128(r, _mm_srli_si128(r, 2));
    return static_cast<uint32_t>(_mm_cvtsi128_si32(r)) & 0xFFFFu;
}

__attribute__((target("avx2"))) uint32_t pack_chunk_avx2(const uint16_t* v, uint8_t* dst) {
    const __m256i* src = reinterpret_cast<const __m256i*>(v);
    const __m256i first = _mm256_set1_epi16(static_cast<short>(v[0]));
    __m256i any = _mm256_setzero_si256(), diff = _mm256_setzero_si256();
    for (int i = 0; i < 256; i += 4) {
        const __m256i a = _mm256_loadu_si256(src + i), b = _mm256_loadu_si256(src + i + 1), c = _mm256_loadu_si256(src + i + 2), d = _mm256_loadu_si256(src + i + 3);
        any = _mm256_or_si256(any, _mm256_or_si256(_mm256_or_si256(a, b), _mm256_or_si256(c, d)));
        diff = _mm256_or_si256(diff, _mm256_or_si256(_mm256_or_si256(_mm256_xor_si256(a, first), _mm256_xor_si256(b, first)),
                                                     _mm256_or_si256(_mm256_xor_si256(c, first), _mm256_xor_si256(d, first))));
    }
    const uint32_t desc = classify(fold_or16(any), fold_or16(diff), v[0]);
    const uint32_t form = desc & 7u;
    __m256i* out = reinterpret_cast<__m256i*>(dst);
    if (form == kPackRaw) {
        std::memcpy(dst, v, kPackSlotBytes);
    } else if (form == kPack8) {
        for (int i = 0; i < 256; i += 2) _mm256_storeu_si256(out + (i >> 1), narrow16(_mm256_loadu_si256(src + i), _mm256_loadu_si256(src + i + 1)));
    } else if (form == kPack4) {
        for (int i = 0; i < 256; i += 4)
            _mm256_storeu_si256(out + (i >> 2), nibbles(narrow16(_mm256_loadu_si256(src + i), _mm256_loadu_si256(src + i + 1)),
                                                        narrow16(_mm256_loadu_si256(src + i + 2), _mm256_loadu_si256(src + i + 3))));
    } else if (form == kPack2) {
        for (int i = 0; i < 256; i += 8) {
            const __m256i n0 = nibbles(narrow16(_mm256_loadu_si256(src + i), _mm256_loadu_si256(src + i + 1)),
                                       narrow16(_mm256_loadu_si256(src + i + 2), _mm256_loadu_si256(src + i + 3)));
            const __m256i n1 = nibbles(narrow16(_mm256_loadu_si256(src + i + 4), _mm256_loadu_si256(src + i + 5)),
                                       narrow16(_mm256_loadu_si256(src + i + 6), _mm256_loadu_si256(src + i + 7)));
            _mm256_storeu_si256(out + (i >> 3), crumbs(n0, n1));
        }
    }
    return desc;
}
#endif

}  // namespace

uint32_t pack_chunk_portable(const uint16_t* v, uint8_t* dst) {
    uint32_t any = 0, diff = 0;
    for (int i = 0; i < 4096; ++i) { any |= v[i]; diff |= static_cast<uint32_t>(v[i] ^ v[0]); }
    const uint32_t desc = classify(any, diff, v[0]);
    switch (desc & 7u) {
        case kPackRaw: std::memcpy(dst, v, kPackSlotBytes); break;
        case kPack8: for (int i = 0; i < 4096; ++i) dst[i] = static_cast<uint8_t>(v[i]); break;
        case kPack4: for (int i = 0; i < 2048; ++i) dst[i] = static_cast<uint8_t>(v[2 * i] | (v[2 * i + 1] << 4)); break;
        case kPack2: for (int i = 0; i < 1024; ++i) dst[i] = static_cast<uint8_t>(v[4 * i] | (v[4 * i + 1] << 2) | (v[4 * i + 2] << 4) | (v[4 * i + 3] << 6)); break;
        default: break;
    }
    return desc;
}

uint32_t pack_chunk(const uint16_t* v, uint8_t* dst) {
#ifdef DSP_HOST_X86
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) return pack_chunk_avx2(v, dst);
#endif
    return pack_chunk_portable(v, dst);
}

PackPool::PackPool(int workers) {
    for (int i = 0; i < workers; ++i) threads_.emplace_back([this] { run(); });
}

PackPool::~PackPool() {
    quit_.store(true);
    { std::lock_guard<std::mutex> g(m_); generation_.fetch_add(1); }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
}

void PackPool::start(const uint16_t* blocks, const uint32_t* palette_lens, uint64_t n, uint32_t* desc, uint8_t* staging, uint32_t keep_of_16) {
    const uint64_t nb = (n + kBlock - 1) / kBlock;
    if (nb > block_cap_) {
        block_cap_ = nb * 2;
        block_done_.reset(new std::atomic<uint8_t>[block_cap_]);
    }
    for (uint64_t b = 0; b < nb; ++b) block_done_[b].store(0, std::memory_order_relaxed);
    done_blocks_.store(0, std::memory_order_relaxed);
    upto_blocks_ = 0;
    {
        std::lock_guard<std::mutex> g(m_);
        job_.blocks = blocks; job_.lens = palette_lens; job_.desc = desc; job_.staging = staging; job_.n = n; job_.n_blocks = nb;
        job_.keep_of_16 = keep_of_16;
        job_.generation = generation_.load() + 1;
        next_.store(static_cast<uint64_t>(job_.generation) << 32, std::memory_order_release);
        generation_.store(job_.generation, std::memory_order_release);
    }
    cv_.notify_all();
}

void PackPool::work(const Job& job) {
    for (;;) {
        uint64_t v = next_.load(std::memory_order_acquire);
        uint64_t b;
        for (;;) {
            if (static_cast<uint32_t>(v >> 32) != job.generation) return;  // a newer job has started: not ours
            b = v & 0xFFFFFFFFull;
            if (b >= job.n_blocks) return;
            if (next_.compare_exchange_weak(v, v + 1, std::memory_order_acq_rel)) break;
        }
        const uint64_t i0 = b * kBlock, i1 = std::min(job.n, i0 + kBlock);
        for (uint64_t i = i0; i < i1; ++i) {
            // Chunk::palette.len() == 1: the chunk has never held anything but air (chunk.rs:58-65) -- not even read
            if (job.lens != nullptr && job.lens[i] == 1u) { job.desc[i] = kPackAir; continue; }
            if ((i & 15u) >= job.keep_of_16) { job.desc[i] = kPackInPlace; continue; }
            job.desc[i] = pack_chunk(job.blocks + i * 4096, job.staging + i * kPackSlotBytes);
        }
        block_done_[b].store(1, std::memory_order_release);
        done_blocks_.fetch_add(1, std::memory_order_acq_rel);
    }
}

void PackPool::run() {
    uint32_t seen = 0;
    for (;;) {
        // a short spin first (back-to-back calls arrive every few hundred microseconds), then sleep
        bool fresh = false;
        for (int spin = 0; spin < 20000 && !fresh; ++spin) {
            fresh = generation_.load(std::memory_order_acquire) != seen;
#ifdef DSP_HOST_X86
            if (!fresh) _mm_pause();
#endif
        }
        Job job;
        {
            std::unique_lock<std::mutex> g(m_);
            cv_.wait(g, [&] { return generation_.load(std::memory_order_acquire) != seen; });
            seen = generation_.load(std::memory_order_acquire);
            if (quit_.load()) return;
            job = job_;
        }
        if (job.generation == seen) work(job);
    }
}

uint64_t PackPool::packed_upto() {
    while (upto_blocks_ < job_.n_blocks && block_done_[upto_blocks_].load(std::memory_order_acquire)) ++upto_blocks_;
    return std::min(job_.n, upto_blocks_ * kBlock);
}

void PackPool::help() { work(job_); }

void PackPool::finish() {
    work(job_);
    while (done_blocks_.load(std::memory_order_acquire) < job_.n_blocks) {
#ifdef DSP_HOST_X86
        _mm_pause();
#endif
    }
    upto_blocks_ = job_.n_blocks;
}

}  // namespace dsp_host
