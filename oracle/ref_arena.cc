// oracle/ref_arena.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see ref_arena.hpp).
#include "ref_arena.hpp"

#include <sys/mman.h>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <new>

namespace {
const std::size_t kRegionBytes = std::size_t(8) << 30; // virtual, per thread
const int kMaxRegions = 1024;

std::atomic<char*> g_regions[kMaxRegions]; // slot == 0: free
std::atomic<int> g_hi(0);                  // slots >= g_hi were never used

// Regions are never unmapped: data allocated by a worker thread (e.g. the keypoint vectors that
// Frame::ExtractORB fills on its own std::thread, src/Frame.cc:82-85) outlives the thread.  A
// finished thread's region is parked and handed to the next new thread.
// A parked region keeps its bump offset: with several Frames in flight (orbref_stereo_bench) another worker's new
// thread may adopt it while the Frame that owns the data is still alive, so it must continue, not restart.
// reset_parked() rewinds all parked regions; harness entry points call it when nothing allocated is alive.
std::atomic<char*> g_parked[kMaxRegions];
std::atomic<std::size_t> g_parked_off[kMaxRegions];

struct ThreadArena {
    char* base;
    std::size_t off;
    ThreadArena() : base(0), off(0) {}
    void init()
    {
        for (int i = 0, n = g_hi.load(); i < n; ++i) {
            const std::size_t o = g_parked_off[i].load();
            char* p = g_parked[i].exchange(0);
            if (p) { base = p; off = o; return; }
        }
        void* p = mmap(0, kRegionBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { std::fprintf(stderr, "ref_arena: mmap failed\n"); std::abort(); }
        base = (char*)p;
        for (int i = 0; i < kMaxRegions; ++i) {
            char* expect = 0;
            if (g_regions[i].compare_exchange_strong(expect, base)) {
                int hi = g_hi.load();
                while (hi < i + 1 && !g_hi.compare_exchange_weak(hi, i + 1)) {}
                return;
            }
        }
        std::fprintf(stderr, "ref_arena: too many threads\n");
        std::abort();
    }
    ~ThreadArena()
    {
        if (!base) return;
        for (int i = 0; i < kMaxRegions; ++i) {
            char* expect = 0;
            // the slot's offset is written before the region becomes visible (a slot is owned by whoever emptied it)
            if (g_parked[i].load() == 0) {
                g_parked_off[i].store(off);
                if (g_parked[i].compare_exchange_strong(expect, base)) break;
            }
        }
        base = 0;
    }
};
thread_local ThreadArena t_arena;
} // namespace

namespace ref_arena {
void* alloc(std::size_t n)
{
    ThreadArena& a = t_arena;
    if (!a.base) a.init();
    std::size_t need = (n + 15) & ~std::size_t(15);
    if (a.off + need > kRegionBytes) { std::fprintf(stderr, "ref_arena: region exhausted\n"); std::abort(); }
    void* p = a.base + a.off;
    a.off += need;
    return p;
}
bool owns(const void* p)
{
    for (int i = 0, n = g_hi.load(std::memory_order_relaxed); i < n; ++i) {
        const char* b = g_regions[i].load(std::memory_order_relaxed);
        if (b && (const char*)p >= b && (const char*)p < b + kRegionBytes) return true;
    }
    return false;
}
void reset_parked()
{
    for (int i = 0; i < kMaxRegions; ++i) g_parked_off[i].store(0);
}
std::size_t mark() { return t_arena.off; }
void rewind(std::size_t m) { t_arena.off = m; }
} // namespace ref_arena

// Replace the global allocation functions for this shared object only (the link uses a
// version script that keeps these symbols local, so nothing else in the process sees them).
void* operator new(std::size_t n) { return ref_arena::alloc(n ? n : 1); }
void* operator new[](std::size_t n) { return ref_arena::alloc(n ? n : 1); }
void* operator new(std::size_t n, const std::nothrow_t&) noexcept { return ref_arena::alloc(n ? n : 1); }
void* operator new[](std::size_t n, const std::nothrow_t&) noexcept { return ref_arena::alloc(n ? n : 1); }
static inline void arena_delete(void* p) noexcept
{
    if (p && !ref_arena::owns(p)) std::free(p);
}
void operator delete(void* p) noexcept { arena_delete(p); }
void operator delete[](void* p) noexcept { arena_delete(p); }
void operator delete(void* p, std::size_t) noexcept { arena_delete(p); }
void operator delete[](void* p, std::size_t) noexcept { arena_delete(p); }
void operator delete(void* p, const std::nothrow_t&) noexcept { arena_delete(p); }
void operator delete[](void* p, const std::nothrow_t&) noexcept { arena_delete(p); }
