// oracle/ref_arena.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Monotonic bump allocator behind the global operator new of the reference harness.
//
// Why: ORBextractor::DistributeOctTree sorts vector<pair<int, ExtractorNode*>>
// (/root/reference/src/ORBextractor.cc:711), so nodes with equal key counts are ordered
// by HEAP ADDRESS.  With glibc malloc the retained keypoint set depends on allocator
// history.  With a bump allocator addresses grow in creation
// order, the function becomes pure, and it equals the rule "among equal sizes the most
// recently created node is split first", which is what the oracle restatement and the
// CUDA path implement.
//
// Each thread owns one lazily mmap'ed region (MAP_NORESERVE).  Harness entry points
// open an ArenaScope: the bump pointer is rewound on exit, after every object that
// was allocated inside the call has been destroyed (cv::Mat buffers are malloc'ed by
// the shim and never live here).  operator delete is a no-op for arena pointers.
#ifndef ORACLE_REF_ARENA_HPP
#define ORACLE_REF_ARENA_HPP

#include <cstddef>

namespace ref_arena {
void* alloc(std::size_t n);
bool owns(const void* p);
std::size_t mark();
void rewind(std::size_t m);
void reset_parked();   // rewind the regions of finished threads; only while nothing they allocated is alive
struct Scope {
    std::size_t m;
    Scope() : m(mark()) {}
    ~Scope() { rewind(m); }
};
} // namespace ref_arena

#endif
