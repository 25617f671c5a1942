// common.cuh — shared declarations of the sm_100a hot path (context, arena, launch helpers).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include <string>
#include <vector>
#include "../../include/mray_b200.h"

namespace mrb
{

static constexpr uint32_t INVALID_U32 = 0xFFFFFFFFu;
static constexpr uint32_t LEAF_FLAG = 0x80000000u; // ChildIndex = KeyT<u32,1,31> (AcceleratorLBVH.h:L46-49)
static constexpr size_t   ARENA_ALIGN = 256;       // MemAlloc::DefaultSystemAlignment (Core/MemAlloc.h)

struct CudaError { cudaError_t code; const char* file; int line; };

#define MRB_CUDA_TRY(expr)                                                                    \
    do { cudaError_t e__ = (expr);                                                            \
         if(e__ != cudaSuccess) throw ::mrb::CudaError{e__, __FILE__, __LINE__}; } while(0)

inline size_t AlignUp(size_t v, size_t a = ARENA_ALIGN) { return (v + a - 1) / a * a; }

// One device allocation carved into 256-byte aligned sub-arrays: the B200 counterpart of the
// reference's DeviceMemory + MemAlloc::AllocateMultiData (Core/MemAlloc.h:L171-209).
class DeviceBlock
{
    void*  base = nullptr;
    size_t capacity = 0;
    public:
    DeviceBlock() = default;
    DeviceBlock(const DeviceBlock&) = delete;
    DeviceBlock& operator=(const DeviceBlock&) = delete;
    ~DeviceBlock() { Free(); }
    void Free() { if(base) cudaFree(base); base = nullptr; capacity = 0; }
    // Grow-only reservation (contents are NOT preserved on growth).
    void Reserve(size_t bytes)
    {
        if(bytes <= capacity) return;
        Free();
        size_t want = AlignUp(bytes + bytes / 8, 1u << 20);
        MRB_CUDA_TRY(cudaMalloc(&base, want));
        capacity = want;
    }
    void*  Base() const { return base; }
    size_t Capacity() const { return capacity; }
};

// Bump sub-allocator over a DeviceBlock. First pass with base == nullptr sizes the block.
class MultiAlloc
{
    char*  base;
    size_t offset = 0;
    public:
    explicit MultiAlloc(void* b = nullptr) : base(static_cast<char*>(b)) {}
    template<class T> T* Take(size_t count)
    {
        size_t at = offset;
        offset = AlignUp(offset + count * sizeof(T));
        return base ? reinterpret_cast<T*>(base + at) : nullptr;
    }
    size_t Total() const { return offset; }
};

// Sampled kernel timing (mrb_context_set_profiling): pairs of CUDA events recorded on the context stream around the
// kernels of every `stride`-th wavefront iteration, read back by mrb_context_get_profile. Off by default.
enum ProfileKind : int { PROF_TRACE_CLOSEST = 0, PROF_SHADE = 1, PROF_TRACE_ANY = 2, PROF_FINISH_RELOAD = 3, PROF_TRACE_TAIL = 4, PROF_KINDS = 5 };
struct KernelProfile
{
    static constexpr int POOL = 320;
    bool        enabled = false, sampleNow = false;
    uint32_t    stride = 16;
    cudaEvent_t ev[POOL][2] = {};
    int         kind[POOL] = {};
    int         used = 0, created = 0;
    double      ms[PROF_KINDS] = {};
    uint64_t    samples[PROF_KINDS] = {};
};

struct Context
{
    int          device = 0;
    cudaStream_t ownStream = nullptr;
    cudaStream_t stream = nullptr;
    int          smCount = 148;
    size_t       totalMem = 0;
    size_t       persistentBytes = 0; // bytes held by live accelerators / renderers
    DeviceBlock  scratch;             // per-call temporaries (grow-only, reused)
    DeviceBlock  traceScratch;        // exact-fallback ray list of the wide traversal
    const uint32_t* lastFallbackCount = nullptr; // device counter of the last wide cast
    uint64_t     launches = 0;
    uint32_t     alphaSeed = 0x9E3779B9u; // seed of the next cast's stochastic alpha test (advances per cast)
    std::string  error;
    cudaEvent_t  ev0 = nullptr, ev1 = nullptr;
    // host-pointer casts: upload / trace / download of consecutive ray chunks overlap on three streams
    static constexpr int PIPE_CHUNKS = 16;
    cudaStream_t copyIn = nullptr, copyOut = nullptr;
    cudaEvent_t  evIn[PIPE_CHUNKS] = {}, evDone[PIPE_CHUNKS] = {};
    cudaEvent_t  evStart = nullptr;
    KernelProfile prof;
    // occupancy of the persistent traversal kernels on THIS device (blocks per SM), filled on first use
    int          occWide[4] = {0, 0, 0, 0}, occWide2[4] = {0, 0, 0, 0};   // [anyHit + 2 * alphaMaps]
    const void*  persistBase = nullptr;   // accelerator the stream's L2 access-policy window currently covers
    cudaStream_t persistStream = nullptr;
};

// Brackets the launches issued inside its scope with a sampled event pair (no-op unless profiling samples this iteration)
struct ProfileScope
{
    Context& ctx; int slot = -1;
    ProfileScope(Context& c, int kind) : ctx(c)
    {
        KernelProfile& p = c.prof;
        if(!p.enabled || !p.sampleNow || p.used >= KernelProfile::POOL) return;
        if(p.used >= p.created)
        {
            if(cudaEventCreate(&p.ev[p.created][0]) != cudaSuccess || cudaEventCreate(&p.ev[p.created][1]) != cudaSuccess) return;
            p.created++;
        }
        slot = p.used++;
        p.kind[slot] = kind;
        cudaEventRecord(p.ev[slot][0], c.stream);
    }
    ~ProfileScope() { if(slot >= 0) cudaEventRecord(ctx.prof.ev[slot][1], ctx.stream); }
};

inline uint32_t DivUp(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// Grid sizing rule of the whole library: grid-stride kernels launch a multiple of the SM count
// (148 on B200), capped by the work available.
inline uint32_t GridFor(const Context& ctx, uint32_t work, uint32_t tpb, uint32_t blocksPerSM = 8)
{
    uint32_t need = DivUp(work, tpb);
    uint32_t cap = uint32_t(ctx.smCount) * blocksPerSM;
    return need < cap ? (need ? need : 1u) : cap;
}

#define MRB_LAUNCH(ctx, kernel, grid, block, smem, ...)                                        \
    do { kernel<<<(grid), (block), (smem), (ctx).stream>>>(__VA_ARGS__);                       \
         (ctx).launches++;                                                                     \
         MRB_CUDA_TRY(cudaGetLastError()); } while(0)

// ---- device algorithms (sort.cu) ------------------------------------------------------------
size_t RadixSortTempBytes(uint32_t count, size_t keyBytes);
// Sorts (keys, values) by bits [bitBegin, bitEnd); result lands back in keys/values.
// temp must hold RadixSortTempBytes(count, sizeof(K)).
void RadixSortPairs(Context& ctx, uint64_t* keys, uint32_t* values, uint32_t count,
                    uint32_t bitBegin, uint32_t bitEnd, void* temp);
void RadixSortPairs(Context& ctx, uint32_t* keys, uint32_t* values, uint32_t count,
                    uint32_t bitBegin, uint32_t bitEnd, void* temp);

} // namespace mrb
