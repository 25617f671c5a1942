// partition.cu — RayPartitioner (Tracer/RayPartitioner.h:L63-256, .cu:L263-411).
//
//   MultiPartition : stable radix sort of (key, index) over the data bit range then the batch bit range
//                    (one sort when the ranges are contiguous), then the partition table {count, start
//                    offsets, first key of each partition}. The reference marks splits (KCFindSplits),
//                    compacts them with cub::DevicePartition::If and reads the table on the HOST before
//                    launching one kernel per partition; here the table stays in device memory (batch
//                    values are small integers, so a dense per-batch table + one block-wide scan replaces
//                    the compaction) and nothing synchronises.
//   BinaryPartition: stable two-way split of an index list by a flag (dead/alive compaction) = a one-bit
//                    pass of the same onesweep sort; the left count comes from the sort histogram.
#include "common.cuh"

namespace mrb
{
namespace
{

constexpr int PTPB = 256;

__global__ void __launch_bounds__(PTPB)
KFindSplits(const uint32_t* __restrict__ sortedKeys, uint32_t n, uint32_t batchShift, uint32_t batchMask,
            uint32_t* __restrict__ firstOfBatch)
{
    for(uint32_t i = blockIdx.x * PTPB + threadIdx.x; i < n; i += gridDim.x * PTPB)
    {
        const uint32_t b = (sortedKeys[i] >> batchShift) & batchMask;
        if(i == 0 || ((sortedKeys[i - 1] >> batchShift) & batchMask) != b) firstOfBatch[b] = i; // one writer per batch value
    }
}

// One block: compacts the dense table in ascending batch order (= ascending start offset).
__global__ void __launch_bounds__(1024)
KCompactSplits(const uint32_t* __restrict__ firstOfBatch, uint32_t tableSize, const uint32_t* __restrict__ sortedKeys,
               uint32_t n, uint32_t maxPartitions,
               uint32_t* __restrict__ outCount, uint32_t* __restrict__ outOffsets, uint32_t* __restrict__ outKeys)
{
    __shared__ uint32_t sWarp[32];
    __shared__ uint32_t sBase;
    if(threadIdx.x == 0) sBase = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for(uint32_t start = 0; start < tableSize; start += 1024u)
    {
        const uint32_t j = start + threadIdx.x;
        const uint32_t first = (j < tableSize) ? firstOfBatch[j] : INVALID_U32;
        const bool valid = first != INVALID_U32;
        const uint32_t bal = __ballot_sync(0xffffffffu, valid);
        if(lane == 0) sWarp[warp] = __popc(bal);
        __syncthreads();
        uint32_t warpBase = 0;
        for(uint32_t w = 0; w < warp; w++) warpBase += sWarp[w];
        uint32_t total = 0;
        for(uint32_t w = 0; w < 32u; w++) total += sWarp[w];
        const uint32_t slot = sBase + warpBase + __popc(bal & ((1u << lane) - 1u));
        if(valid && slot < maxPartitions) { outOffsets[slot] = first; outKeys[slot] = sortedKeys[first]; }
        __syncthreads();
        if(threadIdx.x == 0) sBase += total;
        __syncthreads();
    }
    if(threadIdx.x == 0)
    {
        const uint32_t c = min(sBase, maxPartitions);
        outCount[0] = c;
        outOffsets[c] = n;
    }
}

__global__ void __launch_bounds__(PTPB)
KFlagsToKeys(const uint32_t* __restrict__ indicesIn, const uint8_t* __restrict__ flags, uint32_t n,
             uint32_t* __restrict__ keys, uint32_t* __restrict__ indicesOut, uint32_t* __restrict__ leftCount)
{
    uint32_t local = 0;
    for(uint32_t i = blockIdx.x * PTPB + threadIdx.x; i < n; i += gridDim.x * PTPB)
    {
        const uint32_t idx = indicesIn[i];
        const uint32_t left = flags[idx] ? 1u : 0u; // predicate true -> left side
        keys[i] = left ? 0u : 1u;
        indicesOut[i] = idx;
        local += left;
    }
    for(int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if((threadIdx.x & 31u) == 0 && local) atomicAdd(leftCount, local);
}

} // namespace

size_t MultiPartitionTempBytes(uint32_t count, uint32_t batchBits)
{
    MultiAlloc ma(nullptr);
    ma.Take<char>(RadixSortTempBytes(count, 4));
    ma.Take<uint32_t>(size_t(1) << batchBits);
    return ma.Total();
}

// keys / indices are sorted in place; outputs live in device memory.
void MultiPartition(Context& ctx, uint32_t* keys, uint32_t* indices, uint32_t count,
                    const uint32_t dataBits[2], const uint32_t batchBits[2], bool onlySortForBatches,
                    uint32_t maxPartitions, uint32_t* outCount, uint32_t* outOffsets, uint32_t* outKeys, void* temp)
{
    const uint32_t nBatchBits = batchBits[1] - batchBits[0];
    MultiAlloc ma(temp);
    void* sortTemp = ma.Take<char>(RadixSortTempBytes(count, 4));
    uint32_t* table = ma.Take<uint32_t>(size_t(1) << nBatchBits);
    if(onlySortForBatches) RadixSortPairs(ctx, keys, indices, count, batchBits[0], batchBits[1], sortTemp);
    else if(dataBits[1] == batchBits[0]) RadixSortPairs(ctx, keys, indices, count, dataBits[0], batchBits[1], sortTemp);
    else
    {
        RadixSortPairs(ctx, keys, indices, count, dataBits[0], dataBits[1], sortTemp);
        RadixSortPairs(ctx, keys, indices, count, batchBits[0], batchBits[1], sortTemp);
    }
    MRB_CUDA_TRY(cudaMemsetAsync(table, 0xFF, sizeof(uint32_t) << nBatchBits, ctx.stream));
    if(count) MRB_LAUNCH(ctx, KFindSplits, GridFor(ctx, count, PTPB, 8), PTPB, 0, keys, count, batchBits[0], (1u << nBatchBits) - 1u, table);
    MRB_LAUNCH(ctx, KCompactSplits, 1, 1024, 0, table, 1u << nBatchBits, keys, count, maxPartitions, outCount, outOffsets, outKeys);
}

size_t BinaryPartitionTempBytes(uint32_t count)
{
    MultiAlloc ma(nullptr);
    ma.Take<char>(RadixSortTempBytes(count, 4));
    ma.Take<uint32_t>(count);
    return ma.Total();
}

// indicesOut = [indices whose flag is set, in order][the others, in order]; leftCount (device) = size of the first part.
void BinaryPartition(Context& ctx, uint32_t* indicesOut, uint32_t* leftCount, const uint32_t* indicesIn,
                     const uint8_t* flags, uint32_t count, void* temp)
{
    MultiAlloc ma(temp);
    void* sortTemp = ma.Take<char>(RadixSortTempBytes(count, 4));
    uint32_t* keys = ma.Take<uint32_t>(count);
    MRB_CUDA_TRY(cudaMemsetAsync(leftCount, 0, sizeof(uint32_t), ctx.stream));
    if(count == 0) return;
    MRB_LAUNCH(ctx, KFlagsToKeys, GridFor(ctx, count, PTPB, 8), PTPB, 0, indicesIn, flags, count, keys, indicesOut, leftCount);
    RadixSortPairs(ctx, keys, indicesOut, count, 0, 1, sortTemp);
}

} // namespace mrb
