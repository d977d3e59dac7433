// codec.cu -- the reference's on-disk cube codec (server/src/chunk/codec.rs) on the GPU:
// encode_cubes' run-length records and the record loop of CubeGrid::decode, bit for bit including the
// codec's quirks (the leading empty record, air and the first material sharing wire value 0).
//
// A chunk's 32 768 ids in voxel-index order are cut into 256 segments of 128 ids, one per thread.  A run is
// owned by the thread in whose segment it starts; where it ends is found with a block-wide suffix-min over the
// segments' first run starts, so no thread walks ids outside its segment.  Records per run = ceil(len / 255).
#include "host_common.cuh"

using namespace hb;

namespace {

using u64 = unsigned long long;
constexpr int kCT = 256;
constexpr int kSeg = HB_CHUNK_VOLUME / kCT;  // 128 ids per thread
constexpr int kStageBytes = 16384;

// PaletteMaterialId::to_u16 (material.rs:224-226) / None -> 0 (codec.rs:85-90)
__device__ __forceinline__ uint32_t wire_of_raw(uint32_t raw) { return raw ? raw - 1u : 0u; }

__device__ __forceinline__ void put_record(uint8_t* dst, uint32_t count, uint32_t wire) {
    dst[0] = (uint8_t)count;
    dst[1] = (uint8_t)(wire & 0xFFu);
    dst[2] = (uint8_t)(wire >> 8);
}

// EMIT = false: counts[chunk] = number of records.  EMIT = true: records written at 3 * offsets[chunk].
template <bool EMIT>
__global__ void __launch_bounds__(kCT) k_rle_encode(const uint16_t* __restrict__ ids, size_t n, uint32_t* __restrict__ counts,
                                                    const u64* __restrict__ offsets, uint8_t* __restrict__ out,
                                                    u64 capacity_records, u64* __restrict__ total) {
    __shared__ int s_wmin[kCT / 32];
    __shared__ uint32_t s_wsum[kCT / 32];
    // EMIT: a chunk's records are assembled here and leave as coalesced 32-bit stores (terrain chunks have about
    // 1 100 records = 3.4 KB); a chunk with more than kStageBytes of records falls back to direct byte stores
    __shared__ uint32_t s_stage[EMIT ? kStageBytes / 4 : 1];
    __shared__ unsigned long long s_next;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // chunks are drawn from a ticket counter (total[3], zeroed by the launch wrapper): their cost follows the number of
    // runs, and a static stride leaves a tail of late CTAs
    for (size_t chunk = blockIdx.x; chunk < n;) {
        const uint16_t* src = ids + chunk * HB_CHUNK_VOLUME + tid * kSeg;
        // run-start mask of the segment: bit i set iff ids[i] != ids[i-1]; position 0 always starts a run (the
        // encoder's initial (None, 0) state either merges with a leading None run or is closed as an empty record)
        uint32_t b[4];
        uint32_t prev = tid ? (uint32_t)src[-1] : ~0u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t m = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + q * 4 + j);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t lo = w[k] & 0xFFFFu, hi = w[k] >> 16;
                    m |= (uint32_t)(lo != prev) << (j * 8 + k * 2);
                    m |= (uint32_t)(hi != lo) << (j * 8 + k * 2 + 1);
                    prev = hi;
                }
            }
            b[q] = m;
        }
        const int seg0 = tid * kSeg;
        int first = HB_CHUNK_VOLUME;
#pragma unroll
        for (int q = 3; q >= 0; --q)
            if (b[q]) first = seg0 + q * 32 + __ffs(b[q]) - 1;
        // first run start after this segment: exclusive suffix-min over `first`
        int incl = first;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_down_sync(0xffffffffu, incl, d);
            if (lane + d < 32) incl = min(incl, o);
        }
        int after = __shfl_down_sync(0xffffffffu, incl, 1);
        if (lane == 31) after = HB_CHUNK_VOLUME;
        __syncthreads();  // previous chunk's readers of the shared arrays are done
        if (lane == 0) s_wmin[warp] = incl;
        if (tid == 0) s_next = (unsigned long long)gridDim.x + atomicAdd(total + 3, 1ULL);
        __syncthreads();
        const size_t this_chunk = chunk;
        chunk = (size_t)s_next;  // the loop's `continue`s below move on to the next ticket
        for (int w2 = warp + 1; w2 < kCT / 32; ++w2) after = min(after, s_wmin[w2]);

        // records owned by this thread
        const uint32_t id0 = tid == 0 ? (uint32_t)__ldg(src) : 0u;
        uint32_t recs = (tid == 0 && id0 != 0) ? 1u : 0u;  // the empty leading record (codec.rs:78-91 on the first voxel)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t m = b[q];
            while (m) {
                const int bit = __ffs(m) - 1;
                m &= m - 1;
                const int p = seg0 + q * 32 + bit;
                int next = after;
                if (m) next = seg0 + q * 32 + __ffs(m) - 1;
                else {
#pragma unroll
                    for (int q2 = 3; q2 > q; --q2)
                        if (q2 > q && b[q2]) next = seg0 + q2 * 32 + __ffs(b[q2]) - 1;
                }
                recs += (uint32_t)(next - p + 254) / 255u;
            }
        }
        // block exclusive scan of recs
        uint32_t sc = recs;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, sc, d);
            if (lane >= d) sc += o;
        }
        if (lane == 31) s_wsum[warp] = sc;
        __syncthreads();
        uint32_t before = sc - recs, all = 0;
#pragma unroll
        for (int w2 = 0; w2 < kCT / 32; ++w2) {
            if (w2 < warp) before += s_wsum[w2];
            all += s_wsum[w2];
        }
        if (!EMIT) {
            if (tid == 0) counts[this_chunk] = all;
            continue;
        }
        const u64 off = offsets[this_chunk];
        if (off + (((u64)all + 3ULL) & ~3ULL) > capacity_records) {
            if (tid == 0) atomicOr(total + 2, 1ULL);  // error path only
            continue;
        }
        const bool staged = all * 3u <= (uint32_t)kStageBytes;  // block-uniform
        if (staged) {
            if (tid == 0 && ((all * 3u) & 3u)) s_stage[(all * 3u) >> 2] = 0u;  // tail of the last word: the zero-filled gap
            __syncthreads();
        }
        uint8_t* dst = staged ? reinterpret_cast<uint8_t*>(s_stage) + before * 3u : out + (off + before) * 3ULL;
        if (tid == 0 && id0 != 0) { put_record(dst, 0u, 0u); dst += 3; }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t m = b[q];
            while (m) {
                const int bit = __ffs(m) - 1;
                m &= m - 1;
                const int p = seg0 + q * 32 + bit;
                int next = after;
                if (m) next = seg0 + q * 32 + __ffs(m) - 1;
                else {
#pragma unroll
                    for (int q2 = 3; q2 > q; --q2)
                        if (q2 > q && b[q2]) next = seg0 + q2 * 32 + __ffs(b[q2]) - 1;
                }
                const uint32_t wire = wire_of_raw((uint32_t)__ldg(src + (p - seg0)));
                int len = next - p;
                while (len > 255) { put_record(dst, 255u, wire); dst += 3; len -= 255; }
                put_record(dst, (uint32_t)len, wire);
                dst += 3;
            }
        }
        if (staged) {
            __syncthreads();
            // stream starts are multiples of 4 records = 12 bytes, so the words are aligned in the output too
            uint32_t* outw = reinterpret_cast<uint32_t*>(out + off * 3ULL);
            const uint32_t words = (all * 3u + 3u) >> 2;
            for (uint32_t i = tid; i < words; i += kCT) outw[i] = s_stage[i];
        }
    }
}

// CubeGrid::decode (codec.rs:52-66): thread per record, block scan of the counts gives each record's first voxel.
__global__ void __launch_bounds__(kCT) k_rle_decode(const uint8_t* __restrict__ bytes, const u64* __restrict__ offsets,
                                                    const uint32_t* __restrict__ sizes, size_t n, uint16_t* __restrict__ ids,
                                                    u64* __restrict__ total) {
    __shared__ uint32_t s_wsum[kCT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (size_t chunk = blockIdx.x; chunk < n; chunk += gridDim.x) {
        const uint8_t* src = bytes + offsets[chunk];
        const uint32_t nrec = sizes[chunk] / 3u;  // array_chunks::<3>() drops a partial tail
        uint16_t* dst = ids + chunk * HB_CHUNK_VOLUME;
        uint32_t running = 0;
        for (uint32_t base = 0; base < nrec; base += kCT) {
            const uint32_t r = base + tid;
            uint32_t cnt = 0, raw = 0;
            if (r < nrec) {
                cnt = src[3 * r];
                const uint32_t v = (uint32_t)src[3 * r + 1] | ((uint32_t)src[3 * r + 2] << 8);
                raw = (v + 1u) & 0xFFFFu;  // PaletteMaterialId::new(v) = NonZeroU16::new(v + 1), wrapping as a release build does
            }
            uint32_t sc = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, sc, d);
                if (lane >= d) sc += o;
            }
            __syncthreads();
            if (lane == 31) s_wsum[warp] = sc;
            __syncthreads();
            uint32_t pos = running + sc - cnt, all = 0;
#pragma unroll
            for (int w2 = 0; w2 < kCT / 32; ++w2) {
                if (w2 < warp) pos += s_wsum[w2];
                all += s_wsum[w2];
            }
            for (uint32_t k = 0; k < cnt; ++k)
                if (pos + k < (uint32_t)HB_CHUNK_VOLUME) dst[pos + k] = (uint16_t)raw;
            running += all;
            // block-uniform; one pass adds at most kCT * 255, so `running` can never wrap however long the stream is
            if (running > (uint32_t)HB_CHUNK_VOLUME) break;
        }
        if (running > (uint32_t)HB_CHUNK_VOLUME) {
            if (tid == 0) atomicOr(total + 2, 4ULL);  // the reference indexes past the chunk and panics
            running = HB_CHUNK_VOLUME;
        }
        for (uint32_t L = running + tid; L < (uint32_t)HB_CHUNK_VOLUME; L += kCT) dst[L] = 0;  // untouched voxels stay None
    }
}

}  // namespace

namespace hb {

cudaError_t launch_rle_encode(cudaStream_t s, bool emit, const uint16_t* d_ids, size_t n, uint32_t* d_counts,
                              const uint64_t* d_offsets, uint8_t* d_out, uint64_t capacity_records, uint64_t* d_total) {
    if (n == 0) return cudaSuccess;
    size_t grid = (size_t)num_sms() * 8;
    if (grid > n) grid = n;
    const cudaError_t e0 = cudaMemsetAsync(d_total + 3, 0, sizeof(uint64_t), s);  // the chunk ticket counter
    if (e0 != cudaSuccess) return e0;
    if (emit) k_rle_encode<true><<<(unsigned)grid, kCT, 0, s>>>(d_ids, n, d_counts, (const u64*)d_offsets, d_out, capacity_records, (u64*)d_total);
    else k_rle_encode<false><<<(unsigned)grid, kCT, 0, s>>>(d_ids, n, d_counts, nullptr, nullptr, 0, (u64*)d_total);
    return cudaGetLastError();
}

cudaError_t launch_rle_decode(cudaStream_t s, const uint8_t* d_bytes, const uint64_t* d_offsets, const uint32_t* d_sizes, size_t n,
                              uint16_t* d_ids, uint64_t* d_total) {
    if (n == 0) return cudaSuccess;
    size_t grid = (size_t)num_sms() * 8;
    if (grid > n) grid = n;
    k_rle_decode<<<(unsigned)grid, kCT, 0, s>>>(d_bytes, (const u64*)d_offsets, d_sizes, n, d_ids, (u64*)d_total);
    return cudaGetLastError();
}

}  // namespace hb

extern "C" {

int hb_dev_rle_encode(hb_ctx* c, void* stream, const uint16_t* d_ids, size_t n, void* d_scan_workspace, uint8_t* d_out,
                      uint64_t capacity_records, uint64_t* d_offsets, uint32_t* d_counts, uint64_t* d_total) {
    if (!c || !d_total) return fail(HB_ERR_INVALID, "null argument");
    if (n && (!d_ids || !d_scan_workspace || !d_offsets || !d_counts)) return fail(HB_ERR_INVALID, "null argument");
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    HB_CUDA(cudaMemsetAsync(d_total, 0, 4 * sizeof(uint64_t), s));
    if (n == 0) return HB_OK;
    HB_CUDA(launch_rle_encode(s, false, d_ids, n, d_counts, nullptr, nullptr, 0, d_total));
    HB_CUDA(launch_scan(s, d_counts, n, d_offsets, d_scan_workspace, d_total, false));
    if (d_out) HB_CUDA(launch_rle_encode(s, true, d_ids, n, d_counts, d_offsets, d_out, capacity_records, d_total));
    return HB_OK;
}

int hb_rle_encode(hb_ctx* c, const uint16_t* ids, size_t n, uint8_t* out, uint64_t out_capacity, uint64_t* out_offsets,
                  uint32_t* out_sizes, uint64_t* out_total) {
    if (!c || !out_total) return fail(HB_ERR_INVALID, "null argument");
    *out_total = 0;
    if (n == 0) return HB_OK;
    if (!ids || !out_offsets || !out_sizes) return fail(HB_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(c->mu);
    HB_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    HB_CUDA(c->d_ids.reserve(n * HB_CHUNK_VOLUME * sizeof(uint16_t)));
    HB_CUDA(c->d_counts.reserve(n * sizeof(uint32_t)));
    HB_CUDA(c->d_offsets.reserve(n * sizeof(uint64_t)));
    HB_CUDA(c->d_scan.reserve(scan_workspace_bytes(n)));
    HB_CUDA(c->d_total.reserve(4 * sizeof(uint64_t)));
    HB_CUDA(c->h_total.reserve(4 * sizeof(uint64_t)));
    HB_CUDA(cudaMemcpyAsync(c->d_ids.p, ids, n * HB_CHUNK_VOLUME * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
    int rc = hb_dev_rle_encode(c, s, c->d_ids.as<uint16_t>(), n, c->d_scan.p, nullptr, 0, c->d_offsets.as<uint64_t>(),
                               c->d_counts.as<uint32_t>(), c->d_total.as<uint64_t>());
    if (rc != HB_OK) return rc;
    HB_CUDA(cudaMemcpyAsync(c->h_total.p, c->d_total.p, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaMemcpyAsync(out_offsets, c->d_offsets.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaMemcpyAsync(out_sizes, c->d_counts.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaStreamSynchronize(s));
    uint64_t tot[4];
    memcpy(tot, c->h_total.p, sizeof(tot));
    const uint64_t span_bytes = tot[0] * 3ULL;  // record span incl. the padding that keeps each stream 12-byte aligned
    for (size_t i = 0; i < n; ++i) { out_offsets[i] *= 3ULL; out_sizes[i] *= 3u; }
    *out_total = span_bytes;
    if (span_bytes > out_capacity || !out)
        return fail(HB_ERR_CAPACITY, "encoded streams need %llu bytes, %llu given", (unsigned long long)span_bytes,
                    (unsigned long long)out_capacity);
    HB_CUDA(c->d_out.reserve(span_bytes));
    HB_CUDA(cudaMemsetAsync(c->d_out.p, 0, span_bytes, s));
    HB_CUDA(launch_rle_encode(s, true, c->d_ids.as<uint16_t>(), n, c->d_counts.as<uint32_t>(), c->d_offsets.as<uint64_t>(),
                              c->d_out.as<uint8_t>(), tot[0], c->d_total.as<uint64_t>()));
    HB_CUDA(cudaMemcpyAsync(out, c->d_out.p, span_bytes, cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaStreamSynchronize(s));
    return HB_OK;
}

int hb_rle_decode(hb_ctx* c, const uint8_t* bytes, uint64_t n_bytes, const uint64_t* offsets, const uint32_t* sizes, size_t n,
                  uint16_t* out_ids) {
    if (!c) return fail(HB_ERR_INVALID, "null argument");
    if (n == 0) return HB_OK;
    if (!offsets || !sizes || !out_ids || (n_bytes && !bytes)) return fail(HB_ERR_INVALID, "null argument");
    for (size_t i = 0; i < n; ++i)
        if (offsets[i] > n_bytes || sizes[i] > n_bytes - offsets[i])
            return fail(HB_ERR_INVALID, "stream %zu lies outside the %llu input bytes", i, (unsigned long long)n_bytes);
    std::lock_guard<std::mutex> lk(c->mu);
    HB_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    HB_CUDA(c->d_out.reserve(n_bytes + 16));
    HB_CUDA(c->d_offsets.reserve(n * sizeof(uint64_t)));
    HB_CUDA(c->d_counts.reserve(n * sizeof(uint32_t)));
    HB_CUDA(c->d_ids.reserve(n * HB_CHUNK_VOLUME * sizeof(uint16_t)));
    HB_CUDA(c->d_total.reserve(4 * sizeof(uint64_t)));
    HB_CUDA(c->h_total.reserve(4 * sizeof(uint64_t)));
    HB_CUDA(cudaMemsetAsync(c->d_total.p, 0, 4 * sizeof(uint64_t), s));
    if (n_bytes) HB_CUDA(cudaMemcpyAsync(c->d_out.p, bytes, n_bytes, cudaMemcpyHostToDevice, s));
    HB_CUDA(cudaMemcpyAsync(c->d_offsets.p, offsets, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    HB_CUDA(cudaMemcpyAsync(c->d_counts.p, sizes, n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    HB_CUDA(launch_rle_decode(s, c->d_out.as<uint8_t>(), c->d_offsets.as<uint64_t>(), c->d_counts.as<uint32_t>(), n,
                              c->d_ids.as<uint16_t>(), c->d_total.as<uint64_t>()));
    HB_CUDA(cudaMemcpyAsync(c->h_total.p, c->d_total.p, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaMemcpyAsync(out_ids, c->d_ids.p, n * HB_CHUNK_VOLUME * sizeof(uint16_t), cudaMemcpyDeviceToHost, s));
    HB_CUDA(cudaStreamSynchronize(s));
    uint64_t tot[4];
    memcpy(tot, c->h_total.p, sizeof(tot));
    if (tot[2] & 4) return fail(HB_ERR_INVALID, "a stream decodes to more than %d voxels (the reference panics here)", HB_CHUNK_VOLUME);
    return HB_OK;
}

}  // extern "C"
