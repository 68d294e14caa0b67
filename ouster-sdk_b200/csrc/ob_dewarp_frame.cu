// ob_dewarp_frame.cu -- K3: range image -> world-frame point list in one pass:
// LUT projection + per-column pose + range filter + order-preserving compaction.
//
// What it replaces (reference paths relative to /root/reference):
//   dewarp<T>(LidarFrame, XYZLutT<T>, min_range, max_range)  ouster_core/include/ouster/core/pose_util.h:456-485
//   impl::dewarp_impl (single frame)                         ouster_core/include/ouster/core/impl/dewarp_impl.h:22-76
// The reference projects the whole image (cartesian), then walks the columns between the first and
// the last valid one and appends the posed points that pass the range filter; its own note
// (dewarp_impl.h:27-29) asks for the projection to be folded in.  Here nothing is materialised:
//   count : one warp per (32 columns x 16 rows) block counts the surviving pixels of each column
//   scan  : one CTA finds the first/last valid column, masks excluded columns and turns the counts
//           into write offsets (column-major order, rows ascending inside a column -- the order
//           of the reference's loop)
//   emit  : the count mapping again; each lane projects its pixel from the LUT, applies the
//           column pose (cast from double to T like the reference) and writes at offset + rank.
#include <algorithm>

#include "ob_internal.h"

namespace ob {

constexpr int kSlabRows = 16;

__device__ __forceinline__ float k3_project(uint32_t r, float d, float o) {
    return r == 0 ? 0.0f : __fadd_rn(__fmul_rn(static_cast<float>(r), d), o);
}
__device__ __forceinline__ double k3_project(uint32_t r, double d, double o) {
    return r == 0 ? 0.0 : __dadd_rn(__dmul_rn(static_cast<double>(r), d), o);
}
__device__ __forceinline__ float k3_pose_row(const float* m, float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(m[0], x), __fadd_rn(__fmul_rn(m[1], y), __fmul_rn(m[2], z))), m[3]);
}
__device__ __forceinline__ double k3_pose_row(const double* m, double x, double y, double z) {
    return __dadd_rn(__dadd_rn(__dmul_rn(m[0], x), __dadd_rn(__dmul_rn(m[1], y), __dmul_rn(m[2], z))), m[3]);
}

// cnt[slab * W + col] = pixels of column col, rows of the slab, with min_r <= r <= max_r
__global__ void __launch_bounds__(256) k3_count_kernel(const uint32_t* __restrict__ range, unsigned H, unsigned W,
                                                       uint32_t min_r, uint32_t max_r, unsigned n_cg,
                                                       unsigned n_slabs, uint32_t* __restrict__ cnt) {
    const unsigned gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= n_cg * n_slabs) return;
    const unsigned cg = gw % n_cg, slab = gw / n_cg;
    const unsigned col = cg * 32u + (threadIdx.x & 31u);
    if (col >= W) return;
    const unsigned r0 = slab * kSlabRows, r1 = min(H, r0 + kSlabRows);
    uint32_t c = 0;
    for (unsigned row = r0; row < r1; ++row) {
        const uint32_t r = range[static_cast<size_t>(row) * W + col];
        c += (r >= min_r && r <= max_r) ? 1u : 0u;
    }
    cnt[static_cast<size_t>(slab) * W + col] = c;
}

// One CTA.  base[slab * W + col] = write offset of the first surviving pixel of (col, slab), or
// 0xffffffff for excluded columns; *total = number of points.
__global__ void __launch_bounds__(1024) k3_scan_kernel(const uint32_t* __restrict__ cnt,
                                                       const uint32_t* __restrict__ status, unsigned W,
                                                       unsigned n_slabs, uint32_t* __restrict__ base,
                                                       unsigned long long* __restrict__ total) {
    __shared__ int s_first, s_last;
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const unsigned tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) {
        s_first = 0x7fffffff;
        s_last = -1;
        s_carry = 0;
    }
    __syncthreads();
    // LidarFrame::get_first_valid_column / get_last_valid_column (lidar_frame.cpp:907-925)
    int lf = 0x7fffffff, ll = -1;
    for (unsigned c = tid; c < W; c += nt)
        if ((status[c] & 1u) != 0) {
            lf = min(lf, static_cast<int>(c));
            ll = max(ll, static_cast<int>(c));
        }
    atomicMin(&s_first, lf);
    atomicMax(&s_last, ll);
    __syncthreads();
    const int first = s_first, last = s_last;
    // columns in chunks of blockDim: block-wide exclusive scan with a running carry
    for (unsigned c0 = 0; c0 < W; c0 += nt) {
        const unsigned c = c0 + tid;
        // dewarp_impl.h:59-62: columns outside [first, last] are never visited, status == 0 is skipped
        const bool on = c < W && last >= first && static_cast<int>(c) >= first && static_cast<int>(c) <= last &&
                        status[c] != 0;
        uint32_t mine = 0;
        if (on)
            for (unsigned s = 0; s < n_slabs; ++s) mine += cnt[static_cast<size_t>(s) * W + c];
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if ((tid & 31u) >= static_cast<unsigned>(d)) incl += v;
        }
        if ((tid & 31u) == 31u) s_warp[tid >> 5] = incl;
        __syncthreads();
        if (tid < 32) {
            uint32_t w = tid < (nt >> 5) ? s_warp[tid] : 0u;
            uint32_t wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, wi, d);
                if (tid >= static_cast<unsigned>(d)) wi += v;
            }
            s_warp[tid] = wi - w;  // exclusive prefix of the warp sums
        }
        __syncthreads();
        uint32_t off = s_carry + s_warp[tid >> 5] + (incl - mine);
        if (c < W) {
            for (unsigned s = 0; s < n_slabs; ++s) {
                base[static_cast<size_t>(s) * W + c] = on ? off : 0xffffffffu;
                if (on) off += cnt[static_cast<size_t>(s) * W + c];
            }
        }
        __syncthreads();
        if (tid == nt - 1) s_carry = s_carry + s_warp[tid >> 5] + incl;
        __syncthreads();
    }
    if (tid == 0) *total = s_carry;
}

template <typename T>
__global__ void __launch_bounds__(256) k3_emit_kernel(const uint32_t* __restrict__ range, const T* __restrict__ dir,
                                                      const T* __restrict__ off, const double* __restrict__ poses,
                                                      const uint64_t* __restrict__ timestamps, unsigned H,
                                                      unsigned W, uint32_t min_r, uint32_t max_r, unsigned n_cg,
                                                      unsigned n_slabs, const uint32_t* __restrict__ base,
                                                      T* __restrict__ points, uint32_t* __restrict__ col_idx,
                                                      uint64_t* __restrict__ ts_out) {
    const unsigned gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= n_cg * n_slabs) return;
    const unsigned cg = gw % n_cg, slab = gw / n_cg;
    const unsigned col = cg * 32u + (threadIdx.x & 31u);
    if (col >= W) return;
    uint32_t w = base[static_cast<size_t>(slab) * W + col];
    if (w == 0xffffffffu) return;
    T m[12];  // rows 0..2 of body_to_world[col], cast to T (dewarp_impl.h:64-65)
#pragma unroll
    for (int k = 0; k < 12; ++k) m[k] = static_cast<T>(poses[static_cast<size_t>(col) * 16 + k]);
    const uint64_t ts = ts_out != nullptr ? timestamps[col] : 0ull;
    const unsigned r0 = slab * kSlabRows, r1 = min(H, r0 + kSlabRows);
    for (unsigned row = r0; row < r1; ++row) {
        const size_t px = static_cast<size_t>(row) * W + col;
        const uint32_t r = range[px];
        if (r >= min_r && r <= max_r) {
            const T x = k3_project(r, dir[px * 3], off[px * 3]);
            const T y = k3_project(r, dir[px * 3 + 1], off[px * 3 + 1]);
            const T z = k3_project(r, dir[px * 3 + 2], off[px * 3 + 2]);
            T* o = points + static_cast<size_t>(w) * 3;
            o[0] = k3_pose_row(m, x, y, z);
            o[1] = k3_pose_row(m + 4, x, y, z);
            o[2] = k3_pose_row(m + 8, x, y, z);
            if (col_idx != nullptr) col_idx[w] = col;
            if (ts_out != nullptr) ts_out[w] = ts;
            ++w;
        }
    }
}

size_t dewarp_frame_scratch_bytes(unsigned H, unsigned W) {
    const size_t n_slabs = (H + kSlabRows - 1) / kSlabRows;
    return 2 * n_slabs * W * sizeof(uint32_t) + 16;
}

cudaError_t launch_dewarp_frame_count(const DewarpFrameArgs& a, cudaStream_t st) {
    const unsigned n_cg = (a.W + 31) / 32, n_slabs = (a.H + kSlabRows - 1) / kSlabRows;
    uint32_t* cnt = static_cast<uint32_t*>(a.scratch);
    uint32_t* base = cnt + static_cast<size_t>(n_slabs) * a.W;
    unsigned long long* total = reinterpret_cast<unsigned long long*>(
        static_cast<uint8_t*>(a.scratch) + ((2 * static_cast<size_t>(n_slabs) * a.W * 4 + 7) & ~static_cast<size_t>(7)));
    const unsigned warps = n_cg * n_slabs;
    k3_count_kernel<<<(warps + 7) / 8, 256, 0, st>>>(a.range, a.H, a.W, a.min_r, a.max_r, n_cg, n_slabs, cnt);
    k3_scan_kernel<<<1, 1024, 0, st>>>(cnt, a.status, a.W, n_slabs, base, total);
    count_launch(2);
    return cudaGetLastError();
}

const unsigned long long* dewarp_frame_total_ptr(const DewarpFrameArgs& a) {
    const size_t n_slabs = (a.H + kSlabRows - 1) / kSlabRows;
    return reinterpret_cast<const unsigned long long*>(
        static_cast<const uint8_t*>(a.scratch) + ((2 * n_slabs * a.W * 4 + 7) & ~static_cast<size_t>(7)));
}

cudaError_t launch_dewarp_frame_emit(const DewarpFrameArgs& a, cudaStream_t st) {
    const unsigned n_cg = (a.W + 31) / 32, n_slabs = (a.H + kSlabRows - 1) / kSlabRows;
    const uint32_t* cnt = static_cast<const uint32_t*>(a.scratch);
    const uint32_t* base = cnt + static_cast<size_t>(n_slabs) * a.W;
    const unsigned warps = n_cg * n_slabs;
    if (a.dtype == OB_F64)
        k3_emit_kernel<double><<<(warps + 7) / 8, 256, 0, st>>>(
            a.range, static_cast<const double*>(a.dir), static_cast<const double*>(a.off), a.poses, a.timestamps,
            a.H, a.W, a.min_r, a.max_r, n_cg, n_slabs, base, static_cast<double*>(a.points), a.col_idx, a.ts_out);
    else
        k3_emit_kernel<float><<<(warps + 7) / 8, 256, 0, st>>>(
            a.range, static_cast<const float*>(a.dir), static_cast<const float*>(a.off), a.poses, a.timestamps,
            a.H, a.W, a.min_r, a.max_r, n_cg, n_slabs, base, static_cast<float*>(a.points), a.col_idx, a.ts_out);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Batched form: dewarp(FrameSet, xyzluts, min_range, max_range) (pose_util.h:475, impl/dewarp_impl.h:84-102)
// -- the frames of a set, each with its own LUT / poses / status, in THREE launches for the whole set
// (the single-frame path above needs three per frame plus a host round trip for its count): the points of
// frame f follow those of frame f-1 in the output, exactly like the reference's concatenation.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k3_count_batch_kernel(const K3Frame* __restrict__ frames, uint32_t min_r,
                                                             uint32_t max_r) {
    const K3Frame& fr = frames[blockIdx.y];
    const unsigned gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= fr.n_cg * fr.n_slabs) return;
    const unsigned cg = gw % fr.n_cg, slab = gw / fr.n_cg;
    const unsigned col = cg * 32u + (threadIdx.x & 31u);
    if (col >= fr.W) return;
    const unsigned r0 = slab * kSlabRows, r1 = min(fr.H, r0 + kSlabRows);
    uint32_t c = 0;
    for (unsigned row = r0; row < r1; ++row) {
        const uint32_t r = fr.range[static_cast<size_t>(row) * fr.W + col];
        c += (r >= min_r && r <= max_r) ? 1u : 0u;
    }
    fr.cnt[static_cast<size_t>(slab) * fr.W + col] = c;
}

__global__ void __launch_bounds__(1024) k3_scan_batch_kernel(const K3Frame* __restrict__ frames,
                                                             unsigned long long* __restrict__ totals) {
    const K3Frame& fr = frames[blockIdx.x];
    __shared__ int s_first, s_last;
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const unsigned tid = threadIdx.x, nt = blockDim.x, W = fr.W, n_slabs = fr.n_slabs;
    if (tid == 0) {
        s_first = 0x7fffffff;
        s_last = -1;
        s_carry = 0;
    }
    __syncthreads();
    int lf = 0x7fffffff, ll = -1;
    for (unsigned c = tid; c < W; c += nt)
        if ((fr.status[c] & 1u) != 0) {
            lf = min(lf, static_cast<int>(c));
            ll = max(ll, static_cast<int>(c));
        }
    atomicMin(&s_first, lf);
    atomicMax(&s_last, ll);
    __syncthreads();
    const int first = s_first, last = s_last;
    for (unsigned c0 = 0; c0 < W; c0 += nt) {
        const unsigned c = c0 + tid;
        const bool on = c < W && last >= first && static_cast<int>(c) >= first && static_cast<int>(c) <= last &&
                        fr.status[c] != 0;
        uint32_t mine = 0;
        if (on)
            for (unsigned sl = 0; sl < n_slabs; ++sl) mine += fr.cnt[static_cast<size_t>(sl) * W + c];
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if ((tid & 31u) >= static_cast<unsigned>(d)) incl += v;
        }
        if ((tid & 31u) == 31u) s_warp[tid >> 5] = incl;
        __syncthreads();
        if (tid < 32) {
            uint32_t w = tid < (nt >> 5) ? s_warp[tid] : 0u;
            uint32_t wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, wi, d);
                if (tid >= static_cast<unsigned>(d)) wi += v;
            }
            s_warp[tid] = wi - w;
        }
        __syncthreads();
        uint32_t off = s_carry + s_warp[tid >> 5] + (incl - mine);
        if (c < W) {
            for (unsigned sl = 0; sl < n_slabs; ++sl) {
                fr.base[static_cast<size_t>(sl) * W + c] = on ? off : 0xffffffffu;
                if (on) off += fr.cnt[static_cast<size_t>(sl) * W + c];
            }
        }
        __syncthreads();
        if (tid == nt - 1) s_carry = s_carry + s_warp[tid >> 5] + incl;
        __syncthreads();
    }
    if (tid == 0) totals[blockIdx.x] = s_carry;
}

template <typename T>
__global__ void __launch_bounds__(256) k3_emit_batch_kernel(const K3Frame* __restrict__ frames,
                                                            const unsigned long long* __restrict__ totals,
                                                            uint32_t min_r, uint32_t max_r, T* __restrict__ points,
                                                            uint32_t* __restrict__ frame_idx,
                                                            uint32_t* __restrict__ col_idx, uint64_t* __restrict__ ts_out,
                                                            unsigned long long capacity) {
    const unsigned f = blockIdx.y;
    const K3Frame& fr = frames[f];
    const unsigned gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= fr.n_cg * fr.n_slabs) return;
    const unsigned cg = gw % fr.n_cg, slab = gw / fr.n_cg;
    const unsigned col = cg * 32u + (threadIdx.x & 31u);
    if (col >= fr.W) return;
    const uint32_t b = fr.base[static_cast<size_t>(slab) * fr.W + col];
    if (b == 0xffffffffu) return;
    unsigned long long fbase = 0;  // points of the frames before this one (a set holds a handful of frames)
    for (unsigned i = 0; i < f; ++i) fbase += totals[i];
    unsigned long long w = fbase + b;
    const T* dir = static_cast<const T*>(fr.dir);
    const T* off = static_cast<const T*>(fr.off);
    T m[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) m[k] = static_cast<T>(fr.poses[static_cast<size_t>(col) * 16 + k]);
    const uint64_t ts = (ts_out != nullptr && fr.timestamps != nullptr) ? fr.timestamps[col] : 0ull;
    const unsigned r0 = slab * kSlabRows, r1 = min(fr.H, r0 + kSlabRows);
    for (unsigned row = r0; row < r1; ++row) {
        const size_t px = static_cast<size_t>(row) * fr.W + col;
        const uint32_t r = fr.range[px];
        if (r >= min_r && r <= max_r) {
            if (w < capacity) {
                const T x = k3_project(r, dir[px * 3], off[px * 3]);
                const T y = k3_project(r, dir[px * 3 + 1], off[px * 3 + 1]);
                const T z = k3_project(r, dir[px * 3 + 2], off[px * 3 + 2]);
                T* o = points + w * 3;
                o[0] = k3_pose_row(m, x, y, z);
                o[1] = k3_pose_row(m + 4, x, y, z);
                o[2] = k3_pose_row(m + 8, x, y, z);
                if (frame_idx != nullptr) frame_idx[w] = fr.index;
                if (col_idx != nullptr) col_idx[w] = col;
                if (ts_out != nullptr) ts_out[w] = ts;
            }
            ++w;
        }
    }
}

size_t dewarp_frames_scratch_bytes(unsigned H, unsigned W) {
    const size_t n_slabs = (H + kSlabRows - 1) / kSlabRows;
    return 2 * n_slabs * W * sizeof(uint32_t);
}

cudaError_t launch_dewarp_frames(const K3Frame* frames_dev, unsigned n_frames, unsigned max_warps, uint32_t min_r,
                                 uint32_t max_r, int dtype, unsigned long long* totals_dev, void* points,
                                 uint32_t* frame_idx, uint32_t* col_idx, uint64_t* ts_out, unsigned long long capacity,
                                 cudaStream_t st) {
    if (n_frames == 0) return cudaSuccess;
    if (n_frames > 65535) return cudaErrorInvalidValue;
    const dim3 grid((max_warps + 7) / 8, n_frames);
    k3_count_batch_kernel<<<grid, 256, 0, st>>>(frames_dev, min_r, max_r);
    k3_scan_batch_kernel<<<n_frames, 1024, 0, st>>>(frames_dev, totals_dev);
    if (dtype == OB_F64)
        k3_emit_batch_kernel<double><<<grid, 256, 0, st>>>(frames_dev, totals_dev, min_r, max_r,
                                                            static_cast<double*>(points), frame_idx, col_idx, ts_out, capacity);
    else
        k3_emit_batch_kernel<float><<<grid, 256, 0, st>>>(frames_dev, totals_dev, min_r, max_r,
                                                           static_cast<float*>(points), frame_idx, col_idx, ts_out, capacity);
    count_launch(3);
    return cudaGetLastError();
}

}  // namespace ob
