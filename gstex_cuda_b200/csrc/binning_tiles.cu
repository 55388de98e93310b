// binning_tiles.cu -- fused tile binning for the no-host-sync pipeline: bucket by tile, then sort each tile's
// short list by depth inside one CTA.
//
// The reference sorts ALL (tile | depth) keys globally (torch.sort, utils.py:159): at 1080p that is 6 radix passes
// over M x 12 bytes.  But the high 32 bits of the key only say "which tile" and the tile lists are short (about
// 400 entries at C4), so the same permutation is obtained much cheaper:
//   1. tb_count   : per Gaussian, one atomicAdd per touched tile           -> tile_count[T]
//   2. tb_scan    : exclusive scan over the T tiles (one CTA)              -> tile_bins (the reference's
//                   get_tile_bin_edges output, (0,0) for empty tiles) and the intersection count M
//   3. tb_scatter : per Gaussian, claim a slot in each touched tile's range with an atomic and drop the 64-bit
//                   key (depth bits << 32 | gaussian id) there (order inside the range is arbitrary)
//   4. tb_sort    : one CTA per tile sorts its range by that key in shared memory (bitonic network in its
//                   "mirror" form: every comparator puts the minimum at the lower index, so a virtual +inf padding
//                   never moves and comparators reaching past the end are simply skipped -> any length works).
// The keys are unique (a Gaussian occurs once per tile) and ordering by (depth bits, gaussian id) is exactly the
// order of the reference's stable sort, whose ties keep emission order = ascending Gaussian index
// (forward.cu:45-68).  Depth bits compare as unsigned: emitted depths are > 0.01.
// Results are bit-identical to map_gaussian_to_intersects + sort + get_tile_bin_edges; tests/test_gpu_binning.py
// checks that against both the staged path of this library and the reference CUDA extension.
#include "common.cuh"

namespace gstex {

// Per-tile counters live one per 32-byte sector: with 8 counters in a sector the ~400 atomics each tile receives
// serialise against its 7 neighbours' in the L2 atomic unit (measured on C4: count + scatter 0.142 -> 0.09 ms).
#ifndef GSTEX_TB_STRIDE
#define GSTEX_TB_STRIDE 8
#endif
constexpr int TB_STRIDE = GSTEX_TB_STRIDE;  // ints between consecutive tiles' counters
constexpr int TB_SMALL = 1024;    // keys sorted in static shared memory by 256 threads
constexpr int TB_MEDIUM = 8192;   // keys sorted in 64 KB of dynamic shared memory by 1024 threads
                                  // longer lists: same network, in place in global memory (L2), 1024 threads

__global__ void __launch_bounds__(256) tb_count_kernel(int n, const float2 *__restrict__ centers,
                                                       const float2 *__restrict__ extents, int tiles_x, int tiles_y,
                                                       float fbw, int32_t *__restrict__ tile_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 c = centers[i], e = extents[i];
    if (e.x <= 1e-4 && e.y <= 1e-4) return;  // reference forward.cu:32
    int x0, y0, x1, y1;
    tile_bbox(c.x, c.y, e.x, e.y, tiles_x, tiles_y, fbw, x0, y0, x1, y1);
    for (int ty = y0; ty < y1; ++ty)
        for (int tx = x0; tx < x1; ++tx) atomicAdd(&tile_count[(ty * tiles_x + tx) * TB_STRIDE], 1);
}

// one CTA of 1024 threads: exclusive scan of tile_count -> tile_start, tile_bins, total
__global__ void __launch_bounds__(1024) tb_scan_kernel(int num_tiles, const int32_t *__restrict__ tile_count,
                                                       int32_t *__restrict__ tile_start, int2 *__restrict__ tile_bins,
                                                       int32_t *__restrict__ num_intersects, int64_t cap,
                                                       int32_t *__restrict__ max_seen) {
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < num_tiles; base += 1024) {
        const int t = base + threadIdx.x;
        const int v = t < num_tiles ? tile_count[t * TB_STRIDE] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int s = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += u;
            }
            warp_sums[lane] = s;
        }
        __syncthreads();
        const int carry = carry_s;
        const int start = carry + (warp ? warp_sums[warp - 1] : 0) + incl - v;
        if (t < num_tiles) {
            tile_start[t] = start;
            // entries past the capacity of the key / id buffers are dropped by the scatter
            const int s_c = (int)min((int64_t)start, cap), e_c = (int)min((int64_t)start + v, cap);
            tile_bins[t] = (v > 0 && e_c > s_c) ? make_int2(s_c, e_c) : make_int2(0, 0);
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_sums[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *num_intersects = carry_s;
        // running maximum over the calls that share `max_seen` (same stream): what a no-host-sync caller compares with its
        // capacity once in a while to detect dropped intersections
        if (max_seen && carry_s > *max_seen) *max_seen = carry_s;
    }
}

__global__ void __launch_bounds__(256) tb_scatter_kernel(int n, const float2 *__restrict__ centers,
                                                         const float2 *__restrict__ extents,
                                                         const float *__restrict__ depths, int tiles_x, int tiles_y,
                                                         float fbw, const int32_t *__restrict__ tile_start,
                                                         int32_t *__restrict__ tile_fill, int64_t cap,
                                                         unsigned long long *__restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 c = centers[i], e = extents[i];
    if (e.x <= 1e-4 && e.y <= 1e-4) return;
    int x0, y0, x1, y1;
    tile_bbox(c.x, c.y, e.x, e.y, tiles_x, tiles_y, fbw, x0, y0, x1, y1);
    const unsigned long long key = ((unsigned long long)__float_as_uint(depths[i]) << 32) | (unsigned)i;
    for (int ty = y0; ty < y1; ++ty)
        for (int tx = x0; tx < x1; ++tx) {
            const int t = ty * tiles_x + tx;
            const int64_t pos = (int64_t)tile_start[t] + atomicAdd(&tile_fill[t * TB_STRIDE], 1);
            if (pos < cap) keys[pos] = key;
        }
}

__device__ __forceinline__ int pow2_ceil(int n) { return n <= 1 ? 1 : 1 << (32 - __clz(n - 1)); }

// Mirror-form bitonic network over `a[0..n)` (shared or global memory), P = pow2_ceil(n) virtual elements.
// LD / ST are functors so that the global-memory variant can bypass L1.
template <class LD, class ST>
__device__ __forceinline__ void bitonic_sort(int n, int P, LD ld, ST st) {
    const int tid = threadIdx.x, nthr = blockDim.x, half_p = P >> 1;
    for (int k = 2; k <= P; k <<= 1) {
        const int half = k >> 1;
        for (int c = tid; c < half_p; c += nthr) {  // mirror step: i <-> block_end - i
            const int off = c & (half - 1), b = (c - off) << 1;  // b = (c / half) * k
            const int lo = b + off, hi = b + (k - 1 - off);
            if (hi < n) {
                const unsigned long long x = ld(lo), y = ld(hi);
                if (y < x) { st(lo, y); st(hi, x); }
            }
        }
        __syncthreads();
        for (int j = half >> 1; j > 0; j >>= 1) {  // half-cleaners
            for (int c = tid; c < half_p; c += nthr) {
                const int lo = ((c & ~(j - 1)) << 1) | (c & (j - 1)), hi = lo + j;
                if (hi < n) {
                    const unsigned long long x = ld(lo), y = ld(hi);
                    if (y < x) { st(lo, y); st(hi, x); }
                }
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------
// Register-resident variant of the same network for the short lists (n <= TB_SMALL, the common case: ~400 entries
// per tile on the C4 scene).  128 threads hold E = P/128 consecutive elements each (element i lives in thread i / E,
// slot i % E).  A pass is an XOR mask m - the mirror step of stage k is i <-> i ^ (k-1), a half-cleaner is i <-> i ^ j -
// and, depending on where the partner lives, it is
//   m < E        : a compare-exchange between two registers of the same thread,
//   m / E < 32   : a warp shuffle with lane ^ (m / E)  (the partner's slot is e for a cleaner, E-1-e for a mirror),
//   otherwise    : one round trip through shared memory (write own elements, barrier, read the partners').
// For P = 512 that is 17 + 25 + 3 of the 45 passes, i.e. 3 block barriers pairs instead of 45.  The comparator rule is the one
// of bitonic_sort above (minimum to the lower index, +inf padding never moves), so the output is identical.
// ------------------------------------------------------------------------------------------
constexpr int TBR_THREADS = 128;
constexpr unsigned long long TB_INF = ~0ull;

__device__ __forceinline__ void cswap(unsigned long long &a, unsigned long long &b) {  // a <- min, b <- max
    const unsigned long long lo = a < b ? a : b, hi = a < b ? b : a;
    a = lo;
    b = hi;
}

template <int E, int M>
__device__ __forceinline__ void pass_in_thread(unsigned long long (&v)[E]) {
#pragma unroll
    for (int e = 0; e < E; ++e)
        if ((e ^ M) > e && (e ^ M) < E) cswap(v[e], v[e ^ M]);
}

template <int E>
__device__ __forceinline__ void sort_pass(unsigned long long (&v)[E], int m, bool mirror, unsigned long long *sk,
                                          int t) {
    constexpr int LOGE = E == 1 ? 0 : E == 2 ? 1 : E == 4 ? 2 : 3;
    const int mt = m >> LOGE;
    if (mt == 0) {  // m < E: both elements of every comparator are in this thread
        switch (m) {
            case 1: pass_in_thread<E, 1>(v); break;
            case 2: pass_in_thread<E, 2>(v); break;
            case 3: pass_in_thread<E, 3>(v); break;
            case 4: pass_in_thread<E, 4>(v); break;
            default: pass_in_thread<E, 7>(v); break;
        }
        return;
    }
    const int hb = 1 << (31 - __clz(m));  // element i is the lower end of its comparator iff bit hb of i is clear
    unsigned long long other[E];
    if (mt < 32) {
#pragma unroll
        for (int e = 0; e < E; ++e) other[e] = __shfl_xor_sync(0xffffffffu, v[e], mt);
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) sk[t * E + e] = v[e];
        __syncthreads();
#pragma unroll
        for (int e = 0; e < E; ++e) other[e] = sk[((t ^ mt) << LOGE) + e];
        __syncthreads();
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const unsigned long long o = mirror ? other[E - 1 - e] : other[e];  // slot e ^ (m & (E-1)): E-1-e or e
        const bool lower = (((t << LOGE) + e) & hb) == 0;
        v[e] = lower ? (v[e] < o ? v[e] : o) : (v[e] < o ? o : v[e]);
    }
}

// P is a template parameter so that both loops unroll and every pass's mask, partner distance and comparator rule are
// compile-time constants (no switch / clz / variable shifts at run time).
template <int E, int P, class EMIT>
__device__ __forceinline__ void tile_sort_registers(const unsigned long long *__restrict__ g, int n,
                                                    unsigned long long *sk, EMIT emit) {
    const int t = threadIdx.x;
    unsigned long long v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = (t * E + e) < n ? g[t * E + e] : TB_INF;
#pragma unroll
    for (int k = 2; k <= P; k <<= 1) {
        sort_pass<E>(v, k - 1, true, sk, t);
#pragma unroll
        for (int j = k >> 2; j > 0; j >>= 1) sort_pass<E>(v, j, false, sk, t);
    }
#pragma unroll
    for (int e = 0; e < E; ++e)
        if ((t * E + e) < n) emit(t * E + e, v[e]);
}

struct TileSortArgs {
    const int2 *tile_bins;
    unsigned long long *keys;     // (depth bits << 32 | gaussian id), bucketed by tile
    int32_t *ids_out;             // gaussian_ids_sorted
    int64_t *isect_out;           // optional isect_ids_sorted = tile << 32 | depth bits
};

__device__ __forceinline__ void tile_sort_emit(const TileSortArgs &a, int tile, int start, int idx,
                                               unsigned long long key) {
    a.ids_out[start + idx] = (int32_t)(unsigned)(key & 0xffffffffull);
    if (a.isect_out) a.isect_out[start + idx] = ((int64_t)tile << 32) | (int64_t)(key >> 32);
}

// CLASS 0: n <= TB_SMALL (registers + static smem); CLASS 1: longer lists - dynamic smem up to TB_MEDIUM, global memory beyond
// Class 0 runs one CTA per tile; classes 1 and 2 (rare: more than 1024 entries in a tile) run a small grid whose CTAs
// stride over all tiles looking for theirs, so that a frame without such tiles pays two near-empty launches instead of
// two full grids of 1024-thread CTAs.
template <int CLASS>
__global__ void __launch_bounds__(CLASS == 0 ? TBR_THREADS : 1024) tb_sort_kernel(const TileSortArgs a, int num_tiles) {
    extern __shared__ unsigned long long sk_dyn[];
    __shared__ unsigned long long sk_static[CLASS == 0 ? TB_SMALL : 1];
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {  // CTA-uniform: barriers inside are safe
        const int2 r = a.tile_bins[tile];
        const int n = r.y - r.x;
        if (CLASS == 0 && (n <= 0 || n > TB_SMALL)) continue;
        if (CLASS == 1 && n <= TB_SMALL) continue;  // class 1 = every longer list: shared memory up to TB_MEDIUM, else global
        unsigned long long *__restrict__ g = a.keys + r.x;
        const int P = pow2_ceil(n);
        if (CLASS == 0) {
            auto emit = [&](int i, unsigned long long key) { tile_sort_emit(a, tile, r.x, i, key); };
            if (P <= 32) tile_sort_registers<1, 32>(g, n, sk_static, emit);  // tiny lists: 15 passes, one warp's worth
            else if (P <= TBR_THREADS) tile_sort_registers<1, TBR_THREADS>(g, n, sk_static, emit);
            else if (P == 2 * TBR_THREADS) tile_sort_registers<2, 2 * TBR_THREADS>(g, n, sk_static, emit);
            else if (P == 4 * TBR_THREADS) tile_sort_registers<4, 4 * TBR_THREADS>(g, n, sk_static, emit);
            else tile_sort_registers<8, 8 * TBR_THREADS>(g, n, sk_static, emit);
        } else if (n > TB_MEDIUM) {
            bitonic_sort(n, P, [&](int i) { return __ldcg(g + i); }, [&](int i, unsigned long long v) { __stcg(g + i, v); });
            for (int i = threadIdx.x; i < n; i += blockDim.x) tile_sort_emit(a, tile, r.x, i, __ldcg(g + i));
        } else {
            unsigned long long *sk = sk_dyn;
            for (int i = threadIdx.x; i < n; i += blockDim.x) sk[i] = g[i];
            __syncthreads();
            bitonic_sort(n, P, [&](int i) { return sk[i]; }, [&](int i, unsigned long long v) { sk[i] = v; });
            for (int i = threadIdx.x; i < n; i += blockDim.x) tile_sort_emit(a, tile, r.x, i, sk[i]);
        }
        __syncthreads();  // the shared buffer is reused by the CTA's next tile
    }
}

struct TileBinLayout {
    size_t count_off, fill_off, start_off, keys_off, total;
};

static TileBinLayout tile_bin_layout(int num_tiles, int64_t cap) {
    TileBinLayout L;
    const size_t t = (size_t)(num_tiles > 0 ? num_tiles : 1);
    size_t off = 0;
    L.count_off = off;
    off += sizeof(int32_t) * t * TB_STRIDE;       // count and fill are contiguous: one memset clears both
    L.fill_off = off;
    off = align_up(off + sizeof(int32_t) * t * TB_STRIDE, 256);
    L.start_off = off;
    off = align_up(off + sizeof(int32_t) * t, 256);
    L.keys_off = off;
    off = align_up(off + sizeof(unsigned long long) * (size_t)(cap > 0 ? cap : 1), 256);
    L.total = off;
    return L;
}

}  // namespace gstex

using namespace gstex;

extern "C" size_t gstex_bin_tiles_temp_bytes(int num_tiles, int64_t capacity) {
    return tile_bin_layout(num_tiles, capacity).total;
}

extern "C" int gstex_bin_tiles(int n, const float *centers, const float *extents, const float *depths, int tiles_x,
                               int tiles_y, int block_width, int64_t capacity, int32_t *gaussian_ids_sorted,
                               int64_t *isect_ids_sorted, int32_t *tile_bins, int32_t *num_intersects,
                               int32_t *max_intersects_seen, void *temp, size_t temp_bytes, gstex_stream_t stream) {
    GSTEX_REQUIRE(n >= 0 && block_width > 0 && tiles_x > 0 && tiles_y > 0, GSTEX_E_INVALID,
                  "bin_tiles: n = %d, bw = %d, tiles = %dx%d", n, block_width, tiles_x, tiles_y);
    GSTEX_REQUIRE(capacity >= 0 && capacity < ((int64_t)1 << 31), GSTEX_E_INVALID, "bin_tiles: capacity = %lld",
                  (long long)capacity);
    const int num_tiles = tiles_x * tiles_y;
    const TileBinLayout L = tile_bin_layout(num_tiles, capacity);
    GSTEX_REQUIRE(temp && temp_bytes >= L.total, GSTEX_E_WORKSPACE, "bin_tiles: temp too small (%zu < %zu)", temp_bytes,
                  L.total);
    cudaStream_t s = as_stream(stream);
    char *base = (char *)temp;
    int32_t *tile_count = (int32_t *)(base + L.count_off), *tile_fill = (int32_t *)(base + L.fill_off);
    int32_t *tile_start = (int32_t *)(base + L.start_off);
    unsigned long long *keys = (unsigned long long *)(base + L.keys_off);
    GSTEX_CUDA_OK(cudaMemsetAsync(tile_count, 0, sizeof(int32_t) * 2 * (size_t)num_tiles * TB_STRIDE, s));
    const float fbw = (float)block_width;
    if (n > 0) {
        tb_count_kernel<<<ceil_div(n, 256), 256, 0, s>>>(n, (const float2 *)centers, (const float2 *)extents, tiles_x,
                                                         tiles_y, fbw, tile_count);
        GSTEX_LAUNCH_OK("tb_count_kernel");
    }
    tb_scan_kernel<<<1, 1024, 0, s>>>(num_tiles, tile_count, tile_start, (int2 *)tile_bins, num_intersects, capacity,
                                      max_intersects_seen);
    GSTEX_LAUNCH_OK("tb_scan_kernel");
    if (n == 0 || capacity == 0) return GSTEX_OK;
    tb_scatter_kernel<<<ceil_div(n, 256), 256, 0, s>>>(n, (const float2 *)centers, (const float2 *)extents, depths,
                                                       tiles_x, tiles_y, fbw, tile_start, tile_fill, capacity, keys);
    GSTEX_LAUNCH_OK("tb_scatter_kernel");
    const TileSortArgs a{(const int2 *)tile_bins, keys, gaussian_ids_sorted, isect_ids_sorted};
    tb_sort_kernel<0><<<num_tiles, TBR_THREADS, 0, s>>>(a, num_tiles);
    GSTEX_LAUNCH_OK("tb_sort_kernel<0>");
    static const size_t medium_smem = sizeof(unsigned long long) * TB_MEDIUM;
    {
        static SmemOnceFlags once;
        const int rc = configure_dynamic_smem((const void *)tb_sort_kernel<1>, medium_smem, false, once);
        if (rc != GSTEX_OK) return rc;
    }
    const int rare_grid = num_tiles < 2 * 148 ? num_tiles : 2 * 148;
    tb_sort_kernel<1><<<rare_grid, 1024, medium_smem, s>>>(a, num_tiles);
    GSTEX_LAUNCH_OK("tb_sort_kernel<1>");
    return GSTEX_OK;
}
