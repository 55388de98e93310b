// binning.cu -- tile binning: prefix scan of tile counts, (tile|depth) key emission, a hand-written
// stable LSD radix sort on the 64-bit keys with int32 payload, and the tile-range extraction
// (SURVEY 8a rows a-4..a-7).  Integer work: results are bit-exact against the reference.
//
// Radix sort design:
//   * one histogram kernel reads the keys once and builds the 8 digit histograms (8-bit digits);
//   * a one-block planner marks passes whose digit is constant over all keys as skipped (for
//     tile|depth keys only ceil((32 + log2 tiles)/8) passes survive, typically 6 of 8, and fewer when
//     the depth exponent bits are constant) and assigns ping-pong buffers so that the last executed
//     pass lands in the caller's output;
//   * each executed pass is count -> per-digit scan over blocks -> stable scatter, 4096 keys per
//     block, ranks from warp match_any so equal keys keep their input order (this reproduces the
//     stable cub sort behind torch.sort; ties = emission order = Gaussian index).
//   * the element count can come from device memory (no host sync in the fused path).
#include "common.cuh"

namespace gstex {

// ------------------------------------------------------------------------------------------
// inclusive int32 scan: block partials -> one-block scan of partials -> add
// ------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int warp_incl_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Exclusive block scan of one int per thread (blockDim.x == SCAN_THREADS); returns the block total via
// `total`.  `warp_sums` is a caller-provided smem array of 32 ints.
__device__ __forceinline__ int block_excl_scan(int v, int *warp_sums, int &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int incl = warp_incl_scan(v);
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        int s = lane < nw ? warp_sums[lane] : 0;
        s = warp_incl_scan(s);
        warp_sums[lane] = s;  // inclusive over warps
    }
    __syncthreads();
    const int warp_off = warp == 0 ? 0 : warp_sums[warp - 1];
    total = warp_sums[(blockDim.x >> 5) - 1];
    __syncthreads();
    return warp_off + incl - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_partials_kernel(int n, const int32_t *__restrict__ in,
                                                                     int32_t *__restrict__ partials) {
    __shared__ int warp_sums[32];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) s += in[base + k];
    int total;
    block_excl_scan(s, warp_sums, total);
    if (threadIdx.x == 0) partials[blockIdx.x] = total;
}

// one block: exclusive scan of the partials in place (any count, chunked with a running carry)
__global__ void __launch_bounds__(SCAN_THREADS) scan_spine_kernel(int nparts, int32_t *__restrict__ partials) {
    __shared__ int warp_sums[32];
    int carry = 0;
    for (int start = 0; start < nparts; start += SCAN_THREADS) {
        const int i = start + threadIdx.x;
        const int v = i < nparts ? partials[i] : 0;
        int total;
        const int ex = block_excl_scan(v, warp_sums, total);
        if (i < nparts) partials[i] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(int n, const int32_t *__restrict__ in,
                                                                  const int32_t *__restrict__ partials,
                                                                  int32_t *__restrict__ out) {
    __shared__ int warp_sums[32];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = base + k < n ? in[base + k] : 0;
        s += v[k];
    }
    int total;
    int run = block_excl_scan(s, warp_sums, total) + partials[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        run += v[k];
        if (base + k < n) out[base + k] = run;
    }
}

// ------------------------------------------------------------------------------------------
// key emission: one thread per Gaussian (reference forward.cu:13-71; WRAPPED = the torus mode of forward.cu:34-36, 53-62)
// ------------------------------------------------------------------------------------------
template <bool WRAPPED>
__global__ void __launch_bounds__(256) emit_keys_kernel(int n, const float2 *__restrict__ centers,
                                                        const float2 *__restrict__ extents,
                                                        const float *__restrict__ depths,
                                                        const int32_t *__restrict__ cum_tiles_hit, int tiles_x,
                                                        int tiles_y, float fbw, int64_t cap,
                                                        int64_t *__restrict__ isect_ids,
                                                        int32_t *__restrict__ gaussian_ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 c = centers[i], e = extents[i];
    if (e.x <= 1e-4 && e.y <= 1e-4) return;
    int x0, y0, x1, y1;
    if (WRAPPED) tile_bbox_wrapped(c.x, c.y, e.x, e.y, fbw, x0, y0, x1, y1);
    else tile_bbox(c.x, c.y, e.x, e.y, tiles_x, tiles_y, fbw, x0, y0, x1, y1);
    int32_t cur = i == 0 ? 0 : cum_tiles_hit[i - 1];
    const int64_t depth_id = (int64_t)__float_as_int(depths[i]);  // sign-extends like the reference
    for (int ty = y0; ty < y1; ++ty)
        for (int tx = x0; tx < x1; ++tx) {
            int wy = ty, wx = tx;
            if (WRAPPED) {
                // The reference takes `ti % tile_bounds.y` with an unsigned (dim3) divisor (forward.cu:56-61): a negative
                // index is reduced modulo 2^32 first and its sign fix-up never fires.  Reproduced bit for bit.
                wy = (int)((unsigned)wy % (unsigned)tiles_y);
                wx = (int)((unsigned)wx % (unsigned)tiles_x);
            }
            const int64_t tile = (int64_t)wy * tiles_x + wx;
            if (cur >= 0 && cur < cap) {
                isect_ids[cur] = (tile << 32) | depth_id;
                gaussian_ids[cur] = i;
            }
            ++cur;
        }
}

// ------------------------------------------------------------------------------------------
// radix sort
// ------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;                      // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;    // 4096 keys per block
constexpr int RS_WARP_KEYS = 32 * RS_ITEMS;       // 512 consecutive keys per warp
constexpr int RS_PASSES = 8;

struct SortPlan {
    int32_t count;              // number of valid elements
    int32_t exec[RS_PASSES];    // 1 = pass runs
    int32_t src[RS_PASSES];     // 0 = input, 1 = output, 2 = temp
    int32_t dst[RS_PASSES];     // 1 = output, 2 = temp
    int32_t n_exec;
    int32_t pad[5];
    int32_t digit_base[RS_PASSES][256];  // exclusive prefix of the global digit histogram
};

__device__ __forceinline__ uint64_t sort_bits(int64_t k) { return (uint64_t)k ^ 0x8000000000000000ull; }
__device__ __forceinline__ int digit_of(uint64_t bits, int pass) { return (int)((bits >> (8 * pass)) & 255u); }

__global__ void rs_zero_kernel(int32_t *hist, int nwords) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nwords) hist[i] = 0;
}

// global digit histograms for all passes: hist[pass][256]
__global__ void __launch_bounds__(RS_THREADS) rs_histogram_kernel(int64_t m_cap, const int32_t *__restrict__ d_count,
                                                                  const int64_t *__restrict__ keys, int npass,
                                                                  int32_t *__restrict__ hist) {
    __shared__ int32_t sh[RS_PASSES * 256];
    const int64_t m = d_count ? min((int64_t)max(*d_count, 0), m_cap) : m_cap;
    for (int i = threadIdx.x; i < npass * 256; i += RS_THREADS) sh[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
    if (base < m) {
#pragma unroll 4
        for (int k = 0; k < RS_ITEMS; ++k) {
            const int64_t i = base + (int64_t)k * RS_THREADS + threadIdx.x;
            if (i < m) {
                const uint64_t b = sort_bits(keys[i]);
                for (int p = 0; p < npass; ++p) atomicAdd(&sh[p * 256 + digit_of(b, p)], 1);
            }
        }
    }
    __syncthreads();
    if (base < m)
        for (int i = threadIdx.x; i < npass * 256; i += RS_THREADS)
            if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// one block of 256 threads: decide which passes run and where they read / write
__global__ void __launch_bounds__(256) rs_plan_kernel(int64_t m_cap, const int32_t *__restrict__ d_count, int npass,
                                                      const int32_t *__restrict__ hist, SortPlan *__restrict__ plan) {
    __shared__ int warp_sums[32];
    __shared__ int trivial[RS_PASSES];
    const int64_t m = d_count ? min((int64_t)max(*d_count, 0), m_cap) : m_cap;
    if (threadIdx.x < RS_PASSES) trivial[threadIdx.x] = 0;
    __syncthreads();
    for (int p = 0; p < npass; ++p) {
        const int h = hist[p * 256 + threadIdx.x];
        if (h == (int)m) trivial[p] = 1;  // at most one thread per pass
        int total;
        const int ex = block_excl_scan(h, warp_sums, total);
        plan->digit_base[p][threadIdx.x] = ex;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        plan->count = (int32_t)m;
        int n_exec = 0;
        for (int p = 0; p < RS_PASSES; ++p) {
            const int run = (p < npass) && !trivial[p] && m > 1;
            plan->exec[p] = run;
            n_exec += run;
        }
        plan->n_exec = n_exec;
        int j = 0, prev = 0;
        for (int p = 0; p < RS_PASSES; ++p) {
            if (!plan->exec[p]) {
                plan->src[p] = 0;
                plan->dst[p] = 1;
                continue;
            }
            const int dst = ((n_exec - 1 - j) & 1) ? 2 : 1;
            plan->src[p] = prev;
            plan->dst[p] = dst;
            prev = dst;
            ++j;
        }
    }
}

struct SortBufs {
    const int64_t *k[3];
    const int32_t *v[3];
};

// per-block digit counts for one pass: block_hist[digit * nblocks + block]
__global__ void __launch_bounds__(RS_THREADS) rs_count_kernel(int pass, const SortPlan *__restrict__ plan,
                                                              SortBufs bufs, int nblocks,
                                                              int32_t *__restrict__ block_hist) {
    if (!plan->exec[pass]) return;
    const int64_t m = plan->count;
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
    __shared__ int32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    if (base < m) {
        const int64_t *__restrict__ keys = bufs.k[plan->src[pass]];
#pragma unroll 4
        for (int k = 0; k < RS_ITEMS; ++k) {
            const int64_t i = base + (int64_t)k * RS_THREADS + threadIdx.x;
            if (i < m) atomicAdd(&sh[digit_of(sort_bits(keys[i]), pass)], 1);
        }
    }
    __syncthreads();
    block_hist[threadIdx.x * nblocks + blockIdx.x] = sh[threadIdx.x];
}

// one block per digit: exclusive scan of that digit's counts over the blocks + global digit base
__global__ void __launch_bounds__(SCAN_THREADS) rs_scan_kernel(int pass, const SortPlan *__restrict__ plan, int nblocks,
                                                               int32_t *__restrict__ block_hist) {
    if (!plan->exec[pass]) return;
    __shared__ int warp_sums[32];
    int32_t *row = block_hist + (size_t)blockIdx.x * nblocks;
    int carry = plan->digit_base[pass][blockIdx.x];
    for (int start = 0; start < nblocks; start += SCAN_THREADS) {
        const int i = start + threadIdx.x;
        const int v = i < nblocks ? row[i] : 0;
        int total;
        const int ex = block_excl_scan(v, warp_sums, total);
        if (i < nblocks) row[i] = carry + ex;
        carry += total;
    }
}

// stable scatter of one pass.  Warp w of a block owns keys [w*512, (w+1)*512) of the block's tile and
// walks them 32 at a time in index order; a key's rank among equal digits is
//   (#equal digits in earlier blocks) + (#in earlier warps of this block) + (#earlier in this warp).
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(int pass, const SortPlan *__restrict__ plan, SortBufs bufs,
                                                                int64_t *out_k, int32_t *out_v, int64_t *tmp_k,
                                                                int32_t *tmp_v, int nblocks,
                                                                const int32_t *__restrict__ block_hist) {
    if (!plan->exec[pass]) return;
    const int64_t m = plan->count;
    const int64_t tile_base = (int64_t)blockIdx.x * RS_TILE;
    if (tile_base >= m) return;
    __shared__ int32_t warp_cnt[RS_WARPS][256];
    __shared__ int32_t digit_off[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int src = plan->src[pass], dst = plan->dst[pass];
    const int64_t *__restrict__ keys = bufs.k[src];
    const int32_t *__restrict__ vals = bufs.v[src];
    int64_t *__restrict__ dk = dst == 1 ? out_k : tmp_k;
    int32_t *__restrict__ dv = dst == 1 ? out_v : tmp_v;

    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&warp_cnt[0][0])[i] = 0;
    digit_off[threadIdx.x] = block_hist[threadIdx.x * nblocks + blockIdx.x];
    __syncthreads();

    int64_t key[RS_ITEMS];
    int32_t rank[RS_ITEMS];
    const int64_t warp_base = tile_base + (int64_t)warp * RS_WARP_KEYS;
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const int64_t i = warp_base + r * 32 + lane;
        const bool ok = i < m;
        key[r] = ok ? keys[i] : 0;
        // out-of-range lanes use digit 256+lane-free trick: give them a unique pseudo digit via the valid mask
        const int d = digit_of(sort_bits(key[r]), pass);
        const unsigned okmask = __ballot_sync(0xffffffffu, ok);
        unsigned peers = __match_any_sync(0xffffffffu, ok ? d : (256 + lane));
        peers &= okmask;
        int prev = 0;
        if (ok) prev = warp_cnt[warp][d];
        __syncwarp();
        if (ok && (peers & lt_mask) == 0) warp_cnt[warp][d] = prev + __popc(peers);
        __syncwarp();
        rank[r] = prev + __popc(peers & lt_mask);
    }
    __syncthreads();
    // exclusive scan over warps for each digit (thread = digit)
    {
        int run = digit_off[threadIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const int c = warp_cnt[w][threadIdx.x];
            warp_cnt[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const int64_t i = warp_base + r * 32 + lane;
        if (i < m) {
            const int d = digit_of(sort_bits(key[r]), pass);
            const int64_t pos = (int64_t)warp_cnt[warp][d] + rank[r];
            dk[pos] = key[r];
            dv[pos] = vals[i];
        }
    }
}

// n_exec == 0 (all keys equal, or fewer than two elements): the output is the input
__global__ void __launch_bounds__(256) rs_passthrough_kernel(const SortPlan *__restrict__ plan, const int64_t *__restrict__ keys,
                                                             const int32_t *__restrict__ vals, int64_t *__restrict__ out_k,
                                                             int32_t *__restrict__ out_v) {
    if (plan->n_exec != 0) return;
    const int64_t m = plan->count;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        out_k[i] = keys[i];
        out_v[i] = vals[i];
    }
}

// tile ranges from the sorted keys (reference forward.cu:76-98)
__global__ void __launch_bounds__(256) tile_edges_kernel(int64_t m_cap, const int32_t *__restrict__ d_count,
                                                         const int64_t *__restrict__ keys, int2 *__restrict__ tile_bins) {
    const int64_t m = d_count ? min((int64_t)max(*d_count, 0), m_cap) : m_cap;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int32_t cur = (int32_t)(keys[i] >> 32);
    if (i == 0) tile_bins[cur].x = 0;
    if (i == m - 1) tile_bins[cur].y = (int32_t)m;
    if (i == 0) return;
    const int32_t prev = (int32_t)(keys[i - 1] >> 32);
    if (prev != cur) {
        tile_bins[prev].y = (int32_t)i;
        tile_bins[cur].x = (int32_t)i;
    }
}

struct SortLayout {
    size_t plan_off, hist_off, block_hist_off, tmp_k_off, tmp_v_off, total;
    int nblocks;
};

static SortLayout sort_layout(int64_t m) {
    SortLayout L;
    L.nblocks = (int)ceil_div64(m > 0 ? m : 1, RS_TILE);
    size_t off = 0;
    L.plan_off = off;
    off = align_up(off + sizeof(SortPlan), 256);
    L.hist_off = off;
    off = align_up(off + sizeof(int32_t) * RS_PASSES * 256, 256);
    L.block_hist_off = off;
    off = align_up(off + sizeof(int32_t) * 256 * (size_t)L.nblocks, 256);
    L.tmp_k_off = off;
    off = align_up(off + sizeof(int64_t) * (size_t)(m > 0 ? m : 1), 256);
    L.tmp_v_off = off;
    off = align_up(off + sizeof(int32_t) * (size_t)(m > 0 ? m : 1), 256);
    L.total = off;
    return L;
}

}  // namespace gstex

using namespace gstex;

extern "C" size_t gstex_scan_temp_bytes(int n) {
    return align_up(sizeof(int32_t) * (size_t)ceil_div(n > 0 ? n : 1, SCAN_TILE), 256);
}

extern "C" int gstex_cumsum_i32(int n, const int32_t *in, int32_t *out, void *temp, size_t temp_bytes,
                                gstex_stream_t stream) {
    GSTEX_REQUIRE(n >= 0, GSTEX_E_INVALID, "cumsum: n = %d", n);
    if (n == 0) return GSTEX_OK;
    GSTEX_REQUIRE(temp && temp_bytes >= gstex_scan_temp_bytes(n), GSTEX_E_WORKSPACE,
                  "cumsum: temp too small (%zu < %zu)", temp_bytes, gstex_scan_temp_bytes(n));
    const int nparts = ceil_div(n, SCAN_TILE);
    int32_t *partials = (int32_t *)temp;
    cudaStream_t s = as_stream(stream);
    scan_partials_kernel<<<nparts, SCAN_THREADS, 0, s>>>(n, in, partials);
    GSTEX_LAUNCH_OK("scan_partials_kernel");
    scan_spine_kernel<<<1, SCAN_THREADS, 0, s>>>(nparts, partials);
    GSTEX_LAUNCH_OK("scan_spine_kernel");
    scan_apply_kernel<<<nparts, SCAN_THREADS, 0, s>>>(n, in, partials, out);
    GSTEX_LAUNCH_OK("scan_apply_kernel");
    return GSTEX_OK;
}

static int map_intersects_impl(int n, int64_t num_intersects, const float *centers, const float *extents,
                               const float *depths, const int32_t *cum_tiles_hit, int tiles_x, int tiles_y,
                               int block_width, int wrapped, int64_t *isect_ids, int32_t *gaussian_ids,
                               gstex_stream_t stream) {
    GSTEX_REQUIRE(n >= 0 && block_width > 0 && tiles_x >= 0 && tiles_y >= 0, GSTEX_E_INVALID,
                  "map_gaussian_to_intersects: n = %d, bw = %d", n, block_width);
    GSTEX_REQUIRE(!wrapped || (tiles_x > 0 && tiles_y > 0), GSTEX_E_INVALID,
                  "map_gaussian_to_intersects: wrapped binning needs a non-empty tile grid (%d x %d)", tiles_x, tiles_y);
    if (n == 0) return GSTEX_OK;
    const float2 *c2 = (const float2 *)centers, *e2 = (const float2 *)extents;
    if (wrapped)
        emit_keys_kernel<true><<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
            n, c2, e2, depths, cum_tiles_hit, tiles_x, tiles_y, (float)block_width, num_intersects, isect_ids, gaussian_ids);
    else
        emit_keys_kernel<false><<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
            n, c2, e2, depths, cum_tiles_hit, tiles_x, tiles_y, (float)block_width, num_intersects, isect_ids, gaussian_ids);
    GSTEX_LAUNCH_OK("emit_keys_kernel");
    return GSTEX_OK;
}

extern "C" int gstex_map_gaussian_to_intersects(int n, int64_t num_intersects, const float *centers,
                                                const float *extents, const float *depths,
                                                const int32_t *cum_tiles_hit, int tiles_x, int tiles_y,
                                                int block_width, int64_t *isect_ids, int32_t *gaussian_ids,
                                                gstex_stream_t stream) {
    return map_intersects_impl(n, num_intersects, centers, extents, depths, cum_tiles_hit, tiles_x, tiles_y, block_width,
                               0, isect_ids, gaussian_ids, stream);
}

extern "C" int gstex_map_gaussian_to_intersects_wrapped(int n, int64_t num_intersects, const float *centers,
                                                        const float *extents, const float *depths,
                                                        const int32_t *cum_tiles_hit, int tiles_x, int tiles_y,
                                                        int block_width, int64_t *isect_ids, int32_t *gaussian_ids,
                                                        gstex_stream_t stream) {
    return map_intersects_impl(n, num_intersects, centers, extents, depths, cum_tiles_hit, tiles_x, tiles_y, block_width,
                               1, isect_ids, gaussian_ids, stream);
}

extern "C" size_t gstex_sort_temp_bytes(int64_t m) { return sort_layout(m).total; }

extern "C" int gstex_sort_pairs(int64_t m, const int64_t *keys_in, const int32_t *vals_in, int64_t *keys_out,
                                int32_t *vals_out, int end_bit, const int32_t *d_count, void *temp,
                                size_t temp_bytes, gstex_stream_t stream) {
    GSTEX_REQUIRE(m >= 0 && m < (int64_t)1 << 31, GSTEX_E_INVALID, "sort_pairs: m = %lld", (long long)m);
    GSTEX_REQUIRE(end_bit >= 1 && end_bit <= 64, GSTEX_E_INVALID, "sort_pairs: end_bit = %d", end_bit);
    if (m == 0) return GSTEX_OK;
    const SortLayout L = sort_layout(m);
    GSTEX_REQUIRE(temp && temp_bytes >= L.total, GSTEX_E_WORKSPACE, "sort_pairs: temp too small (%zu < %zu)",
                  temp_bytes, L.total);
    char *base = (char *)temp;
    SortPlan *plan = (SortPlan *)(base + L.plan_off);
    int32_t *hist = (int32_t *)(base + L.hist_off);
    int32_t *block_hist = (int32_t *)(base + L.block_hist_off);
    int64_t *tmp_k = (int64_t *)(base + L.tmp_k_off);
    int32_t *tmp_v = (int32_t *)(base + L.tmp_v_off);
    // with end_bit < 64 the keys are non-negative, so the flipped sign bit is constant and pass 7 is trivial
    const int npass = end_bit >= 57 ? RS_PASSES : ceil_div(end_bit, 8);
    cudaStream_t s = as_stream(stream);
    rs_zero_kernel<<<ceil_div(RS_PASSES * 256, 256), 256, 0, s>>>(hist, RS_PASSES * 256);
    GSTEX_LAUNCH_OK("rs_zero_kernel");
    rs_histogram_kernel<<<L.nblocks, RS_THREADS, 0, s>>>(m, d_count, keys_in, npass, hist);
    GSTEX_LAUNCH_OK("rs_histogram_kernel");
    rs_plan_kernel<<<1, 256, 0, s>>>(m, d_count, npass, hist, plan);
    GSTEX_LAUNCH_OK("rs_plan_kernel");
    SortBufs bufs;
    bufs.k[0] = keys_in;
    bufs.k[1] = keys_out;
    bufs.k[2] = tmp_k;
    bufs.v[0] = vals_in;
    bufs.v[1] = vals_out;
    bufs.v[2] = tmp_v;
    for (int p = 0; p < npass; ++p) {
        rs_count_kernel<<<L.nblocks, RS_THREADS, 0, s>>>(p, plan, bufs, L.nblocks, block_hist);
        GSTEX_LAUNCH_OK("rs_count_kernel");
        rs_scan_kernel<<<256, SCAN_THREADS, 0, s>>>(p, plan, L.nblocks, block_hist);
        GSTEX_LAUNCH_OK("rs_scan_kernel");
        rs_scatter_kernel<<<L.nblocks, RS_THREADS, 0, s>>>(p, plan, bufs, keys_out, vals_out, tmp_k, tmp_v, L.nblocks,
                                                           block_hist);
        GSTEX_LAUNCH_OK("rs_scatter_kernel");
    }
    const int pt_blocks = (int)min((int64_t)1184, ceil_div64(m, 256));
    rs_passthrough_kernel<<<pt_blocks, 256, 0, s>>>(plan, keys_in, vals_in, keys_out, vals_out);
    GSTEX_LAUNCH_OK("rs_passthrough_kernel");
    return GSTEX_OK;
}

extern "C" int gstex_get_tile_bin_edges(int64_t m, const int64_t *isect_ids_sorted, int32_t *tile_bins,
                                        const int32_t *d_count, gstex_stream_t stream) {
    GSTEX_REQUIRE(m >= 0 && m < (int64_t)1 << 31, GSTEX_E_INVALID, "get_tile_bin_edges: m = %lld", (long long)m);
    if (m == 0) return GSTEX_OK;
    tile_edges_kernel<<<(unsigned)ceil_div64(m, 256), 256, 0, as_stream(stream)>>>(m, d_count, isect_ids_sorted,
                                                                                   (int2 *)tile_bins);
    GSTEX_LAUNCH_OK("tile_edges_kernel");
    return GSTEX_OK;
}
