// scan.cu -- cs_cumsum (csparse.py:767-784) as a single-pass decoupled look-back
// exclusive scan.  HBM-bound: reads c once (4 B), writes p and c once (8 B).
//
// Tiles of 2048 ints are claimed in launch order through an atomic ticket, so a
// tile only ever waits on tiles that are already resident (forward progress).
// Each tile publishes {flag, value} in ONE 64-bit word (flag in the top two
// bits), so no fence is needed between flag and payload.
#include "common.cuh"

namespace csb {

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;

constexpr unsigned long long ST_NONE = 0, ST_AGG = 1, ST_PREFIX = 2;

__device__ __forceinline__ unsigned long long st_pack(unsigned long long flag, long long v)
{
    return (flag << 62) | ((unsigned long long)v & 0x3fffffffffffffffull);
}
__device__ __forceinline__ long long st_value(unsigned long long w)
{
    return ((long long)(w << 2)) >> 2;   // sign-extend the 62-bit payload
}
__device__ __forceinline__ unsigned st_flag(unsigned long long w) { return (unsigned)(w >> 62); }

__device__ __forceinline__ long long warp_sum(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool VEC>
__global__ void __launch_bounds__(SCAN_THREADS)
k_excl_scan(csi *__restrict__ p, csi *__restrict__ c, int n,
            volatile unsigned long long *status, unsigned *ticket,
            long long *total, int *maxv)
{
    __shared__ unsigned s_tile;
    __shared__ long long s_warp[SCAN_WARPS];
    __shared__ long long s_tile_excl;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const long long idx0 = (long long)tile * SCAN_TILE + (long long)tid * SCAN_ITEMS;

    int v[SCAN_ITEMS] = {0, 0, 0, 0};
    if (VEC && idx0 + 3 < n) {
        int4 t = *reinterpret_cast<const int4 *>(c + idx0);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++)
            if (idx0 + k < n) v[k] = c[idx0 + k];
    }
    const long long tsum = (long long)v[0] + v[1] + v[2] + v[3];

    if (maxv) {
        int mx = INT_MIN;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++)
            if (idx0 + k < n) mx = max(mx, v[k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0 && mx != INT_MIN) atomicMax(maxv, mx);
    }

    // inclusive scan of the per-thread sums inside each warp
    long long inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();

    if (wid == 0) {
        const long long ws = lane < SCAN_WARPS ? s_warp[lane] : 0;
        long long winc = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        const long long block_sum = __shfl_sync(0xffffffffu, winc, SCAN_WARPS - 1);
        if (lane < SCAN_WARPS) s_warp[lane] = winc - ws;   // exclusive offset of each warp

        long long excl = 0;
        if (tile == 0) {
            if (lane == 0) status[0] = st_pack(ST_PREFIX, block_sum);
        } else {
            if (lane == 0) status[tile] = st_pack(ST_AGG, block_sum);
            long long look = (long long)tile - 1;            // newest tile of the window
            while (true) {
                const long long t = look - lane;             // lane 0 looks at the closest tile
                unsigned long long w = t >= 0 ? status[t] : st_pack(ST_PREFIX, 0);
                while (__any_sync(0xffffffffu, st_flag(w) == ST_NONE)) {
                    if (st_flag(w) == ST_NONE) w = status[t];
                }
                const unsigned pre = __ballot_sync(0xffffffffu, st_flag(w) == ST_PREFIX);
                const long long val = st_value(w);
                if (pre) {
                    const int first = __ffs(pre) - 1;        // closest tile holding a full prefix
                    excl += warp_sum(lane <= first ? val : 0);
                    break;
                }
                excl += warp_sum(val);
                look -= 32;
            }
            if (lane == 0) status[tile] = st_pack(ST_PREFIX, excl + block_sum);
        }
        if (lane == 0) s_tile_excl = excl;
    }
    __syncthreads();

    long long e = s_tile_excl + s_warp[wid] + (inc - tsum);
    if (VEC && idx0 + 3 < n) {
        int4 o;
        o.x = (int)e; e += v[0];
        o.y = (int)e; e += v[1];
        o.z = (int)e; e += v[2];
        o.w = (int)e;
        *reinterpret_cast<int4 *>(p + idx0) = o;
        *reinterpret_cast<int4 *>(c + idx0) = o;
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            const long long idx = idx0 + k;
            if (idx < n) {
                p[idx] = (int)e;
                c[idx] = (int)e;
                e += v[k];
            } else if (idx == n) {
                p[idx] = (int)e;
                *total = e;
            }
        }
    }
}

int launch_excl_scan(csi *d_p, csi *d_c, csi n, long long *d_total, int *d_max)
{
    if (n < 0) return set_error(CSB200_ERR_ARG, "cs_cumsum: n < 0");
    const int tiles = n / SCAN_TILE + 1;   // the tile holding index n always exists
    DevBuf<unsigned long long> st;
    CSB_TRY(st.alloc((size_t)tiles + 1));
    CSB_CUDA(cudaMemsetAsync(st.ptr, 0, ((size_t)tiles + 1) * sizeof(unsigned long long), stream()));
    if (d_max) CSB_CUDA(cudaMemsetAsync(d_max, 0x80, sizeof(int), stream()));
    unsigned *ticket = reinterpret_cast<unsigned *>(st.ptr + tiles);
    const bool vec = (((uintptr_t)d_p | (uintptr_t)d_c) & 15) == 0;
    if (vec)
        k_excl_scan<true><<<tiles, SCAN_THREADS, 0, stream()>>>(d_p, d_c, n, st.ptr, ticket, d_total, d_max);
    else
        k_excl_scan<false><<<tiles, SCAN_THREADS, 0, stream()>>>(d_p, d_c, n, st.ptr, ticket, d_total, d_max);
    CSB_LAUNCHED();
    return CSB200_OK;
}

}  // namespace csb
