// scan.cu -- cs_cumsum (csparse.py:767-784) as a single-pass decoupled look-back
// exclusive scan.  HBM-bound: reads c once (4 B), writes p and c once (8 B).
//
// Tiles of 8192 ints are claimed in launch order through an atomic ticket, so a
// tile only ever waits on tiles that are already resident (forward progress).
// Each tile publishes {flag, value} in ONE 64-bit word (flag in the top two
// bits), so no fence is needed between flag and payload.
#include "common.cuh"

namespace csb {

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ROUNDS = 4;                       // 16-byte chunks per thread
constexpr int SCAN_ITEMS = 4 * SCAN_ROUNDS;          // 16 ints per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS; // 8192 ints per tile
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
__device__ __forceinline__ long long warp_inclusive(long long v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Tile layout: round q (0..3) covers the 16-byte chunks q*512 .. q*512+511 of the tile, thread
// t owns chunk q*512 + t of every round: all four loads are coalesced and in flight together.
template <bool VEC>
__global__ void __launch_bounds__(SCAN_THREADS)
k_excl_scan(csi *__restrict__ p, csi *__restrict__ c, int n,
            volatile unsigned long long *status, unsigned *ticket,
            long long *total, int *maxv)
{
    __shared__ unsigned s_tile;
    __shared__ long long s_part[SCAN_ROUNDS * SCAN_WARPS];   // (round, warp) sums, round-major
    __shared__ long long s_tile_excl;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const long long base = (long long)tile * SCAN_TILE;

    int v[SCAN_ROUNDS][4];
    long long s[SCAN_ROUNDS], inc[SCAN_ROUNDS];
    int mx = INT_MIN;
#pragma unroll
    for (int q = 0; q < SCAN_ROUNDS; q++) {
        const long long idx0 = base + (long long)(q * SCAN_THREADS + tid) * 4;
        if (VEC && idx0 + 3 < n) {
            const int4 t = *reinterpret_cast<const int4 *>(c + idx0);
            v[q][0] = t.x; v[q][1] = t.y; v[q][2] = t.z; v[q][3] = t.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) v[q][k] = idx0 + k < n ? c[idx0 + k] : 0;
        }
    }
#pragma unroll
    for (int q = 0; q < SCAN_ROUNDS; q++) {
        const long long idx0 = base + (long long)(q * SCAN_THREADS + tid) * 4;
        s[q] = (long long)v[q][0] + v[q][1] + v[q][2] + v[q][3];
        if (maxv) {
#pragma unroll
            for (int k = 0; k < 4; k++) if (idx0 + k < n) mx = max(mx, v[q][k]);
        }
        inc[q] = warp_inclusive(s[q], lane);
        if (lane == 31) s_part[q * SCAN_WARPS + wid] = inc[q];
    }
    if (maxv) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0 && mx != INT_MIN) atomicMax(maxv, mx);
    }
    __syncthreads();

    if (wid == 0) {
        // 64 partial sums, two per lane, in tile order
        const long long a = s_part[2 * lane], b = s_part[2 * lane + 1];
        const long long pinc = warp_inclusive(a + b, lane);
        const long long block_sum = __shfl_sync(0xffffffffu, pinc, 31);
        s_part[2 * lane] = pinc - a - b;              // exclusive offsets
        s_part[2 * lane + 1] = pinc - b;

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

    const long long tile_excl = s_tile_excl;
#pragma unroll
    for (int q = 0; q < SCAN_ROUNDS; q++) {
        const long long idx0 = base + (long long)(q * SCAN_THREADS + tid) * 4;
        long long e = tile_excl + s_part[q * SCAN_WARPS + wid] + (inc[q] - s[q]);
        if (VEC && idx0 + 3 < n) {
            int4 o;
            o.x = (int)e; e += v[q][0];
            o.y = (int)e; e += v[q][1];
            o.z = (int)e; e += v[q][2];
            o.w = (int)e;
            *reinterpret_cast<int4 *>(p + idx0) = o;
            *reinterpret_cast<int4 *>(c + idx0) = o;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const long long idx = idx0 + k;
                if (idx < n) {
                    p[idx] = (int)e;
                    c[idx] = (int)e;
                    e += v[q][k];
                } else if (idx == n) {
                    p[idx] = (int)e;
                    *total = e;
                }
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
