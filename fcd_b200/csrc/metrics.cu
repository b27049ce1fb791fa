// Voxel-level evaluation counts on the device (SURVEY 8f rank 4 "voxel metrics on GPU"): what the reference's evaluate loop
// hands to MONAI's DiceMetric / ConfusionMatrixMetric (metrics.py:74-126, called from train.py:220) and to
// utils_common.evaluate_fp (utils/utils_common.py:37-60) after moving whole volumes to the host.  Everything here is
// integer counting, so the results are bit-exact against the numpy oracle (oracle/metrics.py); the ratios (Dice,
// precision, ...) are a handful of scalar operations done by the caller on the [items][4] count table.
//
//   fcd_confusion_counts  : per item (one (subject, channel) volume of n voxels): tp, fp, tn, fn of (pred > thr_p) against
//                           (label > thr_l) -- one streaming pass, HBM bound: 8 B/voxel (fp32 pred) or 5 B/voxel (uint8)
//   fcd_component_overlap : for a component-id volume cc (ids 1..max_id, 0 = background; e.g. out_lab of
//                           fcd_post_process) and a ground-truth label: out[0] = number of distinct ids present,
//                           out[1] = number of those with at least one voxel where label != 0, out[2] = voxels whose id
//                           is outside [0, max_id] (must be 0).  evaluate_fp = out[0] - out[1].
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kVoxPerBlock = kThreads * 4 * 8;       // 8 vectors of 4 voxels per thread

template <bool U8>
__global__ void __launch_bounds__(kThreads) confusion_kernel(const float* __restrict__ pred_f,
                                                             const unsigned char* __restrict__ pred_u8,
                                                             const float* __restrict__ label, float thr_p, float thr_l,
                                                             long long n, unsigned long long* __restrict__ counts) {
    const int item = blockIdx.y;
    const float* lab = label + (long long)item * n;
    const float* pf = U8 ? nullptr : pred_f + (long long)item * n;
    const unsigned char* pu = U8 ? pred_u8 + (long long)item * n : nullptr;
    const long long base = (long long)blockIdx.x * kVoxPerBlock;
    // vector path only when the item's rows start 16-byte aligned (n % 4 == 0 keeps every item aligned)
    const bool vec = (n & 3) == 0 && ((reinterpret_cast<uintptr_t>(label) & 15) == 0) &&
                     (U8 ? (reinterpret_cast<uintptr_t>(pred_u8) & 3) == 0 : (reinterpret_cast<uintptr_t>(pred_f) & 15) == 0);
    int tp = 0, fp = 0, fn = 0, tot = 0;
#pragma unroll 4
    for (int k = 0; k < 8; ++k) {
        const long long v0 = base + ((long long)k * kThreads + threadIdx.x) * 4;
        if (v0 >= n) break;
        float l4[4], p4[4];
        int cnt = 4;
        if (vec) {
            const float4 l = __ldg(reinterpret_cast<const float4*>(lab + v0));
            l4[0] = l.x; l4[1] = l.y; l4[2] = l.z; l4[3] = l.w;
            if (U8) {
                const uchar4 p = __ldg(reinterpret_cast<const uchar4*>(pu + v0));
                p4[0] = p.x; p4[1] = p.y; p4[2] = p.z; p4[3] = p.w;
            } else {
                const float4 p = __ldg(reinterpret_cast<const float4*>(pf + v0));
                p4[0] = p.x; p4[1] = p.y; p4[2] = p.z; p4[3] = p.w;
            }
        } else {
            cnt = (int)min(4LL, n - v0);
            for (int i = 0; i < 4; ++i) {
                l4[i] = i < cnt ? lab[v0 + i] : 0.f;
                p4[i] = i < cnt ? (U8 ? (float)pu[v0 + i] : pf[v0 + i]) : 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i < cnt) {
                const bool p = p4[i] > thr_p, t = l4[i] > thr_l;
                tp += p && t; fp += p && !t; fn += !p && t;
            }
        }
        tot += cnt;
    }
    int vals[4] = {tp, fp, tot - tp - fp - fn, fn};
    __shared__ int sh[kThreads / 32][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vals[j] += __shfl_xor_sync(0xffffffffu, vals[j], o);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) sh[threadIdx.x >> 5][j] = vals[j];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        int t = 0;
        for (int w = 0; w < kThreads / 32; ++w) t += sh[w][threadIdx.x];
        if (t) atomicAdd(counts + (long long)item * 4 + threadIdx.x, (unsigned long long)t);   // integer: order-free
    }
}

// flags: [0, max_id] present, [max_id + 1, 2 max_id + 1] overlapping the label
__global__ void __launch_bounds__(kThreads) component_mark_kernel(const float* __restrict__ cc,
                                                                  const float* __restrict__ label, long long V,
                                                                  long long max_id, unsigned char* __restrict__ flags,
                                                                  unsigned long long* __restrict__ out) {
    int bad = 0;
    auto mark = [&](float c, float l) {
        if (c == 0.f) return;
        const long long id = (long long)c;
        if (id < 1 || id > max_id || (float)id != c) { ++bad; return; }
        // same-value byte stores from many threads: benign
        if (flags[id] == 0) flags[id] = 1;
        if (l != 0.f && flags[max_id + 1 + id] == 0) flags[max_id + 1 + id] = 1;
    };
    if ((V & 3) == 0 && ((reinterpret_cast<uintptr_t>(cc) | reinterpret_cast<uintptr_t>(label)) & 15) == 0) {
        const long long V4 = V >> 2, stride = (long long)gridDim.x * kThreads;
        const float4* c4 = reinterpret_cast<const float4*>(cc);
        const float4* l4 = reinterpret_cast<const float4*>(label);
        for (long long v = (long long)blockIdx.x * kThreads + threadIdx.x; v < V4; v += 2 * stride) {
            // two independent 16-byte loads per stream in flight
            const bool two = v + stride < V4;
            const float4 ca = __ldg(c4 + v), la = __ldg(l4 + v);
            const float4 cb = two ? __ldg(c4 + v + stride) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 lb = two ? __ldg(l4 + v + stride) : make_float4(0.f, 0.f, 0.f, 0.f);
            mark(ca.x, la.x); mark(ca.y, la.y); mark(ca.z, la.z); mark(ca.w, la.w);
            mark(cb.x, lb.x); mark(cb.y, lb.y); mark(cb.z, lb.z); mark(cb.w, lb.w);
        }
    } else {
        const long long stride = (long long)gridDim.x * kThreads;
        for (long long v = (long long)blockIdx.x * kThreads + threadIdx.x; v < V; v += stride) mark(cc[v], label[v]);
    }
    if (bad) atomicAdd(out + 2, (unsigned long long)bad);
}

__global__ void __launch_bounds__(kThreads) component_count_kernel(const unsigned char* __restrict__ flags,
                                                                   long long max_id,
                                                                   unsigned long long* __restrict__ out) {
    const long long stride = (long long)gridDim.x * kThreads;
    int present = 0, hit = 0;
    for (long long id = 1 + (long long)blockIdx.x * kThreads + threadIdx.x; id <= max_id; id += stride) {
        present += flags[id];
        hit += flags[max_id + 1 + id];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        present += __shfl_xor_sync(0xffffffffu, present, o);
        hit += __shfl_xor_sync(0xffffffffu, hit, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (present) atomicAdd(out, (unsigned long long)present);
        if (hit) atomicAdd(out + 1, (unsigned long long)hit);
    }
}

}  // namespace

FCD_API int fcd_confusion_counts(const float* pred_f, const void* pred_u8, const float* label, float thr_pred,
                                 float thr_label, long long n, int items, long long* counts, cudaStream_t st) {
    if (n < 1 || items < 1 || items > 65535 || ((pred_f == nullptr) == (pred_u8 == nullptr))) return -1;
    const long long nblk = (n + kVoxPerBlock - 1) / kVoxPerBlock;
    if (nblk > 0x7fffffffLL) return -1;
    cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(long long) * 4 * (size_t)items, st);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)nblk, (unsigned)items);
    auto* c = reinterpret_cast<unsigned long long*>(counts);
    if (pred_u8)
        confusion_kernel<true><<<grid, kThreads, 0, st>>>(nullptr, (const unsigned char*)pred_u8, label, thr_pred,
                                                          thr_label, n, c);
    else
        confusion_kernel<false><<<grid, kThreads, 0, st>>>(pred_f, nullptr, label, thr_pred, thr_label, n, c);
    return (int)cudaGetLastError();
}

FCD_API long long fcd_component_overlap_ws_bytes(long long max_id) { return max_id < 1 ? -1 : 2 * (max_id + 1); }

FCD_API int fcd_component_overlap(const float* cc, const float* label, long long V, long long max_id, void* ws,
                                  long long ws_bytes, long long* out, cudaStream_t st) {
    if (V < 1 || max_id < 1 || ws_bytes < 2 * (max_id + 1)) return -1;
    cudaError_t e = cudaMemsetAsync(ws, 0, (size_t)(2 * (max_id + 1)), st);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(out, 0, sizeof(long long) * 3, st);
    if (e != cudaSuccess) return (int)e;
    auto* o = reinterpret_cast<unsigned long long*>(out);
    const int g1 = (int)min((V + kThreads - 1) / kThreads, 148LL * 16);
    component_mark_kernel<<<g1, kThreads, 0, st>>>(cc, label, V, max_id, (unsigned char*)ws, o);
    const int g2 = (int)min((max_id + kThreads - 1) / kThreads, 148LL * 8);
    component_count_kernel<<<g2, kThreads, 0, st>>>((const unsigned char*)ws, max_id, o);
    return (int)cudaGetLastError();
}
