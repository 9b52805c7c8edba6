// kmeans.cu — sem_cluster_by's core: faiss.Kmeans(d, k, niter).train(x) followed by kmeans.index.search(x, 1)
// (lotus/utils.py:61-65), restated from faiss/Clustering.cpp exactly as oracle/faiss_flat.c `orc_kmeans` does:
//   host control flow : subsample to 256*k points with rand_perm(seed) (std::mt19937), initial centroids = first k of
//                       rand_perm(seed+1), niter Lloyd iterations, split_clusters with RandomGenerator(1234)
//   assignment        : L2 top-1 of every point against the k centroids through the SAME exact pipeline as search
//                       (tcgen05 filter -> canonical re-score -> certificate); argmin ties -> lowest centroid id
//   centroid update   : faiss sums the member points in POINT ORDER in fp32 and scales by 1/count. That order is
//                       kept: a stable counting sort builds per-centroid member lists in point order, then one thread
//                       per (centroid, dimension) adds its column sequentially -> bit-identical centroids.
#include <algorithm>
#include <chrono>
#include <random>
#include <vector>

#include "index.cuh"

namespace b2 {
namespace {

constexpr unsigned FULL = 0xffffffffu;

// faiss/utils/random.cpp rand_perm: Fisher-Yates with rng.rand_int(n - i) = mt() % (n - i)
void rand_perm(std::vector<int64_t>& perm, int64_t n, int64_t seed) {
    perm.resize(n);
    for (int64_t i = 0; i < n; ++i) perm[i] = i;
    std::mt19937 mt((unsigned int)seed);
    for (int64_t i = 0; i + 1 < n; ++i) {
        const int64_t i2 = i + (int64_t)(mt() % (uint32_t)(n - i));
        std::swap(perm[i], perm[i2]);
    }
}

__global__ void rows_to_f32_kernel(const void* x, int dtype, int d, const int64_t* ids, int64_t m, float* out) {
    const int64_t total = m * d;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / d;
        const int c = (int)(t - r * d);
        const int64_t src = ids ? ids[r] : r;
        out[t] = dtype == B2_F32 ? reinterpret_cast<const float*>(x)[src * d + c]
                                 : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[src * d + c]);
    }
}

__global__ void sum_f32_kernel(const float* v, int64_t n, double* out) {
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc += (double)v[i];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(FULL, acc, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// ---- stable counting sort of points by centroid (member lists in point order) ---------------------------------------
// block b (ONE warp) owns the contiguous point range [b*L, (b+1)*L)
__global__ void km_count_kernel(const int64_t* assign, int64_t n, int64_t L, int k, int32_t* cnt) {
    extern __shared__ int32_t s_cnt[];
    for (int c = threadIdx.x; c < k; c += 32) s_cnt[c] = 0;
    __syncwarp();
    const int64_t lo = blockIdx.x * L, hi = min(n, lo + L);
    for (int64_t i = lo + threadIdx.x; i < hi; i += 32) atomicAdd(&s_cnt[(int)assign[i]], 1);
    __syncwarp();
    for (int c = threadIdx.x; c < k; c += 32) cnt[(size_t)blockIdx.x * k + c] = s_cnt[c];
}

// per centroid: exclusive scan of the block counts; totals[c] = cluster size
__global__ void km_scan_blocks_kernel(int32_t* cnt, int nb, int k, int32_t* totals) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= k) return;
    int32_t run = 0;
    for (int b = 0; b < nb; ++b) {
        const int32_t t = cnt[(size_t)b * k + c];
        cnt[(size_t)b * k + c] = run;
        run += t;
    }
    totals[c] = run;
}

__global__ void km_scan_totals_kernel(const int32_t* totals, int k, int64_t* offsets) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int64_t run = 0;
        for (int c = 0; c < k; ++c) {
            offsets[c] = run;
            run += totals[c];
        }
        offsets[k] = run;
    }
}

__global__ void km_fill_kernel(const int64_t* assign, int64_t n, int64_t L, int k, const int32_t* blk_start, const int64_t* offsets,
                               int32_t* members) {
    extern __shared__ int32_t s_run[];  // next free slot (relative to the cluster's list) for this block's points
    const int lane = threadIdx.x;
    for (int c = lane; c < k; c += 32) s_run[c] = blk_start[(size_t)blockIdx.x * k + c];
    __syncwarp();
    const int64_t lo = blockIdx.x * L, hi = min(n, lo + L);
    for (int64_t base = lo; base < hi; base += 32) {
        const int64_t i = base + lane;
        const bool active = i < hi;
        const unsigned amask = __ballot_sync(FULL, active);
        if (active) {
            const int a = (int)assign[i];
            const unsigned same = __match_any_sync(amask, a);
            const int rank = __popc(same & ((1u << lane) - 1));
            const int32_t slot = s_run[a] + rank;
            members[offsets[a] + slot] = (int32_t)i;
            __syncwarp(amask);
            if (rank == 0) s_run[a] += __popc(same);
        }
        __syncwarp();
    }
}

// thread (c, j): centroid[c][j] = (sum over members of c, in point order, of x[p][j]) * (1 / count)   [all fp32]
__global__ void km_accumulate_kernel(const void* x, int dtype, int d, const int64_t* ids, const int32_t* members,
                                     const int64_t* offsets, float* centroids, float* hassign, int normalize) {
    const int c = blockIdx.x;
    const int64_t o0 = offsets[c], o1 = offsets[c + 1];
    const float cntf = (float)(o1 - o0);
    if (threadIdx.x == 0 && blockIdx.y == 0) hassign[c] = cntf;
    const int j = blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= d) return;
    float acc = 0.f;
    int64_t o = o0;
    for (; o + 4 <= o1; o += 4) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t p = members[o + u];
            const int64_t r = ids ? ids[p] : p;
            v[u] = dtype == B2_F32 ? reinterpret_cast<const float*>(x)[r * d + j]
                                   : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[r * d + j]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc = __fadd_rn(acc, v[u]);
    }
    for (; o < o1; ++o) {
        const int64_t p = members[o];
        const int64_t r = ids ? ids[p] : p;
        const float v = dtype == B2_F32 ? reinterpret_cast<const float*>(x)[r * d + j]
                                        : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[r * d + j]);
        acc = __fadd_rn(acc, v);
    }
    if (normalize && o1 > o0) {
        const float norm = __fdiv_rn(1.0f, cntf);
        acc = __fmul_rn(acc, norm);
    }
    centroids[(size_t)c * d + j] = acc;
}

// faiss/Clustering.cpp split_clusters on the host copy (EPS = 1/1024, RandomGenerator rng(1234))
int split_clusters_host(int d, int k, int64_t n, std::vector<float>& hassign, std::vector<float>& centroids) {
    const double EPS = 1 / 1024.;
    int nsplit = 0;
    std::mt19937 mt(1234u);
    for (int ci = 0; ci < k; ci++) {
        if (hassign[ci] == 0) {
            int cj;
            for (cj = 0; true; cj = (cj + 1) % k) {
                const float p = (hassign[cj] - 1.0) / (float)(n - k);
                const float r = mt() / float(mt.max());
                if (r < p) break;
            }
            memcpy(centroids.data() + (size_t)ci * d, centroids.data() + (size_t)cj * d, sizeof(float) * d);
            for (int j = 0; j < d; j++) {
                if (j % 2 == 0) {
                    centroids[(size_t)ci * d + j] *= 1 + EPS;
                    centroids[(size_t)cj * d + j] *= 1 - EPS;
                } else {
                    centroids[(size_t)ci * d + j] *= 1 - EPS;
                    centroids[(size_t)cj * d + j] *= 1 + EPS;
                }
            }
            hassign[ci] = hassign[cj] / 2;
            hassign[cj] -= hassign[ci];
            nsplit++;
        }
    }
    return nsplit;
}

struct KmWork {
    DevBuf cent, cent_filt, cent_norm2, scalar, dis, assign, members, offsets, totals, blk, hassign, ids, train, obj;
    void release() {
        DevBuf* all[] = {&cent, &cent_filt, &cent_norm2, &scalar, &dis, &assign, &members, &offsets, &totals, &blk, &hassign, &ids, &train, &obj};
        for (DevBuf* b : all) b->release();
    }
};

// searchable view of the fp32 centroid matrix; the filter operand matches the point dtype (bf16 points -> bf16 copy)
int centroid_view(const float* cent, int k, int d, int point_dtype, KmWork& w, MatView& v, cudaStream_t st) {
    v.store = cent;
    v.n = k;
    v.d = d;
    v.dtype = B2_F32;
    if (point_dtype == B2_BF16) {
        v.filt_dtype = B2_BF16;
        v.filt_pitch = round_up(d, 8);
        B2_TRY(w.cent_filt.ensure((size_t)k * v.filt_pitch * 2));
        B2_TRY(launch_convert_pad(cent, B2_F32, k, d, w.cent_filt.p, B2_BF16, v.filt_pitch, st));
        v.filt = w.cent_filt.p;
    } else {
        v.filt_dtype = B2_F32;
        if (d % 4 == 0) {
            v.filt = cent;
            v.filt_pitch = d;
        } else {
            v.filt_pitch = round_up(d, 4);
            B2_TRY(w.cent_filt.ensure((size_t)k * v.filt_pitch * 4));
            B2_TRY(launch_convert_pad(cent, B2_F32, k, d, w.cent_filt.p, B2_F32, v.filt_pitch, st));
            v.filt = w.cent_filt.p;
        }
    }
    B2_TRY(w.cent_norm2.ensure((size_t)k * sizeof(float)));
    B2_TRY(w.scalar.ensure(64));
    B2_TRY(launch_row_norms(cent, B2_F32, k, d, w.cent_norm2.as<float>(), w.scalar.as<float>(), st));
    float mx = 0.f;
    B2_CUDA(cudaMemcpyAsync(&mx, w.scalar.p, sizeof(float), cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    v.norm2 = w.cent_norm2.as<float>();
    v.max_norm = mx;
    return B2_OK;
}

// assign[m], dis[m] for the rows pts[m,d] (device, index dtype) against the centroids
int assign_points(b2_index* idx, const void* pts, int64_t m, const float* cent, int k, KmWork& w, float* dis, int64_t* assign,
                  cudaStream_t st) {
    MatView cv;
    B2_TRY(centroid_view(cent, k, idx->d, idx->dtype, w, cv, st));
    return search_core(idx, cv, B2_METRIC_L2, pts, idx->dtype, m, 1, nullptr, 0, dis, assign, st);
}

int update_centroids(b2_index* idx, const void* x, const int64_t* row_ids, int64_t n, const int64_t* assign, int k, KmWork& w,
                     float* cent, cudaStream_t st, int normalize = 1) {
    const int d = idx->d;
    // one warp per block walks its point range in order; 1024 ranges keep the per-centroid scan over blocks short
    int64_t nb = std::min<int64_t>(1024, std::max<int64_t>(1, ceil_div(n, 256)));
    while (nb > 1 && nb * (int64_t)k > ((int64_t)1 << 26)) nb /= 2;
    const int64_t L = ceil_div(n, nb);
    nb = ceil_div(n, L);
    const size_t smem = (size_t)k * sizeof(int32_t);
    if (smem > 200 * 1024) {
        set_error("k=%d centroids exceed the shared-memory budget of the member-list kernels", k);
        return B2_ERANGE;
    }
    if (smem > 48 * 1024) {
        B2_CUDA(cudaFuncSetAttribute(km_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        B2_CUDA(cudaFuncSetAttribute(km_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    B2_TRY(w.blk.ensure((size_t)nb * k * sizeof(int32_t)));
    B2_TRY(w.totals.ensure((size_t)k * sizeof(int32_t)));
    B2_TRY(w.offsets.ensure((size_t)(k + 1) * sizeof(int64_t)));
    B2_TRY(w.members.ensure((size_t)std::max<int64_t>(n, 1) * sizeof(int32_t)));
    B2_TRY(w.hassign.ensure((size_t)k * sizeof(float)));
    km_count_kernel<<<(unsigned)nb, 32, smem, st>>>(assign, n, L, k, w.blk.as<int32_t>());
    B2_LAUNCH_CHECK();
    km_scan_blocks_kernel<<<(unsigned)ceil_div(k, 128), 128, 0, st>>>(w.blk.as<int32_t>(), (int)nb, k, w.totals.as<int32_t>());
    B2_LAUNCH_CHECK();
    km_scan_totals_kernel<<<1, 32, 0, st>>>(w.totals.as<int32_t>(), k, w.offsets.as<int64_t>());
    B2_LAUNCH_CHECK();
    km_fill_kernel<<<(unsigned)nb, 32, smem, st>>>(assign, n, L, k, w.blk.as<int32_t>(), w.offsets.as<int64_t>(), w.members.as<int32_t>());
    B2_LAUNCH_CHECK();
    dim3 grid((unsigned)k, (unsigned)ceil_div(d, 128));
    km_accumulate_kernel<<<grid, 128, 0, st>>>(x, idx->dtype, d, row_ids, w.members.as<int32_t>(), w.offsets.as<int64_t>(), cent,
                                               w.hassign.as<float>(), normalize);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

// B2_KM_TIMING=1: per-phase wall clock of b2_kmeans on stderr (each phase closed by a stream synchronise, so the phases
// no longer overlap with the host work between them: a diagnostic, not a benchmark mode)
struct KmTimer {
    bool on;
    cudaStream_t st;
    std::chrono::steady_clock::time_point t0;
    double acc[6] = {0, 0, 0, 0, 0, 0};  // setup, assign, assign:filter, update, split/host, final
    KmTimer(cudaStream_t s) : st(s) {
        const char* e = getenv("B2_KM_TIMING");
        on = e && atoi(e) != 0;
        t0 = std::chrono::steady_clock::now();
    }
    void lap(int slot) {
        if (!on) return;
        cudaStreamSynchronize(st);
        const auto t1 = std::chrono::steady_clock::now();
        acc[slot] += std::chrono::duration<double, std::milli>(t1 - t0).count();
        t0 = t1;
    }
    void report(int64_t m, int64_t nx, int k, int d, int niter) const {
        if (!on) return;
        fprintf(stderr,
                "[b2 kmeans timing] m=%lld train=%lld k=%d d=%d niter=%d | setup %.2f ms | %d x assign %.2f ms (filter kernel %.2f) | "
                "%d x update %.2f ms | obj/split host %.2f ms | final assign %.2f ms\n",
                (long long)m, (long long)nx, k, d, niter, acc[0], niter, acc[1], acc[2], niter, acc[3], acc[4], acc[5]);
    }
};

int kmeans_impl(b2_index* idx, const int64_t* ids_host, int64_t m, int k, int niter, int64_t seed, int full_lloyd, int64_t* out_assign,
                float* out_centroids, float* out_obj, KmWork& w) {
    const int d = idx->d;
    cudaStream_t st = idx->stream;
    KmTimer tm(st);
    const size_t es = esize(idx->dtype);
    // the point set: rows ids[0..m) of the index (device id list), or all rows
    const int64_t* ids_dev = nullptr;
    if (ids_host) {
        for (int64_t i = 0; i < m; ++i)
            if (ids_host[i] < 0 || ids_host[i] >= idx->n) {
                set_error("ids contains a position outside [0, %lld)", (long long)idx->n);
                return B2_ERANGE;
            }
        B2_TRY(w.ids.ensure((size_t)std::max<int64_t>(m, 1) * sizeof(int64_t)));
        B2_CUDA(cudaMemcpyAsync(w.ids.p, ids_host, (size_t)m * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        ids_dev = w.ids.as<int64_t>();
    }
    // materialise the point matrix P[m,d] (index dtype) when it is not simply the whole index
    const void* P = idx->store.p;
    DevBuf pts;
    struct Guard { DevBuf& b; ~Guard() { b.release(); } } pts_guard{pts};
    if (ids_dev) {
        B2_TRY(pts.ensure((size_t)std::max<int64_t>(m, 1) * d * es));
        B2_TRY(idx->scalar.ensure(64));
        int* err = reinterpret_cast<int*>(idx->scalar.as<char>() + 16);
        B2_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
        B2_TRY(launch_gather_rows(idx->store.p, idx->dtype, d, ids_dev, m, idx->n, pts.p, err, st));
        P = pts.p;
    }
    // training set (faiss ClusteringParameters: max_points_per_centroid = 256)
    int64_t nx = m;
    const void* T = P;
    const int64_t max_pts = (int64_t)k * 256;
    if (!full_lloyd && nx > max_pts) {
        std::vector<int64_t> perm;
        rand_perm(perm, nx, seed);
        perm.resize(max_pts);
        nx = max_pts;
        DevBuf perm_dev;
        int rc = perm_dev.ensure((size_t)nx * sizeof(int64_t));
        if (rc == B2_OK) rc = w.train.ensure((size_t)nx * d * es);
        if (rc == B2_OK) {
            cudaMemcpyAsync(perm_dev.p, perm.data(), (size_t)nx * sizeof(int64_t), cudaMemcpyHostToDevice, st);
            int* err = reinterpret_cast<int*>(idx->scalar.as<char>() + 16);
            rc = launch_gather_rows(P, idx->dtype, d, perm_dev.as<int64_t>(), nx, m, w.train.p, err, st);
            cudaStreamSynchronize(st);
        }
        perm_dev.release();
        if (rc != B2_OK) return rc;
        T = w.train.p;
    }
    B2_TRY(w.cent.ensure((size_t)k * d * sizeof(float)));
    float* cent = w.cent.as<float>();
    std::vector<float> h_obj(std::max(niter, 1), 0.f);
    if (nx == k) {
        // "Number of training points same as number of centroids, just copying"
        rows_to_f32_kernel<<<148, 256, 0, st>>>(T, idx->dtype, d, nullptr, k, cent);
        B2_LAUNCH_CHECK();
    } else {
        std::vector<int64_t> perm;
        rand_perm(perm, nx, seed + 1);
        DevBuf perm_dev;
        int rc = perm_dev.ensure((size_t)k * sizeof(int64_t));
        if (rc != B2_OK) return rc;
        cudaMemcpyAsync(perm_dev.p, perm.data(), (size_t)k * sizeof(int64_t), cudaMemcpyHostToDevice, st);
        rows_to_f32_kernel<<<148, 256, 0, st>>>(T, idx->dtype, d, perm_dev.as<int64_t>(), k, cent);
        g_stats[ST_LAUNCHES]++;
        cudaStreamSynchronize(st);
        perm_dev.release();
        B2_TRY(w.dis.ensure((size_t)nx * sizeof(float)));
        B2_TRY(w.assign.ensure((size_t)nx * sizeof(int64_t)));
        B2_TRY(w.obj.ensure(64));
        std::vector<float> hassign(k), hcent;
        tm.lap(0);
        for (int it = 0; it < niter; ++it) {
            B2_TRY(assign_points(idx, T, nx, cent, k, w, w.dis.as<float>(), w.assign.as<int64_t>(), st));
            tm.lap(1);
            if (tm.on && idx->last_filter_ms > 0) tm.acc[2] += idx->last_filter_ms;
            B2_CUDA(cudaMemsetAsync(w.obj.p, 0, sizeof(double), st));
            sum_f32_kernel<<<148, 256, 0, st>>>(w.dis.as<float>(), nx, w.obj.as<double>());
            B2_LAUNCH_CHECK();
            B2_TRY(update_centroids(idx, T, nullptr, nx, w.assign.as<int64_t>(), k, w, cent, st));
            tm.lap(3);
            double obj = 0;
            B2_CUDA(cudaMemcpyAsync(&obj, w.obj.p, sizeof(double), cudaMemcpyDeviceToHost, st));
            B2_CUDA(cudaMemcpyAsync(hassign.data(), w.hassign.p, (size_t)k * sizeof(float), cudaMemcpyDeviceToHost, st));
            B2_CUDA(cudaStreamSynchronize(st));
            h_obj[it] = (float)obj;
            bool any_empty = false;
            for (int c = 0; c < k; ++c) any_empty |= hassign[c] == 0;
            if (any_empty) {
                hcent.resize((size_t)k * d);
                B2_CUDA(cudaMemcpy(hcent.data(), cent, (size_t)k * d * sizeof(float), cudaMemcpyDeviceToHost));
                split_clusters_host(d, k, nx, hassign, hcent);
                B2_CUDA(cudaMemcpy(cent, hcent.data(), (size_t)k * d * sizeof(float), cudaMemcpyHostToDevice));
            }
            tm.lap(4);
        }
    }
    tm.lap(0);
    // lotus/utils.py:65 kmeans.index.search(vec_set, 1) over ALL m points
    DevBuf fin_dis, fin_assign;
    int rc = fin_dis.ensure((size_t)std::max<int64_t>(m, 1) * sizeof(float));
    if (rc == B2_OK) rc = fin_assign.ensure((size_t)std::max<int64_t>(m, 1) * sizeof(int64_t));
    if (rc == B2_OK) rc = assign_points(idx, P, m, cent, k, w, fin_dis.as<float>(), fin_assign.as<int64_t>(), st);
    if (rc == B2_OK) {
        cudaMemcpyAsync(out_assign, fin_assign.p, (size_t)m * sizeof(int64_t), cudaMemcpyDeviceToHost, st);
        if (out_centroids) cudaMemcpyAsync(out_centroids, cent, (size_t)k * d * sizeof(float), cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) {
            set_error("k-means failed on the device: %s", cudaGetErrorString(cudaGetLastError()));
            rc = B2_ECUDA;
        }
    }
    tm.lap(5);
    tm.report(m, nx, k, d, niter);
    fin_dis.release();
    fin_assign.release();
    if (rc == B2_OK && out_obj)
        for (int it = 0; it < niter; ++it) out_obj[it] = h_obj[it];
    return rc;
}

}  // namespace
}  // namespace b2

using namespace b2;

extern "C" {

int b2_kmeans(b2_index* idx, const int64_t* ids, int64_t m, int32_t k, int32_t niter, int64_t seed, int32_t full_lloyd,
              int64_t* out_assign, float* out_centroids, float* out_obj) {
    if (!idx) { set_error("Index not loaded"); return B2_EINVAL; }
    if (!ids) m = idx->n;
    if (k <= 0 || niter < 0 || m < 0 || !out_assign) { set_error("bad k-means arguments (k=%d niter=%d m=%lld)", k, niter, (long long)m); return B2_EINVAL; }
    if (m < k) { set_error("Number of training points (%lld) should be at least as large as number of clusters (%d)", (long long)m, k); return B2_EINVAL; }
    DeviceGuard guard(idx->device);
    KmWork w;
    const int rc = kmeans_impl(idx, ids, m, k, niter, seed, full_lloyd, out_assign, out_centroids, out_obj, w);
    cudaStreamSynchronize(idx->stream);
    w.release();
    return rc;
}

int b2_kmeans_accumulate(b2_index* idx, const int64_t* ids, int64_t m, const int64_t* assign, int32_t k, float* out_sums,
                         float* out_counts) {
    if (!idx) { set_error("Index not loaded"); return B2_EINVAL; }
    if (!ids) m = idx->n;
    if (k <= 0 || m < 0 || !assign || !out_sums || !out_counts) { set_error("bad arguments"); return B2_EINVAL; }
    for (int64_t i = 0; i < m; ++i) {
        if (assign[i] < 0 || assign[i] >= k) { set_error("assign[%lld] = %lld outside [0, %d)", (long long)i, (long long)assign[i], k); return B2_ERANGE; }
        if (ids && (ids[i] < 0 || ids[i] >= idx->n)) { set_error("ids contains a position outside [0, %lld)", (long long)idx->n); return B2_ERANGE; }
    }
    DeviceGuard guard(idx->device);
    cudaStream_t st = idx->stream;
    KmWork w;
    DevBuf d_assign;
    int rc = w.cent.ensure((size_t)k * idx->d * sizeof(float));
    if (rc == B2_OK) rc = d_assign.ensure((size_t)std::max<int64_t>(m, 1) * sizeof(int64_t));
    const int64_t* ids_dev = nullptr;
    if (rc == B2_OK && ids) {
        rc = w.ids.ensure((size_t)std::max<int64_t>(m, 1) * sizeof(int64_t));
        if (rc == B2_OK) {
            cudaMemcpyAsync(w.ids.p, ids, (size_t)m * sizeof(int64_t), cudaMemcpyHostToDevice, st);
            ids_dev = w.ids.as<int64_t>();
        }
    }
    if (rc == B2_OK) {
        cudaMemcpyAsync(d_assign.p, assign, (size_t)m * sizeof(int64_t), cudaMemcpyHostToDevice, st);
        rc = update_centroids(idx, idx->store.p, ids_dev, m, d_assign.as<int64_t>(), k, w, w.cent.as<float>(), st, /*normalize=*/0);
    }
    if (rc == B2_OK) {
        cudaMemcpyAsync(out_sums, w.cent.p, (size_t)k * idx->d * sizeof(float), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(out_counts, w.hassign.p, (size_t)k * sizeof(float), cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) {
            set_error("k-means accumulate failed on the device: %s", cudaGetErrorString(cudaGetLastError()));
            rc = B2_ECUDA;
        }
    } else {
        cudaStreamSynchronize(st);
    }
    w.release();
    d_assign.release();
    return rc;
}

int b2_kmeans_assign(b2_index* idx, const int64_t* ids, int64_t m, const float* centroids, int32_t k, int64_t* out_assign,
                     float* out_dist) {
    if (!idx) { set_error("Index not loaded"); return B2_EINVAL; }
    if (!ids) m = idx->n;
    if (k <= 0 || m < 0 || !centroids || (m > 0 && !out_assign)) { set_error("bad arguments"); return B2_EINVAL; }
    if (m == 0) return B2_OK;
    DeviceGuard guard(idx->device);
    cudaStream_t st = idx->stream;
    const int d = idx->d;
    KmWork w;
    DevBuf pts, dis, asg;
    auto cleanup = [&]() { cudaStreamSynchronize(st); w.release(); pts.release(); dis.release(); asg.release(); };
    int rc = w.cent.ensure((size_t)k * d * sizeof(float));
    if (rc == B2_OK) rc = dis.ensure((size_t)m * sizeof(float));
    if (rc == B2_OK) rc = asg.ensure((size_t)m * sizeof(int64_t));
    const void* P = idx->store.p;
    if (rc == B2_OK && ids) {
        rc = w.ids.ensure((size_t)m * sizeof(int64_t));
        if (rc == B2_OK) rc = pts.ensure((size_t)m * d * esize(idx->dtype));
        if (rc == B2_OK) rc = idx->scalar.ensure(64);
        if (rc == B2_OK) {
            int* err = reinterpret_cast<int*>(idx->scalar.as<char>() + 16);
            cudaMemsetAsync(err, 0, sizeof(int), st);
            cudaMemcpyAsync(w.ids.p, ids, (size_t)m * sizeof(int64_t), cudaMemcpyHostToDevice, st);
            rc = launch_gather_rows(idx->store.p, idx->dtype, d, w.ids.as<int64_t>(), m, idx->n, pts.p, err, st);
            int herr = 0;
            cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, st);
            cudaStreamSynchronize(st);
            if (rc == B2_OK && herr) { set_error("ids contains a position outside [0, %lld)", (long long)idx->n); rc = B2_ERANGE; }
            P = pts.p;
        }
    }
    if (rc == B2_OK) {
        cudaMemcpyAsync(w.cent.p, centroids, (size_t)k * d * sizeof(float), cudaMemcpyHostToDevice, st);
        rc = assign_points(idx, P, m, w.cent.as<float>(), k, w, dis.as<float>(), asg.as<int64_t>(), st);
    }
    if (rc == B2_OK) {
        cudaMemcpyAsync(out_assign, asg.p, (size_t)m * sizeof(int64_t), cudaMemcpyDeviceToHost, st);
        if (out_dist) cudaMemcpyAsync(out_dist, dis.p, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) {
            set_error("k-means assignment failed on the device: %s", cudaGetErrorString(cudaGetLastError()));
            rc = B2_ECUDA;
        }
    }
    cleanup();
    return rc;
}

}  // extern "C"
