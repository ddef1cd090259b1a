// Retrieval metrics of the evaluation step: reid/evaluation_metrics/ranking.py:18-115 (cmc, mean_ap)
// as called by reid/evaluators.py:88-133 -- without sorting the q x g distance matrix.
//
// Both metrics only need, for every MATCH g of query i (gallery entry with the query's id that is not filtered out), how
// many valid gallery entries are ranked before it:
//   * CMC (ranking.py:43-75): a match at position k among the valid entries, being the j-th match, adds to ret[k - j];
//     k - j = number of valid NON-matching entries ranked before it.  Ranking order = (distance, gallery index), i.e. a
//     stable argsort (the reference's np.argsort leaves ties unspecified).
//   * AP (ranking.py:105-111 -> sklearn.metrics.average_precision_score: one threshold per DISTINCT score, ties form one
//     threshold): AP = (1/|M|) * sum over matches g of TP(d_g) / N(d_g), TP(v) = #{matches with d <= v},
//     N(v) = #{valid entries with d <= v}.
// "valid" (ranking.py:48-49, 103-104): not (same id AND same camera); with separate_camera_set also not same camera.
// One CTA per query: pass 1 classifies the gallery once (one code byte per entry in shared memory: bit 0 valid, bit 1
// same id) and collects the matches of the row (ballot-compacted, ascending gallery index); pass 2 gives every warp one
// match at a time and counts over the row (coalesced, L1/L2-resident after the first sweep); integer warp reductions,
// the per-query AP is summed in a fixed order in float64.
#include <limits.h>

#include "common.cuh"
#include "kernels.h"

namespace ssg {

constexpr int RM_NT = 256;
constexpr int RM_CAP = SSG_RANK_MAX_MATCHES;       // matches kept per query (more -> nmatch = -1 and the flag)
constexpr size_t RM_MAX_CODE_BYTES = 200 * 1024;   // gallery entries classified in shared memory (MSMT17: 82 161)

template <typename T>
__global__ void __launch_bounds__(RM_NT)
rank_metrics_kernel(const T* __restrict__ dist, int m, int n, const long long* __restrict__ q_ids,
                    const long long* __restrict__ g_ids, const long long* __restrict__ q_cams,
                    const long long* __restrict__ g_cams, int separate_camera_set, double* __restrict__ ap,
                    int* __restrict__ nmatch, int* __restrict__ slots, int* __restrict__ flags) {
    extern __shared__ unsigned char s_code[];             // [n] bit 0: valid, bit 1: same id
    __shared__ int s_idx[RM_CAP];
    __shared__ int s_base;
    __shared__ int wsum[RM_NT / 32];
    __shared__ double s_ap[RM_NT / 32];
    const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const T* row = dist + (size_t)i * n;
    const long long qid = q_ids[i], qcam = q_cams[i];
    if (tid == 0) s_base = 0;
    __syncthreads();
    // pass 1: the matches of this query in ascending gallery index
    for (int j0 = 0; j0 < n; j0 += RM_NT) {
        const int j = j0 + tid;
        bool is_match = false;
        if (j < n) {
            const bool same_id = g_ids[j] == qid, same_cam = g_cams[j] == qcam;
            const bool valid = !(same_id && same_cam) && !(separate_camera_set && same_cam);
            is_match = valid && same_id;
            s_code[j] = (unsigned char)((valid ? 1 : 0) | (same_id ? 2 : 0));
        }
        const unsigned mask = __ballot_sync(0xffffffffu, is_match);
        if (lane == 0) wsum[wid] = __popc(mask);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < RM_NT / 32; ++w) { if (w < wid) before += wsum[w]; total += wsum[w]; }
        const int base = s_base;
        if (is_match) {
            const int pos = base + before + __popc(mask & ((1u << lane) - 1u));
            if (pos < RM_CAP) s_idx[pos] = j;
        }
        __syncthreads();
        if (tid == 0) s_base = base + total;
        __syncthreads();
    }
    const int nm_all = s_base;
    if (nm_all > RM_CAP) { if (tid == 0) { flags[0] = 1; nmatch[i] = -1; } return; }
    const int nm = nm_all;
    // pass 2: one warp per match
    double ap_part = 0.0;
    for (int s = wid; s < nm; s += RM_NT / 32) {
        const int g = s_idx[s];
        const T dg = row[g];
        int n_le = 0, tp_le = 0, nonmatch_before = 0;
        for (int j = lane; j < n; j += 32) {
            const int code = s_code[j];
            if (!(code & 1)) continue;
            const bool same_id = (code & 2) != 0;
            const T dj = row[j];
            const bool le = dj <= dg;
            n_le += le;
            tp_le += le && same_id;
            nonmatch_before += (!same_id) && (dj < dg || (dj == dg && j < g));
        }
        n_le = warp_sum_i(n_le);
        tp_le = warp_sum_i(tp_le);
        nonmatch_before = warp_sum_i(nonmatch_before);
        if (lane == 0) {
            slots[(size_t)i * RM_CAP + s] = nonmatch_before;
            ap_part += (double)tp_le / (double)n_le;
        }
    }
    if (lane == 0) s_ap[wid] = ap_part;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < RM_NT / 32; ++w) t += s_ap[w];      // fixed order
        ap[i] = nm > 0 ? t / (double)nm : 0.0;
        nmatch[i] = nm;
    }
}

}  // namespace ssg

using namespace ssg;

extern "C" int ssg_rank_metrics(const void* d_dist, int dtype, int m, int n, const long long* d_query_ids,
                                const long long* d_gallery_ids, const long long* d_query_cams,
                                const long long* d_gallery_cams, int separate_camera_set, double* d_ap, int* d_nmatch,
                                int* d_slots, int* d_flags, void* stream) {
    if (!d_dist || !d_query_ids || !d_gallery_ids || !d_query_cams || !d_gallery_cams || !d_ap || !d_nmatch || !d_slots ||
        !d_flags || m <= 0 || n <= 0)
        return ssg_set_error(SSG_ERR_INVALID, "rank_metrics: bad arguments (m=%d, n=%d)", m, n);
    const size_t code_bytes = ((size_t)n + 15) & ~(size_t)15;
    if (code_bytes > RM_MAX_CODE_BYTES)
        return ssg_set_error(SSG_ERR_INVALID, "rank_metrics: gallery of %d entries exceeds the %d this kernel classifies in shared memory",
                             n, (int)RM_MAX_CODE_BYTES);
    cudaStream_t st = (cudaStream_t)stream;
    SSG_CUDA_TRY(cudaMemsetAsync(d_flags, 0, sizeof(int), st));
    SSG_CUDA_TRY(cudaFuncSetAttribute(rank_metrics_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RM_MAX_CODE_BYTES));
    SSG_CUDA_TRY(cudaFuncSetAttribute(rank_metrics_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RM_MAX_CODE_BYTES));
    if (dtype == SSG_F32)
        rank_metrics_kernel<float><<<m, RM_NT, code_bytes, st>>>((const float*)d_dist, m, n, d_query_ids, d_gallery_ids, d_query_cams,
                                                        d_gallery_cams, separate_camera_set, d_ap, d_nmatch, d_slots, d_flags);
    else if (dtype == SSG_F64)
        rank_metrics_kernel<double><<<m, RM_NT, code_bytes, st>>>((const double*)d_dist, m, n, d_query_ids, d_gallery_ids, d_query_cams,
                                                         d_gallery_cams, separate_camera_set, d_ap, d_nmatch, d_slots, d_flags);
    else
        return ssg_set_error(SSG_ERR_INVALID, "rank_metrics: dtype %d", dtype);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
