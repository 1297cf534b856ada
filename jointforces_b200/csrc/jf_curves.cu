// Radial displacement curves on the device: what
// extract_deformation_curve_solver does on the host with numpy after loading a whole solver file
// (/root/reference/jointforces/simulation.py:296-299): for every log-spaced radial bin the median
// of |u|/r_inner over the nodes whose |x|/r_inner lies in [x_i, x_i+1).  Bin membership depends on
// the mesh only, so the caller classifies the nodes once (same float comparisons as numpy) and
// hands over the per-bin node lists; per solve one CTA per (system, bin) computes the norms with
// numpy's operation order (no FMA contraction, IEEE sqrt and divide) and selects the median by
// rank counting over the non-NaN entries -- bit-identical to np.nanmedian on the same displacements.  D2H per load case
// drops from N*24 bytes to n_bins*8.
#include <math.h>

#include "jf_internal.h"

namespace {

__global__ void __launch_bounds__(256) jf_curve_kernel(const int32_t *__restrict__ bin_ptr, const int32_t *__restrict__ bin_nodes,
                                                       const double *__restrict__ U, double *__restrict__ scratch,
                                                       double *__restrict__ out, const jf_sys_state *__restrict__ state,
                                                       long long N, int n_bins, long long scratch_stride, double r_inner) {
    const int bin = blockIdx.x, s = blockIdx.y;
    const int beg = bin_ptr[bin], n = bin_ptr[bin + 1] - beg;
    double *v = scratch + (long long)s * scratch_stride + beg;
    const double *Us = U + (long long)s * N * 3;
    if (n == 0) {
        if (threadIdx.x == 0) out[(long long)s * n_bins + bin] = nan("");  // np.nanmedian of an empty bin
        return;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double *u = Us + 3ll * bin_nodes[beg + i];
        const double q = __dadd_rn(__dadd_rn(__dmul_rn(u[0], u[0]), __dmul_rn(u[1], u[1])), __dmul_rn(u[2], u[2]));
        v[i] = __ddiv_rn(__dsqrt_rn(q), r_inner);
    }
    __syncthreads();
    // np.nanmedian: NaN entries (a diverged solve) are left out; a bin without any finite-or-infinite entry gives NaN.
    // Rank selection over the m non-NaN entries: element i has rank = #{j : v_j < v_i} + #{j < i : v_j == v_i}; the ranks
    // of the non-NaN elements are a permutation of 0..m-1 (comparisons with NaN are false, so NaNs never count).
    __shared__ double mid[2];
    __shared__ int n_valid;
    if (threadIdx.x == 0) {
        mid[0] = mid[1] = nan("");
        n_valid = 0;
    }
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mine += v[i] == v[i];
    if (mine) atomicAdd(&n_valid, mine);
    __syncthreads();
    const int m = n_valid;
    const int k_hi = m / 2, k_lo = (m - 1) / 2;  // middle element(s): equal for odd m
    for (int i = threadIdx.x; i < n && m > 0; i += blockDim.x) {
        const double vi = v[i];
        if (vi != vi) continue;
        int rank = 0;
        for (int j = 0; j < n; j++) {
            const double vj = v[j];
            rank += (vj < vi) || (vj == vi && j < i);
        }
        if (rank == k_lo) mid[0] = vi;
        if (rank == k_hi) mid[1] = vi;
    }
    __syncthreads();
    if (threadIdx.x == 0)
        out[(long long)s * n_bins + bin] = m == 0 ? nan("") : ((k_lo == k_hi) ? mid[0] : __dmul_rn(__dadd_rn(mid[0], mid[1]), 0.5));
}

}  // namespace

extern "C" int jf_set_curve_bins(jf_ctx *c, int n_bins, const int32_t *bin_ptr, const int32_t *bin_nodes, double r_inner) {
    if (!c || !bin_ptr || n_bins < 1 || !(r_inner > 0)) {
        jf_set_error("jf_set_curve_bins: bad arguments");
        return JF_ERR_ARG;
    }
    JF_CUDA(cudaSetDevice(c->device));
    const int32_t total = bin_ptr[n_bins];
    if (bin_ptr[0] != 0 || total < 0 || (total > 0 && !bin_nodes)) {
        jf_set_error("jf_set_curve_bins: malformed bin lists");
        return JF_ERR_ARG;
    }
    std::vector<int32_t> nodes_new((size_t)total);
    for (int b = 0; b < n_bins; b++) {
        if (bin_ptr[b + 1] < bin_ptr[b]) {
            jf_set_error("jf_set_curve_bins: bin_ptr must be non-decreasing");
            return JF_ERR_ARG;
        }
        for (int32_t i = bin_ptr[b]; i < bin_ptr[b + 1]; i++) {
            if (bin_nodes[i] < 0 || bin_nodes[i] >= c->N) {
                jf_set_error("jf_set_curve_bins: node %d outside [0,%lld)", bin_nodes[i], (long long)c->N);
                return JF_ERR_ARG;
            }
            nodes_new[i] = c->node_old2new[bin_nodes[i]];
        }
    }
    if (c->d_bin_ptr) cudaFreeAsync(c->d_bin_ptr, c->stream);
    if (c->d_bin_nodes) cudaFreeAsync(c->d_bin_nodes, c->stream);
    if (c->d_curve_scratch) cudaFreeAsync(c->d_curve_scratch, c->stream);
    if (c->d_curves) cudaFreeAsync(c->d_curves, c->stream);
    c->d_bin_ptr = c->d_bin_nodes = nullptr;
    c->d_curve_scratch = c->d_curves = nullptr;
    int rc;
    if ((rc = jf_dev_alloc(c, &c->d_bin_ptr, (size_t)n_bins + 1))) return rc;
    if ((rc = jf_dev_alloc(c, &c->d_bin_nodes, (size_t)total))) return rc;
    if ((rc = jf_dev_alloc(c, &c->d_curve_scratch, (size_t)c->n_sys * (total > 0 ? total : 1)))) return rc;
    if ((rc = jf_dev_alloc(c, &c->d_curves, (size_t)c->n_sys * n_bins))) return rc;
    JF_CUDA(cudaMemcpy(c->d_bin_ptr, bin_ptr, ((size_t)n_bins + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (total > 0) JF_CUDA(cudaMemcpy(c->d_bin_nodes, nodes_new.data(), (size_t)total * sizeof(int32_t), cudaMemcpyHostToDevice));
    c->stats.bytes_h2d += (int64_t)((n_bins + 1 + total) * sizeof(int32_t));
    c->n_bins = n_bins;
    c->bin_total = total;
    c->curve_r_inner = r_inner;
    return JF_OK;
}

extern "C" int jf_get_curves(jf_ctx *c, double *curves) {
    if (!c || !curves) {
        jf_set_error("jf_get_curves: null argument");
        return JF_ERR_ARG;
    }
    if (c->n_bins < 1) {
        jf_set_error("jf_get_curves before jf_set_curve_bins");
        return JF_ERR_STATE;
    }
    JF_CUDA(cudaSetDevice(c->device));
    // every system, whether it stopped early or not: temporarily evaluate all by using a state-independent kernel
    dim3 grid((unsigned)c->n_bins, (unsigned)c->n_sys);
    jf_curve_kernel<<<grid, 256, 0, c->stream>>>(c->d_bin_ptr, c->d_bin_nodes, c->d_U, c->d_curve_scratch, c->d_curves, c->d_state,
                                                c->N, c->n_bins, c->bin_total > 0 ? c->bin_total : 1, c->curve_r_inner);
    JF_CUDA(cudaGetLastError());
    c->stats.kernel_launches++;
    JF_CUDA(cudaMemcpyAsync(curves, c->d_curves, (size_t)c->n_sys * c->n_bins * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    JF_CUDA(cudaStreamSynchronize(c->stream));
    c->stats.bytes_d2h += (int64_t)((size_t)c->n_sys * c->n_bins * sizeof(double));
    return JF_OK;
}
