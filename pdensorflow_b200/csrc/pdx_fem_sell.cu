// pdx_fem_sell.cu - matrix-format and ghost plumbing of the partitioned FEM path, on the device:
//   * CSR -> SELL-32 conversion next to the device assembly (pdx_fem_assemble): per-slice widths, then one thread per
//     row scatters its entries column-major into the slice and pads the tail with (col = own index, val = 0).  The
//     host-side NumPy conversion (fem_dist.csr_to_sell32) is what it replaces and is checked against.
//   * pdx_push_rows: the general ghost exchange of a row-block partition whose ghosts are NOT the two neighbours'
//     edge rows (unstructured meshes, many ranks): after a step every rank stores the rows other ranks hold as ghosts
//     straight into THEIR full-length buffers over NVLink peer memory, from an index list built once at set-up.
// The reference keeps everything on the host (gpuSolve/matrices/globalMatrices.py:54-74, 612-675) and has no
// multi-GPU path.
#include "pdx_internal.h"

namespace pdx {
namespace {

// one warp per 32-row slice: width = longest row of the slice
__global__ void __launch_bounds__(256) sell_width_kernel(int n, const int32_t* __restrict__ rowptr, int32_t* __restrict__ width) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nsl = (n + 31) >> 5;
    if (warp >= nsl) return;
    const int row = warp * 32 + lane;
    int len = row < n ? rowptr[row + 1] - rowptr[row] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if (lane == 0) width[warp] = len;
}

// one thread per (padded) row: entry k of row r of slice s lands at slice_ptr[s] + 32 k + (r mod 32)
__global__ void __launch_bounds__(256) sell_fill_kernel(int n, int row_lo, const int32_t* __restrict__ rowptr,
                                                        const int32_t* __restrict__ col, const float* __restrict__ val,
                                                        const int32_t* __restrict__ slice_ptr, int32_t* __restrict__ scol,
                                                        float* __restrict__ sval) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int nsl = (n + 31) >> 5;
    const int sl = row >> 5, lane = row & 31;
    if (sl >= nsl) return;
    const int base = slice_ptr[sl];
    const int width = (slice_ptr[sl + 1] - base) >> 5;
    const int beg = row < n ? rowptr[row] : 0, len = row < n ? rowptr[row + 1] - beg : 0;
    const int own = row_lo + min(row, n - 1);
    for (int k = 0; k < width; ++k) {
        const int dst = base + 32 * k + lane;
        scol[dst] = k < len ? col[beg + k] : own;
        sval[dst] = k < len ? val[beg + k] : 0.0f;
    }
}

__global__ void __launch_bounds__(256) push_rows_kernel(const float* __restrict__ src, int count,
                                                        const int32_t* __restrict__ rows, const int32_t* __restrict__ peer,
                                                        float* const* __restrict__ bufs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int r = rows[i];
    bufs[peer[i]][r] = src[r];
}

}  // namespace
}  // namespace pdx

using namespace pdx;

extern "C" {

int pdx_sell_slice_widths(int32_t n, const int32_t* rowptr, int32_t* widths, void* stream) {
    if (n < 0 || (n > 0 && (!rowptr || !widths))) { set_error("pdx_sell_slice_widths: bad arguments"); return 1; }
    if (n == 0) return 0;
    const int nsl = (n + 31) / 32;
    sell_width_kernel<<<(unsigned)((nsl * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, rowptr, widths);
    count_launch();
    return check_cuda(cudaGetLastError(), "sell_width_kernel launch");
}

int pdx_sell_fill(int32_t n, int32_t row_lo, const int32_t* rowptr, const int32_t* col, const float* val,
                  const int32_t* slice_ptr, int32_t* scol, float* sval, void* stream) {
    if (n < 0 || (n > 0 && (!rowptr || !col || !val || !slice_ptr || !scol || !sval))) {
        set_error("pdx_sell_fill: bad arguments");
        return 1;
    }
    if (n == 0) return 0;
    const int nsl = (n + 31) / 32;
    sell_fill_kernel<<<(unsigned)((nsl * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, row_lo, rowptr, col, val,
                                                                                             slice_ptr, scol, sval);
    count_launch();
    return check_cuda(cudaGetLastError(), "sell_fill_kernel launch");
}

int pdx_push_rows(const float* src, int32_t count, const int32_t* rows, const int32_t* peer, float* const* peer_bufs,
                  void* stream) {
    if (count < 0 || (count > 0 && (!src || !rows || !peer || !peer_bufs))) { set_error("pdx_push_rows: bad arguments"); return 1; }
    if (count == 0) return 0;
    push_rows_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, count, rows, peer, peer_bufs);
    count_launch();
    return check_cuda(cudaGetLastError(), "push_rows_kernel launch");
}

}  // extern "C"
